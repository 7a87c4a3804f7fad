// Minibatch optimal-transport coupling on the device.
//
// Reference: OTPlanSampler.get_map, torchcfm/optimal_transport.py:59-94 -- uniform marginals a = b = 1/n over two minibatches
// of equal size, squared-Euclidean cost, exact plan from POT's network simplex (pot.emd, un-vendored third-party C++; POT is
// not installed here).  For equal-size uniform marginals the exact plan is a vertex of the Birkhoff polytope, i.e. 1/n times a
// permutation matrix, so get_map is a linear assignment problem and sample_plan (:120-142) draws i uniformly and returns
// (x0[i], x1[perm[i]]).  The reference ships the matrix to the host, solves it there (n = 256) and ships the plan back; here the
// assignment is solved where the data lives, so the dynamic matcher's step needs no host round trip and can be graph-captured.
//
// Algorithm: shortest augmenting paths with dual variables (Jonker-Volgenant / the Hungarian method in its O(n^3) form), fp64,
// exact, after a greedy column-reduction start.  One CTA, one thread per column: a thread keeps its column's tentative distance,
// visited flag and dual v in registers; every search step is one relaxation of all columns against the row just reached (cost
// recomputed from the coordinates, never materialised) plus one block-wide argmin (warp shuffles, one shared-memory exchange,
// one barrier per step).
#include "idff_common.cuh"

namespace idff {
namespace ot {

constexpr int MAXN = 1024;

struct Best {
  double val;
  int col;    // column index (MAXN: none)
  int pad;
};

// (distance, column) order: lowest distance, then lowest column index -- any minimal column is a valid next Dijkstra vertex
__device__ __forceinline__ bool better(double v, int c, double bv, int bc) { return v < bv || (v == bv && c < bc); }

__device__ __forceinline__ void warp_argmin(double& bv, int& bc) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
    const int oc = __shfl_xor_sync(0xffffffffu, bc, off);
    if (better(ov, oc, bv, bc)) { bv = ov; bc = oc; }
  }
}

// DFIX > 0: D == DFIX, both point sets are staged on chip (x1[j] in registers, x0 in shared memory); DFIX == 0: any D, read
// through L1.
template <int DFIX>
__global__ void __launch_bounds__(MAXN, 1) k_ot_assign(const float* __restrict__ x0, const float* __restrict__ x1, int n, int D,
                                                       int64_t* __restrict__ perm, double* __restrict__ cost_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* u = reinterpret_cast<double*>(smraw);                    // [n] row duals
  Best* wbest = reinterpret_cast<Best*>(u + n);                     // [2][32]
  int* col4row = reinterpret_cast<int*>(wbest + 64);                // [n]
  int* row4col = col4row + n;                                        // [n]
  int* path = row4col + n;                                           // [n]
  int* claim = path + n;                                             // [n] greedy start: lowest column that wants row i
  float* x0s = reinterpret_cast<float*>(claim + n);                 // [n][DFIX]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
  const int j = tid;
  const bool has = j < n;
  const double INF = 1.0e300;

  float x1r[DFIX > 0 ? DFIX : 1];
  if (DFIX > 0 && has) {
#pragma unroll
    for (int d = 0; d < DFIX; ++d) {
      x1r[d] = x1[(size_t)j * DFIX + d];
      x0s[j * DFIX + d] = x0[(size_t)j * DFIX + d];
    }
  }
  if (has) {
    u[j] = 0.0;
    col4row[j] = -1;
    row4col[j] = -1;
    path[j] = -1;
    claim[j] = MAXN;
  }
  __syncthreads();

  auto cost = [&](int i) -> double {
    double c = 0.0;
    if (DFIX > 0) {
#pragma unroll
      for (int d = 0; d < DFIX; ++d) {
        const double df = (double)x0s[i * DFIX + d] - (double)x1r[d];
        c = fma(df, df, c);
      }
    } else {
      for (int d = 0; d < D; ++d) {
        const double df = (double)x0[(size_t)i * D + d] - (double)x1[(size_t)j * D + d];
        c = fma(df, df, c);
      }
    }
    return c;
  };

  // Greedy start (column reduction): v[j] = min_i c(i, j) with u = 0 is dual feasible and makes every column's cheapest edge
  // tight; a row takes the lowest column that wants it.  Only the rows left without a column need an augmenting path.
  double v = 0.0;
  int want = -1;
  if (has) {
    double best = INF;
    for (int i = 0; i < n; ++i) {
      const double c = cost(i);
      if (c < best) { best = c; want = i; }
    }
    v = best;
    atomicMin(&claim[want], j);
  }
  __syncthreads();
  if (has && claim[want] == j) {
    row4col[j] = want;
    col4row[want] = j;
  }
  __syncthreads();

  int it = 0;
  for (int cur = 0; cur < n; ++cur) {
    if (col4row[cur] >= 0) continue;   // matched by the greedy start (uniform across the block)
    double shortest = INF;
    bool visited = false;
    const int my_row = has ? row4col[j] : -1;
    int i = cur, sink = -1;
    double min_val = 0.0;
    while (sink < 0) {
      if (has && !visited) {
        const double r = min_val + cost(i) - u[i] - v;
        if (r < shortest) {
          shortest = r;
          path[j] = i;
        }
      }
      // block argmin over the unvisited columns: warp shuffles, one shared-memory exchange, one barrier
      double bv = (has && !visited) ? shortest : INF;
      int bc = (has && !visited) ? j : MAXN;
      warp_argmin(bv, bc);
      Best* slot = wbest + (it & 1) * 32;
      if (lane == 0) { slot[warp].val = bv; slot[warp].col = bc; }
      __syncthreads();
      bv = lane < nwarps ? slot[lane].val : INF;
      bc = lane < nwarps ? slot[lane].col : MAXN;
      warp_argmin(bv, bc);
      ++it;
      min_val = bv;
      if (j == bc) visited = true;
      const int br = row4col[bc];
      if (br < 0) sink = bc; else i = br;
    }
    // dual update (uses the assignment as it was during the search)
    if (has && visited) {
      const double delta = min_val - shortest;
      if (j != sink) u[my_row] += delta;
      v -= delta;
    }
    if (tid == 0) u[cur] += min_val;
    __syncthreads();   // path[] complete, u[] updated, every thread has read row4col[] for this search
    if (tid == 0) {    // augment along the path
      int jj = sink;
      while (true) {
        const int ii = path[jj];
        row4col[jj] = ii;
        const int nxt = col4row[ii];
        col4row[ii] = jj;
        jj = nxt;
        if (ii == cur) break;
      }
    }
    __syncthreads();
  }
  if (has) perm[j] = (int64_t)col4row[j];
  if (cost_out) {   // total cost, fixed order: thread j contributes cost(row4col[j], j)
    double c = has ? cost(row4col[j]) : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    __syncthreads();
    double* red = reinterpret_cast<double*>(wbest);
    if (lane == 0) red[warp] = c;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < nwarps; ++w) t += red[w];
      *cost_out = t;
    }
  }
}

static size_t smem_bytes(int n, int dfix) { return (size_t)n * 8 + 64 * sizeof(Best) + (size_t)4 * n * 4 + (size_t)n * dfix * 4 + 16; }

}  // namespace ot
}  // namespace idff

using namespace idff;

extern "C" int idff_ot_assign(const float* d_x0, const float* d_x1, int32_t n, int32_t D, int64_t* d_perm, double* d_cost,
                              void* stream) {
  IDFF_CHECK_ARG(n >= 0 && n <= ot::MAXN, "idff_ot_assign: n = %d outside [0, %d]", n, ot::MAXN);
  IDFF_CHECK_ARG(D >= 1, "idff_ot_assign: D = %d", D);
  if (n == 0) return IDFF_OK;
  IDFF_CHECK_ARG(d_x0 && d_x1 && d_perm, "idff_ot_assign: null pointer");
  const int threads = ((n + 31) / 32) * 32;
  const size_t sm = ot::smem_bytes(n, D == 2 ? 2 : 0);
  cudaStream_t s = (cudaStream_t)stream;
  if (D == 2) ot::k_ot_assign<2><<<1, threads, sm, s>>>(d_x0, d_x1, n, D, d_perm, d_cost);
  else ot::k_ot_assign<0><<<1, threads, sm, s>>>(d_x0, d_x1, n, D, d_perm, d_cost);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}
