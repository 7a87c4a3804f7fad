// Minibatch optimal-transport coupling on the device.
//
// Reference: OTPlanSampler.get_map, torchcfm/optimal_transport.py:59-94 -- uniform marginals a = b = 1/n over two minibatches
// of equal size, squared-Euclidean cost, exact plan from POT's network simplex (pot.emd, un-vendored third-party C++; POT is
// not installed here).  For equal-size uniform marginals the exact plan is a vertex of the Birkhoff polytope, i.e. 1/n times a
// permutation matrix, so get_map is a linear assignment problem and sample_plan (:120-142) draws i uniformly and returns
// (x0[i], x1[perm[i]]).  The reference ships the matrix to the host, solves it there (n = 256) and ships the plan back; here the
// assignment is solved where the data lives, so the dynamic matcher's step needs no host round trip and can be graph-captured.
//
// Two kernels: shortest augmenting paths (below; n <= 64 and the yardstick) and an epsilon-scaling auction (further down; the
// default above 64 rows).
// Augmenting paths: shortest augmenting paths with dual variables (Jonker-Volgenant / the Hungarian method in its O(n^3) form), fp64,
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
  x0 += (size_t)blockIdx.x * n * D; x1 += (size_t)blockIdx.x * n * D;   // one CTA per problem of a batch (idff_ot_assign_batch)
  perm += (size_t)blockIdx.x * n;
  if (cost_out) cost_out += blockIdx.x;
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


// ---- epsilon-scaling auction (Bertsekas), the default above 64 rows -----------------------------------------------------------
// The augmenting-path search above is a chain of ~n^2 / 2 dependent block-wide argmins (10 ms at n = 256).  The auction is
// parallel over the unassigned rows: in a round every unassigned row i (one warp each) scans all columns for its best and
// second-best value -c(i, j) - p[j] and bids p[j1] + (w1 - w2) + eps on the best; every column takes its highest bid (ties: the
// lowest row), its previous owner becomes unassigned.  Prices only rise; a phase ends when every row owns a column, and the
// assignment then costs at most n * eps more than the optimum.  eps is scaled down by 6 from max(c) / 2 to max(c) * 2^-40, prices
// carried over, so the final plan is the exact optimum whenever the optimum beats every other plan by more than n max(c) 2^-40
// (2e-10 of the largest cost at n = 256: below the rounding of the cost sums themselves).  All arithmetic in fp64, costs
// recomputed from the coordinates; every step is deterministic (ordered atomics only).
template <int DFIX>
__global__ void __launch_bounds__(512, 1) k_ot_auction(const float* __restrict__ x0, const float* __restrict__ x1, int n, int D,
                                                       int64_t* __restrict__ perm, double* __restrict__ cost_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  x0 += (size_t)blockIdx.x * n * D; x1 += (size_t)blockIdx.x * n * D;   // one CTA per problem of a batch (idff_ot_assign_batch)
  perm += (size_t)blockIdx.x * n;
  if (cost_out) cost_out += blockIdx.x;
  double* price = reinterpret_cast<double*>(smraw);                                // [n]
  double* bidval = price + n;                                                      // [n] bid of row i this round
  unsigned long long* maxbid = reinterpret_cast<unsigned long long*>(bidval + n);  // [n] highest bid on column j (bit pattern)
  double* wred = reinterpret_cast<double*>(maxbid + n);                            // [32]
  int* owner = reinterpret_cast<int*>(wred + 32);                                  // [n] row that owns column j, -1
  int* mine = owner + n;                                                           // [n] column owned by row i, -1
  int* winner = mine + n;                                                          // [n] lowest row among the highest bidders
  int* bidcol = winner + n;                                                        // [n] column row i bid on this round
  int* list = bidcol + n;                                                          // [n] unassigned rows, in row order
  int* bidcol2 = list + n;                                                         // [n] the other list buffer
  int* ccnt = bidcol2 + n;                                                         // [32] list lengths
  double* xs0 = reinterpret_cast<double*>(ccnt + 32);                              // [n][DFIX] coordinates, converted to fp64 once
  double* xs1 = xs0 + (DFIX > 0 ? n * DFIX : 0);                                   // [n][DFIX] (the conversions run on the XU pipe)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;   // 16 warps

  if (DFIX > 0)
    for (int i = tid; i < n * DFIX; i += blockDim.x) { xs0[i] = (double)x0[i]; xs1[i] = (double)x1[i]; }
  for (int j = tid; j < n; j += blockDim.x) { price[j] = 0.0; maxbid[j] = 0ull; winner[j] = 0x7fffffff; }
  __syncthreads();
  auto cost = [&](int i, int j) -> double {
    double c = 0.0;
    if (DFIX > 0) {
#pragma unroll
      for (int d = 0; d < DFIX; ++d) {
        const double df = xs0[i * DFIX + d] - xs1[j * DFIX + d];
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
  // largest cost: the epsilon scale
  double cmax = 0.0;
  for (int e = tid; e < n * n; e += blockDim.x) cmax = fmax(cmax, cost(e / n, e % n));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, off));
  if (lane == 0) wred[warp] = cmax;
  __syncthreads();
  cmax = wred[0];
  for (int w = 1; w < nwarps; ++w) cmax = fmax(cmax, wred[w]);
  __syncthreads();
  const double eps_final = fmax(cmax * 9.094947017729282e-13, 1e-300);   // max(c) 2^-40
  double eps = fmax(cmax * 0.5, eps_final);
  const long round_cap = 64L * n + 4096;   // non-finite inputs must not hang the device: give up on a phase after this many rounds

  for (bool last = false; !last;) {
    last = eps <= eps_final;
    for (int i = tid; i < n; i += blockDim.x) { owner[i] = -1; mine[i] = -1; }
    __syncthreads();
    // the unassigned rows of the next round are collected by the rows themselves (losing bidders, displaced owners) with an
    // atomic counter: the list order varies from run to run, the result does not -- the bids of a round do not depend on each
    // other and a column's winner is chosen by value, then by row index
    for (int i = tid; i < n; i += blockDim.x) list[i] = i;
    if (tid == 0) { ccnt[0] = n; ccnt[1] = 0; }
    __syncthreads();
    int cur = 0;   // ccnt[cur]: rows in list (this round), ccnt[cur ^ 1]: counter of the next round's list
    int* lists[2] = {list, bidcol2};
    for (long round = 0; round < round_cap; ++round) {
      const int nun = ccnt[cur];
      if (nun == 0) break;
      const int* L = lists[cur];
      int* Lnext = lists[cur ^ 1];
      if (nun == 1) {
        // the common tail of a phase: ONE unassigned row.  Its bid displaces an owner, who is then the only unassigned row, and so
        // on until a free column is taken: warp 0 walks that chain alone (same bids, no block barriers)
        if (warp == 0) {
          int i = L[0];
          for (long hop = 0; hop < round_cap; ++hop) {
            double w1 = -1.0e300, w2 = -1.0e300;
            int j1 = 0x7fffffff;
            for (int j = lane; j < n; j += 32) {
              const double v = -cost(i, j) - price[j];
              if (v > w1) { w2 = w1; w1 = v; j1 = j; }
              else if (v > w2) w2 = v;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              const double o1 = __shfl_xor_sync(0xffffffffu, w1, off), o2 = __shfl_xor_sync(0xffffffffu, w2, off);
              const int oj = __shfl_xor_sync(0xffffffffu, j1, off);
              if (o1 > w1 || (o1 == w1 && oj < j1)) { w2 = fmax(w1, o2); w1 = o1; j1 = oj; }
              else w2 = fmax(w2, o1);
            }
            if (j1 >= n) { j1 = i; w2 = w1; }
            const int prev = owner[j1];
            __syncwarp();
            if (lane == 0) {
              price[j1] = price[j1] + (n > 1 ? w1 - w2 : 0.0) + eps;
              owner[j1] = i;
              mine[i] = j1;
              if (prev >= 0) mine[prev] = -1;
            }
            __syncwarp();
            if (prev < 0) break;
            i = prev;
          }
          if (lane == 0) { ccnt[cur] = 0; ccnt[cur ^ 1] = 0; }
        }
        __syncthreads();
        continue;   // the next iteration sees an empty list and ends the phase (a capped chain leaves its row unassigned)
      }
      // (1) bids: one warp per unassigned row scans all columns for its best and second-best value
      for (int u = warp; u < nun; u += nwarps) {
        const int i = L[u];
        double w1 = -1.0e300, w2 = -1.0e300;
        int j1 = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
          const double v = -cost(i, j) - price[j];
          if (v > w1) { w2 = w1; w1 = v; j1 = j; }
          else if (v > w2) w2 = v;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double o1 = __shfl_xor_sync(0xffffffffu, w1, off), o2 = __shfl_xor_sync(0xffffffffu, w2, off);
          const int oj = __shfl_xor_sync(0xffffffffu, j1, off);
          if (o1 > w1 || (o1 == w1 && oj < j1)) { w2 = fmax(w1, o2); w1 = o1; j1 = oj; }
          else w2 = fmax(w2, o1);
        }
        if (lane == 0) {
          if (j1 >= n) { j1 = i; w2 = w1; }   // nothing comparable (non-finite input): bid on the own index, no gap
          const double bid = price[j1] + (n > 1 ? w1 - w2 : 0.0) + eps;
          bidval[i] = bid;
          bidcol[i] = j1;
          atomicMax(&maxbid[j1], (unsigned long long)__double_as_longlong(bid));   // bids are positive: bit order = value order
        }
      }
      if (tid == 0) ccnt[cur ^ 1] = 0;
      __syncthreads();
      // (2) the lowest row among a column's highest bidders wins
      for (int u = tid; u < nun; u += blockDim.x) {
        const int i = L[u], j1 = bidcol[i];
        if ((unsigned long long)__double_as_longlong(bidval[i]) == maxbid[j1]) atomicMin(&winner[j1], i);
      }
      __syncthreads();
      // (3) winners take their columns at the price they bid and reset the column's bid state; losers and displaced owners
      // enter the next round's list
      for (int u = tid; u < nun; u += blockDim.x) {
        const int i = L[u], j1 = bidcol[i];
        if (winner[j1] == i) {
          const int prev = owner[j1];
          if (prev >= 0) {
            mine[prev] = -1;
            Lnext[atomicAdd(&ccnt[cur ^ 1], 1)] = prev;
          }
          owner[j1] = i;
          mine[i] = j1;
          price[j1] = bidval[i];
        } else {
          Lnext[atomicAdd(&ccnt[cur ^ 1], 1)] = i;
        }
      }
      __syncthreads();
      for (int u = tid; u < nun; u += blockDim.x) {   // (after every bidder has read winner[]: all bidders of a column write the same)
        const int j1 = bidcol[L[u]];
        maxbid[j1] = 0ull;
        winner[j1] = 0x7fffffff;
      }
      cur ^= 1;
      __syncthreads();
    }
#ifndef IDFF_OT_THETA
#define IDFF_OT_THETA 6.0
#endif
    eps = fmax(eps * (1.0 / IDFF_OT_THETA), eps_final);
  }
  __syncthreads();
  // rows left unassigned by the round cap (non-finite inputs only) take the free columns in order: still a permutation
  if (tid == 0) {
    int fj = 0;
    for (int i = 0; i < n; ++i)
      if (mine[i] < 0) {
        while (owner[fj] >= 0) ++fj;
        mine[i] = fj;
        owner[fj] = i;
      }
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) perm[i] = (int64_t)mine[i];
  if (cost_out) {   // total cost in an order that does not depend on the block size: per-row costs, then warp 0 (lane l sums rows l, l + 32, ...)
    for (int i = tid; i < n; i += blockDim.x) bidval[i] = cost(i, mine[i]);
    __syncthreads();
    if (warp == 0) {
      double c = 0.0;
      for (int i = lane; i < n; i += 32) c += bidval[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
      if (lane == 0) *cost_out = c;
    }
  }
}

static size_t auction_smem_bytes(int n, int dfix) {
  return (size_t)n * 8 * 3 + 32 * 8 + (size_t)n * 4 * 6 + 32 * 4 + (size_t)2 * n * dfix * 8 + 16;
}

static size_t smem_bytes(int n, int dfix) { return (size_t)n * 8 + 64 * sizeof(Best) + (size_t)4 * n * 4 + (size_t)n * dfix * 4 + 16; }

}  // namespace ot
}  // namespace idff

using namespace idff;

static thread_local int g_ot_threads = 0;   // diagnostic: block size of the auction kernel (0 = chosen by the batch size), idff_debug_ot_threads
extern "C" int idff_debug_ot_threads(int32_t threads) {
  IDFF_CHECK_ARG(threads == 0 || (threads >= 64 && threads <= 512 && threads % 32 == 0), "idff_debug_ot_threads: %d", threads);
  g_ot_threads = threads;
  return IDFF_OK;
}
static thread_local int g_ot_sap = 0;   // diagnostic: 1 = always the augmenting-path kernel (idff_debug_ot_sap)
extern "C" int idff_debug_ot_sap(int32_t on) {
  g_ot_sap = on ? 1 : 0;
  return IDFF_OK;
}

// P independent problems (x0, x1 [P][n][D] -> perm [P][n], cost [P]), one CTA each: the kernels are latency-bound one-CTA solvers,
// so a batch of problems -- the minibatches of the next S training steps, which do not depend on the parameters -- fills the
// machine: several CTAs per SM, 148 SMs.
extern "C" int idff_ot_assign_batch(const float* d_x0, const float* d_x1, int32_t n_problems, int32_t n, int32_t D, int64_t* d_perm,
                                    double* d_cost, void* stream) {
  IDFF_CHECK_ARG(n >= 0 && n <= ot::MAXN, "idff_ot_assign: n = %d outside [0, %d]", n, ot::MAXN);
  IDFF_CHECK_ARG(D >= 1, "idff_ot_assign: D = %d", D);
  IDFF_CHECK_ARG(n_problems >= 0 && n_problems <= (1 << 20), "idff_ot_assign_batch: n_problems = %d outside [0, 2^20]", n_problems);
  if (n == 0 || n_problems == 0) return IDFF_OK;
  IDFF_CHECK_ARG(d_x0 && d_x1 && d_perm, "idff_ot_assign: null pointer");
  PtrDeviceGuard guard(d_perm);
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid(n_problems);
  if (n > 64 && !g_ot_sap) {   // auction: parallel over the unassigned rows
    const size_t sm = ot::auction_smem_bytes(n, D == 2 ? 2 : 0);
    // one problem: 16 warps bid in parallel (latency).  A batch that fills the machine several times over: smaller CTAs, more of
    // them resident per SM (throughput) -- same rounds, same result for any block size
    int sms = 148;
    (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, guard.dev);
    // (measured, profiles/r2_ot_batch_rate.txt: n = 256, 4096 problems: 23.1 / 11.8 / 9.6 us per problem at 512 / 256 / 128 threads)
    const int threads = g_ot_threads ? g_ot_threads : n_problems <= sms ? 512 : n <= 512 ? 128 : 256;
    if (D == 2) {
      IDFF_CUDA(cudaFuncSetAttribute(ot::k_ot_auction<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      ot::k_ot_auction<2><<<grid, threads, sm, s>>>(d_x0, d_x1, n, D, d_perm, d_cost);
    } else {
      IDFF_CUDA(cudaFuncSetAttribute(ot::k_ot_auction<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      ot::k_ot_auction<0><<<grid, threads, sm, s>>>(d_x0, d_x1, n, D, d_perm, d_cost);
    }
  } else {                     // augmenting paths: small problems, and the yardstick behind idff_debug_ot_sap
    const int threads = ((n + 31) / 32) * 32;
    const size_t sm = ot::smem_bytes(n, D == 2 ? 2 : 0);
    if (D == 2) ot::k_ot_assign<2><<<grid, threads, sm, s>>>(d_x0, d_x1, n, D, d_perm, d_cost);
    else ot::k_ot_assign<0><<<grid, threads, sm, s>>>(d_x0, d_x1, n, D, d_perm, d_cost);
  }
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}

extern "C" int idff_ot_assign(const float* d_x0, const float* d_x1, int32_t n, int32_t D, int64_t* d_perm, double* d_cost,
                              void* stream) {
  return idff_ot_assign_batch(d_x0, d_x1, 1, n, D, d_perm, d_cost, stream);
}
