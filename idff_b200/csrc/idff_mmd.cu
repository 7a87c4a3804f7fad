// K3: RBF-kernel MMD partial sums (compute_mmd_rbf, Autograd_example_order2.py:260-278; compute_mmd,
// MD_simulation.py:50-86).  S(A,B) = sum_i sum_j exp(-|A_i - B_j|^2 / (2 sigma^2)) for the three pairings
// (x,x), (y,y), (x,y); the biased V-statistic is Sxx/N^2 + Syy/M^2 - 2 Sxy/(N M).
//
// The squared distance is evaluated directly as sum_d (a_d - b_d)^2 in fp32 (the reference's cdist uses the
// |a|^2+|b|^2-2ab expansion, which cancels catastrophically on degree-valued data); exponentials go through
// MUFU.EX2 on coordinates pre-scaled by sqrt(log2(e)/(2 sigma^2)).  Per-thread fp32 partial sums over short
// runs, fp64 from the warp level up.  Bound: D <= 4 -> MUFU.EX2 issue rate (16/clk/SM); larger D -> FMA pipe.
#include "idff_common.cuh"

namespace idff {

struct MmdJob {
  const float* a;   // rows [a_begin, a_end) of this set ...
  const float* b;   // ... against all nb rows of this set
  long a_begin, a_end, nb;
  int symmetric;   // a == b and the rows cover the whole set: evaluate only the upper block triangle
};
struct MmdJobs {
  MmdJob job[3];
  float scale;   // sqrt(log2(e) / (2 sigma^2))
  int D;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void block_accumulate(double v, double* dst) {
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
    atomicAdd(dst, r);
  }
}

// ---- small D (<= 4): A points in registers, B chunk broadcast from shared memory --------------------------------
constexpr int kSmallRA = 2;        // A rows per thread
constexpr int kSmallBC = 256 * kSmallRA;   // B rows per CTA (= A rows per CTA: square blocks, so symmetry is per block)
template <int D>
__device__ __forceinline__ void mmd_small_body(const MmdJob& jb, const float scale, double* __restrict__ dst) {
  const long a0 = jb.a_begin + (long)blockIdx.x * (256 * kSmallRA);
  const long b0 = (long)blockIdx.y * kSmallBC;
  if (a0 >= jb.a_end || b0 >= jb.nb) return;
  // K(a,b) = K(b,a): blocks below the diagonal are skipped, blocks above it count twice
  if (jb.symmetric && blockIdx.y < blockIdx.x) return;
  const double weight = (jb.symmetric && blockIdx.y > blockIdx.x) ? 2.0 : 1.0;
  __shared__ __align__(16) float bs[kSmallBC * D];
  const long bcnt = jb.nb - b0 < kSmallBC ? jb.nb - b0 : kSmallBC;
  for (long i = threadIdx.x; i < bcnt * D; i += 256) bs[i] = jb.b[b0 * D + i] * scale;
  float av[kSmallRA][D];
  bool valid[kSmallRA];
#pragma unroll
  for (int r = 0; r < kSmallRA; ++r) {
    const long row = a0 + threadIdx.x + (long)r * 256;
    valid[r] = row < jb.a_end;
#pragma unroll
    for (int d = 0; d < D; ++d) av[r][d] = valid[r] ? jb.a[row * D + d] * scale : 0.0f;
  }
  __syncthreads();
  double tot = 0.0;
  for (long j0 = 0; j0 < bcnt; j0 += 256) {
    const int jn = (int)(bcnt - j0 < 256 ? bcnt - j0 : 256);
    float acc[kSmallRA];
#pragma unroll
    for (int r = 0; r < kSmallRA; ++r) acc[r] = 0.0f;
#pragma unroll 4
    for (int j = 0; j < jn; ++j) {
      float bv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) bv[d] = bs[(j0 + j) * D + d];
#pragma unroll
      for (int r = 0; r < kSmallRA; ++r) {
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float df = av[r][d] - bv[d];
          s = fmaf(-df, df, s);
        }
        acc[r] += ex2_approx(s);
      }
    }
#pragma unroll
    for (int r = 0; r < kSmallRA; ++r)
      if (valid[r]) tot += (double)acc[r];
  }
  block_accumulate(tot * weight, dst);
}
template <int D>
__global__ void __launch_bounds__(256) k_mmd_small(const __grid_constant__ MmdJobs jobs, double* __restrict__ sums) {
  mmd_small_body<D>(jobs.job[blockIdx.z], jobs.scale, sums + blockIdx.z);
}
// gamma sweeps: one fixed set x [N][D] against G generated sets y_g [M][D] (contiguous).  blockIdx.z = 0: (x, x);
// 1 + 2g: (y_g, y_g); 2 + 2g: (x, y_g).  sums[blockIdx.z] receives the pair sum.
template <int D>
__global__ void __launch_bounds__(256) k_mmd_small_sweep(const float* __restrict__ x, long N, const float* __restrict__ y, long M,
                                                         float scale, double* __restrict__ sums) {
  const int z = blockIdx.z;
  MmdJob jb;
  if (z == 0) {
    jb = {x, x, 0, N, N, 1};
  } else {
    const float* yg = y + (size_t)((z - 1) >> 1) * (size_t)M * D;
    if ((z - 1) & 1) jb = {x, yg, 0, N, M, 0};
    else jb = {yg, yg, 0, M, M, 1};
  }
  mmd_small_body<D>(jb, scale, sums + z);
}

// ---- any D: 64 x 64 pair tile per iteration, 4 x 4 pairs per thread, coordinates staged in chunks of 16 ---------
constexpr int kBigBC = 1024;       // B rows per CTA
__global__ void __launch_bounds__(256) k_mmd_big(const __grid_constant__ MmdJobs jobs, double* __restrict__ sums) {
  const MmdJob& jb = jobs.job[blockIdx.z];
  const int D = jobs.D;
  const long a0 = jb.a_begin + (long)blockIdx.x * 64;
  const long bbase = (long)blockIdx.y * kBigBC;
  if (a0 >= jb.a_end || bbase >= jb.nb) return;
  __shared__ float as[16][64 + 4];
  __shared__ float bsm[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long bend = bbase + kBigBC < jb.nb ? bbase + kBigBC : jb.nb;
  double tot = 0.0;
  for (long b0 = bbase; b0 < bend; b0 += 64) {
    float d2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) d2[i][j] = 0.0f;
    for (int k0 = 0; k0 < D; k0 += 16) {
      __syncthreads();
      for (int i = threadIdx.x; i < 16 * 64; i += 256) {
        const int r = i >> 4, k = i & 15;   // consecutive threads walk along a row (coalesced for D >= 16)
        const long ra = a0 + r, rb = b0 + r;
        as[k][r] = (ra < jb.a_end && k0 + k < D) ? jb.a[ra * D + k0 + k] * jobs.scale : 0.0f;
        bsm[k][r] = (rb < bend && k0 + k < D) ? jb.b[rb * D + k0 + k] * jobs.scale : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a4[4], b4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a4[i] = as[k][ty * 4 + i]; b4[i] = bsm[k][tx * 4 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float df = a4[i] - b4[j];
            d2[i][j] = fmaf(-df, df, d2[i][j]);
          }
      }
    }
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (a0 + ty * 4 + i < jb.a_end && b0 + tx * 4 + j < bend) acc += ex2_approx(d2[i][j]);
    tot += (double)acc;
  }
  block_accumulate(tot, sums + blockIdx.z);
}

}  // namespace idff

using namespace idff;

extern "C" int idff_mmd_rbf(const float* d_x, int64_t N, const float* d_y, int64_t M, int32_t D, float sigma_k,
                            double* d_sums, int64_t x_row_begin, int64_t x_row_end, int64_t y_row_begin,
                            int64_t y_row_end, void* stream) {
  IDFF_CHECK_ARG(d_x && d_y && d_sums, "idff_mmd_rbf: null pointer");
  IDFF_CHECK_ARG(N >= 1 && M >= 1 && D >= 1, "idff_mmd_rbf: empty set (N=%lld M=%lld D=%d)", (long long)N, (long long)M, D);
  IDFF_CHECK_ARG(sigma_k > 0.0f, "idff_mmd_rbf: sigma_k must be positive");
  IDFF_CHECK_ARG(0 <= x_row_begin && x_row_begin <= x_row_end && x_row_end <= N, "idff_mmd_rbf: bad x row range");
  IDFF_CHECK_ARG(0 <= y_row_begin && y_row_begin <= y_row_end && y_row_end <= M, "idff_mmd_rbf: bad y row range");
  PtrDeviceGuard guard(d_sums);
  MmdJobs jobs;
  const bool small = D <= 4;   // the symmetric shortcut is implemented by the small-D kernel
  jobs.job[0] = {d_x, d_x, (long)x_row_begin, (long)x_row_end, (long)N, small && x_row_begin == 0 && x_row_end == N};
  jobs.job[1] = {d_y, d_y, (long)y_row_begin, (long)y_row_end, (long)M, small && y_row_begin == 0 && y_row_end == M};
  jobs.job[2] = {d_x, d_y, (long)x_row_begin, (long)x_row_end, (long)M, 0};
  jobs.scale = (float)sqrt(1.4426950408889634 / (2.0 * (double)sigma_k * (double)sigma_k));
  jobs.D = D;
  const long amax = (x_row_end - x_row_begin) > (y_row_end - y_row_begin) ? (x_row_end - x_row_begin) : (y_row_end - y_row_begin);
  const long bmax = N > M ? N : M;
  if (amax == 0) return IDFF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (D <= 4) {
    dim3 grid((unsigned)((amax + 256 * kSmallRA - 1) / (256 * kSmallRA)), (unsigned)((bmax + kSmallBC - 1) / kSmallBC), 3);
    IDFF_CHECK_ARG(grid.y <= 65535, "idff_mmd_rbf: set too large (%ld rows)", bmax);
    switch (D) {
      case 1: k_mmd_small<1><<<grid, 256, 0, s>>>(jobs, d_sums); break;
      case 2: k_mmd_small<2><<<grid, 256, 0, s>>>(jobs, d_sums); break;
      case 3: k_mmd_small<3><<<grid, 256, 0, s>>>(jobs, d_sums); break;
      default: k_mmd_small<4><<<grid, 256, 0, s>>>(jobs, d_sums); break;
    }
  } else {
    dim3 grid((unsigned)((amax + 63) / 64), (unsigned)((bmax + kBigBC - 1) / kBigBC), 3);
    IDFF_CHECK_ARG(grid.y <= 65535, "idff_mmd_rbf: set too large (%ld rows)", bmax);
    k_mmd_big<<<grid, 256, 0, s>>>(jobs, d_sums);
  }
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}

extern "C" int idff_mmd_rbf_sweep(const float* d_x, int64_t N, const float* d_y, int64_t M, int32_t n_sets, int32_t D,
                                  float sigma_k, double* d_sums, void* stream) {
  IDFF_CHECK_ARG(d_x && d_y && d_sums, "idff_mmd_rbf_sweep: null pointer");
  IDFF_CHECK_ARG(N >= 1 && M >= 1 && n_sets >= 1 && n_sets <= 32767, "idff_mmd_rbf_sweep: bad sizes (N=%lld M=%lld sets=%d)",
                 (long long)N, (long long)M, n_sets);
  IDFF_CHECK_ARG(sigma_k > 0.0f, "idff_mmd_rbf_sweep: sigma_k must be positive");
  if (D < 1 || D > 4) {
    set_error("idff_mmd_rbf_sweep: D = %d (the batched kernel serves D <= 4; call idff_mmd_rbf per set)", D);
    return IDFF_EUNSUPPORTED;
  }
  PtrDeviceGuard guard(d_sums);
  const float scale = (float)sqrt(1.4426950408889634 / (2.0 * (double)sigma_k * (double)sigma_k));
  const long amax = N > M ? N : M;
  dim3 grid((unsigned)((amax + 256 * kSmallRA - 1) / (256 * kSmallRA)), (unsigned)((amax + kSmallBC - 1) / kSmallBC),
            (unsigned)(1 + 2 * n_sets));
  IDFF_CHECK_ARG(grid.y <= 65535, "idff_mmd_rbf_sweep: set too large (%ld rows)", amax);
  cudaStream_t s = (cudaStream_t)stream;
  switch (D) {
    case 1: k_mmd_small_sweep<1><<<grid, 256, 0, s>>>(d_x, (long)N, d_y, (long)M, scale, d_sums); break;
    case 2: k_mmd_small_sweep<2><<<grid, 256, 0, s>>>(d_x, (long)N, d_y, (long)M, scale, d_sums); break;
    case 3: k_mmd_small_sweep<3><<<grid, 256, 0, s>>>(d_x, (long)N, d_y, (long)M, scale, d_sums); break;
    default: k_mmd_small_sweep<4><<<grid, 256, 0, s>>>(d_x, (long)N, d_y, (long)M, scale, d_sums); break;
  }
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}
