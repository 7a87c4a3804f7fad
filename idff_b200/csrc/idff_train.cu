// Training-side kernels: conditional path sample (K2), lambda(t), and the fused losses with their
// gradients (K2').  All are HBM-bound streaming kernels: 128-bit loads/stores, grid sized to the SM count.
// Bit-exactness contract for the path sample: the reference's fp32 operation order
// (torchcfm/conditional_flow_matching.py:56-77,98-123,148) with round-to-nearest mul/add and no contraction.
#include "idff_common.cuh"

namespace idff {

__device__ __forceinline__ float sigma_t_of(float t, float sigma, int mode) {
  if (mode == 0) return sigma;
  // ExactOT matcher: sigma * sqrt(t * (1 - t))   (conditional_flow_matching.py:231-233)
  return __fmul_rn(sigma, __fsqrt_rn(__fmul_rn(t, __fsub_rn(1.0f, t))));
}

__device__ __forceinline__ void path_elem(float x0, float x1, float t, float st, float eps, float& xt, float& ut) {
  const float mu = __fadd_rn(__fmul_rn(t, x1), __fmul_rn(__fsub_rn(1.0f, t), x0));
  xt = __fadd_rn(mu, __fmul_rn(st, eps));
  ut = __fsub_rn(x1, x0);
}

// Times of the 4 consecutive flat elements 4i .. 4i+3 (row = element / D).  DM = 1, 2: D == 1 / 2, one aligned vector load
// (equal rows share a register, so the per-row scalars -- sigma_t, the loss weight -- are computed once per row);
// DM = 4: D % 4 == 0, the four elements lie in one row; DM = 0: any D, 32-bit division (the host picks DM = 0 only when
// the flat index fits 32 bits, DM = -1 otherwise: 64-bit division).
template <int DM>
__device__ __forceinline__ void load_t4(const float* __restrict__ t, long i, long D, float (&tv)[4]) {
  if (DM == 1) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(t) + i);
    tv[0] = v.x; tv[1] = v.y; tv[2] = v.z; tv[3] = v.w;
  } else if (DM == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(t) + i);
    tv[0] = v.x; tv[1] = v.x; tv[2] = v.y; tv[3] = v.y;
  } else if (DM == 4) {
    const float v = __ldg(t + (unsigned)i / (unsigned)(D >> 2));
    tv[0] = v; tv[1] = v; tv[2] = v; tv[3] = v;
  } else if (DM == 0) {
    const unsigned e0 = 4u * (unsigned)i, Du = (unsigned)D;
    unsigned row = e0 / Du, rem = e0 - row * Du;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      tv[q] = __ldg(t + row);
      if (++rem == Du) { rem = 0; ++row; }
    }
  } else {
    const long e0 = 4 * i;
    long row = e0 / D, rem = e0 - row * D;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      tv[q] = __ldg(t + row);
      if (++rem == D) { rem = 0; ++row; }
    }
  }
}
// host side: which load_t4 variant serves (t, D, n4)
static int pick_dm(const void* t, long D, long n4) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(t);
  if (n4 >= (1L << 30)) return -1;
  if (D == 1 && (a & 15) == 0) return 1;
  if (D == 2 && (a & 7) == 0) return 2;
  if (D % 4 == 0) return 4;
  return 0;
}

// Vectorised path: no gather, (B*D) % 4 == 0 and 16-byte aligned pointers.  Each thread owns 4 consecutive
// elements per iteration; rows are recovered from the flat index (D is arbitrary).
template <int DM>
__global__ void __launch_bounds__(256, 6) k_path_sample_vec(const float4* __restrict__ x0, const float4* __restrict__ x1,
                                                            const float* __restrict__ t, const float4* __restrict__ eps,
                                                            float sigma, int mode, float4* __restrict__ xt,
                                                            float4* __restrict__ ut, long n4, long D) {
  const long stride = (long)gridDim.x * blockDim.x;
  // two independent 128-bit groups per thread and iteration: twice the loads in flight
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const long i2 = i + stride;
    const bool two = i2 < n4;
    const float4 a = __ldcs(x0 + i), b = __ldcs(x1 + i);
    const float4 e = eps ? __ldcs(eps + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a2 = a, b2 = b, e2 = e;
    if (two) {
      a2 = __ldcs(x0 + i2);
      b2 = __ldcs(x1 + i2);
      if (eps) e2 = __ldcs(eps + i2);
    }
    float tv[4], tw[4];
    load_t4<DM>(t, i, D, tv);
    if (two) load_t4<DM>(t, i2, D, tw);
    float4 o, u;
    path_elem(a.x, b.x, tv[0], sigma_t_of(tv[0], sigma, mode), e.x, o.x, u.x);
    path_elem(a.y, b.y, tv[1], sigma_t_of(tv[1], sigma, mode), e.y, o.y, u.y);
    path_elem(a.z, b.z, tv[2], sigma_t_of(tv[2], sigma, mode), e.z, o.z, u.z);
    path_elem(a.w, b.w, tv[3], sigma_t_of(tv[3], sigma, mode), e.w, o.w, u.w);
    __stcs(xt + i, o);
    if (ut) __stcs(ut + i, u);
    if (two) {
      path_elem(a2.x, b2.x, tw[0], sigma_t_of(tw[0], sigma, mode), e2.x, o.x, u.x);
      path_elem(a2.y, b2.y, tw[1], sigma_t_of(tw[1], sigma, mode), e2.y, o.y, u.y);
      path_elem(a2.z, b2.z, tw[2], sigma_t_of(tw[2], sigma, mode), e2.z, o.z, u.z);
      path_elem(a2.w, b2.w, tw[3], sigma_t_of(tw[3], sigma, mode), e2.w, o.w, u.w);
      __stcs(xt + i2, o);
      if (ut) __stcs(ut + i2, u);
    }
  }
}

// Scalar path: ragged sizes and/or the OT re-pairing gather x0[idx0[b]], x1[idx1[b]].
__global__ void __launch_bounds__(256) k_path_sample_gather(const float* __restrict__ x0, const float* __restrict__ x1,
                                                            const float* __restrict__ t, const float* __restrict__ eps,
                                                            const int64_t* __restrict__ idx0,
                                                            const int64_t* __restrict__ idx1, float sigma, int mode,
                                                            float* __restrict__ xt, float* __restrict__ ut, long n,
                                                            long D) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long row = i / D, d = i - row * D;
    const long r0 = idx0 ? idx0[row] : row, r1 = idx1 ? idx1[row] : row;
    const float tv = __ldg(t + row);
    float o, u;
    path_elem(x0[r0 * D + d], x1[r1 * D + d], tv, sigma_t_of(tv, sigma, mode), eps ? eps[i] : 0.0f, o, u);
    xt[i] = o;
    if (ut) ut[i] = u;
  }
}

__global__ void k_lambda(const float* __restrict__ t, float sigma, int mode, int dynamic, float* __restrict__ out,
                         long B) {
  // conditional_flow_matching.py:210-211: 2 * sigma_t**2 / (sigma**2 + 1e-8); _dynamic.py:213-214: 2 * sigma_t / (...)
  const float den = (float)((double)sigma * (double)sigma + 1e-8);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long)gridDim.x * blockDim.x) {
    const float st = sigma_t_of(t[i], sigma, mode);
    const float num = dynamic ? __fmul_rn(2.0f, st) : __fmul_rn(2.0f, __fmul_rn(st, st));
    out[i] = __fdiv_rn(num, den);
  }
}

// ---- losses -------------------------------------------------------------------------------------------------------
constexpr int kLossMaxBlocks = 1024;
struct LossScratch {
  unsigned int ticket;
  unsigned int pad[3];
  double part[3][kLossMaxBlocks];
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;   // valid in thread 0
}

// kind 0: flow loss; 1: score loss; 2: periodic MD loss.  One pass computes the loss terms and, when asked,
// the gradient wrt the network output `a`.  VEC = 4: 128-bit loads/stores of 4 consecutive elements (rows are
// recovered from the flat index); VEC = 1: ragged / unaligned.  Partial sums are fp64 per block, combined in block
// order by the last block to finish (deterministic).
template <int KIND>
__device__ __forceinline__ void loss_elem(float av, float b1, float x0v, float xtv, float tv, float sigma0_sq,
                                          float gk, float& s0, float& s1, float& s2, float& gout) {
  if (KIND == 0) {
    // mean(((1/(1e-2+1-t)) * (vt - x1))**2)           Autograd_example_order2.py:105
    const float wgt = __fdiv_rn(1.0f, __fsub_rn(1.01f, tv));
    const float r = __fmul_rn(wgt, __fsub_rn(av, b1));
    s0 += __fmul_rn(r, r);
    gout = 2.0f * r * wgt * gk;
  } else if (KIND == 1) {
    // mean((st - logp)**2), logp per example_order1_with_prediction_head.py:42-50
    const float omt = __fsub_rn(1.0f, tv);
    const float mu = __fadd_rn(__fmul_rn(tv, b1), __fmul_rn(omt, x0v));
    const float s2v = __fadd_rn(__fmul_rn(__fmul_rn(sigma0_sq, tv), omt), 1e-4f);
    const float diff = __fsub_rn(xtv, mu);
    const float lp = __fmul_rn(-0.5f, __fadd_rn(logf(__fmul_rn(6.283185307179586f, s2v)),
                                                __fdiv_rn(__fmul_rn(diff, diff), s2v)));
    const float r = __fsub_rn(av, lp);
    s0 += __fmul_rn(r, r);
    gout = 2.0f * r * gk;
  } else {
    // mean((vt-x1)^2) + mean((vt-(x1+360))^2) + mean((vt-(x1-360))^2)     MD_simulation.py:137-141
    const float r0 = __fsub_rn(av, b1), r1 = __fsub_rn(av, __fadd_rn(b1, 360.0f)),
                r2 = __fsub_rn(av, __fsub_rn(b1, 360.0f));
    s0 += __fmul_rn(r0, r0);
    s1 += __fmul_rn(r1, r1);
    s2 += __fmul_rn(r2, r2);
    gout = 2.0f * (r0 + r1 + r2) * gk;
  }
}

template <int KIND, int VEC, int DM>
__global__ void __launch_bounds__(256, 6) k_loss(const float* __restrict__ a, const float* __restrict__ x1,
                                              const float* __restrict__ x0, const float* __restrict__ xt,
                                              const float* __restrict__ t, float sigma0_sq, float gscale,
                                              float* __restrict__ loss, float* __restrict__ grad, long n, long D,
                                              LossScratch* __restrict__ sc) {
  __shared__ double sh[8];
  __shared__ bool last;
  const float gk = (float)(1.0 / (double)n) * gscale;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  if (VEC == 4) {
    const long n4 = n >> 2;
    const float4 *a4 = reinterpret_cast<const float4*>(a), *x14 = reinterpret_cast<const float4*>(x1),
                 *x04 = reinterpret_cast<const float4*>(x0), *xt4 = reinterpret_cast<const float4*>(xt);
    float4* g4 = reinterpret_cast<float4*>(grad);
    int cnt = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 av = __ldcs(a4 + i), bv = __ldcs(x14 + i);
      float4 x0v = make_float4(0.f, 0.f, 0.f, 0.f), xtv = x0v;
      if (KIND == 1) { x0v = __ldcs(x04 + i); xtv = __ldcs(xt4 + i); }
      float tv[4] = {0.f, 0.f, 0.f, 0.f};
      if (KIND != 2) load_t4<DM>(t, i, D, tv);
      float4 go;
      loss_elem<KIND>(av.x, bv.x, x0v.x, xtv.x, tv[0], sigma0_sq, gk, s0, s1, s2, go.x);
      loss_elem<KIND>(av.y, bv.y, x0v.y, xtv.y, tv[1], sigma0_sq, gk, s0, s1, s2, go.y);
      loss_elem<KIND>(av.z, bv.z, x0v.z, xtv.z, tv[2], sigma0_sq, gk, s0, s1, s2, go.z);
      loss_elem<KIND>(av.w, bv.w, x0v.w, xtv.w, tv[3], sigma0_sq, gk, s0, s1, s2, go.w);
      if (grad) __stcs(g4 + i, go);
      if (++cnt == 2) {   // flush the short fp32 runs (8 terms) into fp64
        d0 += s0; d1 += s1; d2 += s2;
        s0 = s1 = s2 = 0.f;
        cnt = 0;
      }
    }
  } else {
    int cnt = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const long row = i / D;
      float go;
      loss_elem<KIND>(a[i], x1[i], KIND == 1 ? x0[i] : 0.f, KIND == 1 ? xt[i] : 0.f, KIND != 2 ? __ldg(t + row) : 0.f,
                      sigma0_sq, gk, s0, s1, s2, go);
      if (grad) grad[i] = go;
      if (++cnt == 8) {
        d0 += s0; d1 += s1; d2 += s2;
        s0 = s1 = s2 = 0.f;
        cnt = 0;
      }
    }
  }
  d0 += s0; d1 += s1; d2 += s2;
  const double b0 = block_sum(d0, sh);
  const double b1s = KIND == 2 ? block_sum(d1, sh) : 0.0;
  const double b2s = KIND == 2 ? block_sum(d2, sh) : 0.0;
  if (threadIdx.x == 0) {
    sc->part[0][blockIdx.x] = b0;
    sc->part[1][blockIdx.x] = b1s;
    sc->part[2][blockIdx.x] = b2s;
    __threadfence();
    last = atomicAdd(&sc->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      t0 += sc->part[0][b]; t1 += sc->part[1][b]; t2 += sc->part[2][b];
    }
    t0 = block_sum(t0, sh);
    t1 = block_sum(t1, sh);
    t2 = block_sum(t2, sh);
    if (threadIdx.x == 0) {
      const double dn = (double)n;
      *loss = KIND == 2 ? __fadd_rn(__fadd_rn((float)(t0 / dn), (float)(t1 / dn)), (float)(t2 / dn)) : (float)(t0 / dn);
      sc->ticket = 0;   // ready for the next call on this scratch
    }
  }
}

static int stream_grid(long work_items, int threads, int device, int blocks_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  long want = (work_items + threads - 1) / threads;
  long cap = (long)sms * blocks_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace idff

using namespace idff;

extern "C" int idff_path_sample(const float* d_x0, const float* d_x1, const float* d_t, const float* d_eps,
                                const int64_t* d_idx0, const int64_t* d_idx1, float sigma, int32_t sigma_mode,
                                float* d_xt, float* d_ut, int64_t B, int64_t D, void* stream) {
  IDFF_CHECK_ARG(B >= 0 && D >= 1, "idff_path_sample: bad shape B=%lld D=%lld", (long long)B, (long long)D);
  if (B == 0) return IDFF_OK;
  IDFF_CHECK_ARG(d_x0 && d_x1 && d_t && d_xt, "idff_path_sample: null pointer");
  IDFF_CHECK_ARG(sigma_mode == 0 || sigma_mode == 1, "idff_path_sample: sigma_mode %d", sigma_mode);
  PtrDeviceGuard guard(d_xt);
  const int dev = guard.dev;
  const long n = (long)B * (long)D;
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = !d_idx0 && !d_idx1 && (n % 4 == 0) && al16(d_x0) && al16(d_x1) && al16(d_eps) && al16(d_xt) && al16(d_ut);
  if (vec) {
    const long n4 = n / 4;
    const int grid = stream_grid((n4 + 1) / 2, 256, dev, 6);
    const int dm = pick_dm(d_t, (long)D, n4);
#define IDFF_PS_LAUNCH(DM)                                                                                   \
  k_path_sample_vec<DM><<<grid, 256, 0, (cudaStream_t)stream>>>(                                             \
      reinterpret_cast<const float4*>(d_x0), reinterpret_cast<const float4*>(d_x1), d_t,                     \
      reinterpret_cast<const float4*>(d_eps), sigma, sigma_mode, reinterpret_cast<float4*>(d_xt),            \
      reinterpret_cast<float4*>(d_ut), n4, (long)D)
    if (dm == 1) IDFF_PS_LAUNCH(1); else if (dm == 2) IDFF_PS_LAUNCH(2); else if (dm == 4) IDFF_PS_LAUNCH(4);
    else if (dm == 0) IDFF_PS_LAUNCH(0); else IDFF_PS_LAUNCH(-1);
#undef IDFF_PS_LAUNCH
  } else {
    const int grid = stream_grid(n, 256, dev, 8);
    k_path_sample_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x0, d_x1, d_t, d_eps, d_idx0, d_idx1, sigma,
                                                                  sigma_mode, d_xt, d_ut, n, (long)D);
  }
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}

extern "C" int idff_compute_lambda(const float* d_t, float sigma, int32_t sigma_mode, int32_t dynamic, float* d_out,
                                   int64_t B, void* stream) {
  if (B == 0) return IDFF_OK;
  IDFF_CHECK_ARG(d_t && d_out && B > 0, "idff_compute_lambda: bad argument");
  PtrDeviceGuard guard(d_out);
  const int dev = guard.dev;
  k_lambda<<<stream_grid(B, 256, dev, 4), 256, 0, (cudaStream_t)stream>>>(d_t, sigma, sigma_mode, dynamic, d_out, (long)B);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}

extern "C" int64_t idff_loss_scratch_bytes(void) { return (int64_t)sizeof(LossScratch); }

extern "C" int idff_loss(int32_t kind, const float* d_a, const float* d_x1, const float* d_x0, const float* d_xt,
                         const float* d_t, float sigma0_sq, float grad_scale, float* d_loss, float* d_grad, int64_t B,
                         int64_t D, void* d_scratch, void* stream) {
  IDFF_CHECK_ARG(kind >= 0 && kind <= 2, "idff_loss: kind %d", kind);
  IDFF_CHECK_ARG(B >= 1 && D >= 1, "idff_loss: empty batch (the reference's mean over zero elements is NaN)");
  IDFF_CHECK_ARG(d_a && d_x1 && d_loss && d_scratch, "idff_loss: null pointer");
  IDFF_CHECK_ARG(kind == 2 || d_t, "idff_loss: kind %d needs t", kind);
  IDFF_CHECK_ARG(kind != 1 || (d_x0 && d_xt), "idff_loss: the score loss needs x0 and xt");
  PtrDeviceGuard guard(d_loss);
  const int dev = guard.dev;
  const long n = (long)B * (long)D;
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (n % 4 == 0) && al16(d_a) && al16(d_x1) && al16(d_x0) && al16(d_xt) && al16(d_grad);
  int grid = stream_grid(vec ? n / 4 : n, 256, dev, 6);
  if (grid > kLossMaxBlocks) grid = kLossMaxBlocks;
  LossScratch* sc = reinterpret_cast<LossScratch*>(d_scratch);
  cudaStream_t s = (cudaStream_t)stream;
  const int dm = (vec && kind != 2) ? pick_dm(d_t, (long)D, n / 4) : -1;
#define IDFF_LOSS_LAUNCH(K, V, DM) \
  k_loss<K, V, DM><<<grid, 256, 0, s>>>(d_a, d_x1, d_x0, d_xt, d_t, sigma0_sq, grad_scale, d_loss, d_grad, n, (long)D, sc)
#define IDFF_LOSS_DM(K)                                                                                       \
  do {                                                                                                        \
    if (dm == 1) IDFF_LOSS_LAUNCH(K, 4, 1); else if (dm == 2) IDFF_LOSS_LAUNCH(K, 4, 2);                      \
    else if (dm == 4) IDFF_LOSS_LAUNCH(K, 4, 4); else if (dm == 0) IDFF_LOSS_LAUNCH(K, 4, 0);                 \
    else IDFF_LOSS_LAUNCH(K, 4, -1);                                                                          \
  } while (0)
  if (vec) {
    if (kind == 0) IDFF_LOSS_DM(0); else if (kind == 1) IDFF_LOSS_DM(1); else IDFF_LOSS_LAUNCH(2, 4, -1);
  } else {
    if (kind == 0) IDFF_LOSS_LAUNCH(0, 1, -1); else if (kind == 1) IDFF_LOSS_LAUNCH(1, 1, -1); else IDFF_LOSS_LAUNCH(2, 1, -1);
  }
#undef IDFF_LOSS_DM
#undef IDFF_LOSS_LAUNCH
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}
