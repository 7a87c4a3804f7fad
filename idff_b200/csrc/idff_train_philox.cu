// Path sample with the noise drawn INSIDE the kernel (BASELINE north_star (2): x_t, u_t and the noise in one fused, HBM-bound
// kernel).  The reference draws eps = torch.randn_like(x0) (torchcfm/conditional_flow_matching.py:150-151,187) and then
// forms xt from it (:98-123): on a GPU that is one ATen RNG kernel writing eps (4 D bytes per row) plus a kernel reading
// it back.  Here eps never touches HBM unless the caller asks for it (return_noise=True): 36 instead of 44 B per row at D = 2.
//
// Bit-exact noise layout.  ATen fills a contiguous fp32 tensor with distribution_elementwise_grid_stride_kernel
// (ATen/native/cuda/DistributionTemplates.h): block 256, grid = min(SMs * (maxThreadsPerSM / 256), ceil(numel / 256)),
// thread idx owns Philox subsequence idx (curand_init(seed, idx, offset)); loop iteration j draws curand_normal4 and element
// idx + S (4 j + ii), S = grid * 256, receives component ii.  The generator's offset then advances by
// ((numel - 1) / (S * 4) + 1) * 4.  This kernel walks exactly that schedule with cuRAND's own device functions (the Philox4x32-10
// block function and the Box-Muller on logf / sqrtf / __sincosf that curand_normal4 is made of), so for the same (seed, offset)
// every element sees the number randn_like would have written; idff_philox_plan tells the caller the grid and the offset increment to book on the
// torch generator.  xt / ut use the reference's fp32 operation order (no contraction), as in idff_train.cu.
#include <curand_kernel.h>

#include "idff_common.cuh"

namespace idff {

__device__ __forceinline__ float sigma_t_of_p(float t, float sigma, int mode) {
  if (mode == 0) return sigma;
  return __fmul_rn(sigma, __fsqrt_rn(__fmul_rn(t, __fsub_rn(1.0f, t))));   // ExactOT matcher, conditional_flow_matching.py:231-233
}

// One element of the schedule: element e receives normal `eps`.
template <int DM>
__device__ __forceinline__ void philox_elem(unsigned e, float a, float b, float tt, float eps, float sigma, int mode,
                                            float* __restrict__ xt, float* __restrict__ ut, float* __restrict__ eps_out) {
  const float mu = __fadd_rn(__fmul_rn(tt, b), __fmul_rn(__fsub_rn(1.0f, tt), a));
  __stcs(xt + e, __fadd_rn(mu, __fmul_rn(sigma_t_of_p(tt, sigma, mode), eps)));
  if (ut) __stcs(ut + e, __fsub_rn(b, a));
  if (eps_out) __stcs(eps_out + e, eps);
}

// Thread idx is Philox subsequence idx: key = seed, counter = (offset / 4 + iteration, idx) -- what curand_init(seed, idx, offset)
// followed by one curand_normal4 per iteration walks when offset is a multiple of 4 (torch's always is: the generator state
// machine of curand4 then only ever returns whole 4-word outputs, so it is skipped here and the raw Philox4x32-10 block function
// and Box-Muller of cuRAND's headers are called directly).  32-bit element indices (numel < 2^31, and the rounded-up schedule
// stays below 2^32); every iteration but the last has all four of its elements inside the tensor and runs without bounds checks.
template <int DM>   // DM = 1, 2: D == 1 / 2; 0: any D
__global__ void __launch_bounds__(256, 4) k_path_sample_philox(const float* __restrict__ x0, const float* __restrict__ x1,
                                                               const float* __restrict__ t, unsigned long long seed,
                                                               unsigned long long offset, float sigma, int mode,
                                                               float* __restrict__ xt, float* __restrict__ ut,
                                                               float* __restrict__ eps_out, unsigned n, unsigned D) {
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  const unsigned S = gridDim.x * 256u;
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  unsigned long long c01 = offset >> 2;                 // counter words 0, 1; words 2, 3 = the subsequence (idx, 0)
  const unsigned iters = (n - 1u) / (4u * S) + 1u;      // the same for every thread (ATen's rounded-up loop bound)
  auto row_of = [&](unsigned e) -> unsigned { return DM == 1 ? e : (DM == 2 ? e >> 1 : e / D); };
  // (Tried on this loop, same 127 us at 2^24 x 2 each time, profiles/r2_ncu_k_path_sample_philox.txt: loads issued one iteration
  // ahead with the values moved, or with two register sets and the loop unrolled by two; 8 instead of 4 resident blocks per SM.)
  unsigned base = idx;
  for (unsigned it = 0; it + 1u < iters; ++it, base += 4u * S, ++c01) {
    float a[4], b[4], tv[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {   // the four elements of this iteration: loads in flight during the Philox / Box-Muller arithmetic
      const unsigned e = base + S * ii;
      a[ii] = __ldcs(x0 + e);
      b[ii] = __ldcs(x1 + e);
      tv[ii] = __ldg(t + row_of(e));
    }
    const uint4 r = curand_Philox4x32_10(make_uint4((unsigned)c01, (unsigned)(c01 >> 32), idx, 0u), key);
    const float2 n0 = _curand_box_muller(r.x, r.y), n1 = _curand_box_muller(r.z, r.w);
    const float eps[4] = {n0.x, n0.y, n1.x, n1.y};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) philox_elem<DM>(base + S * ii, a[ii], b[ii], tv[ii], eps[ii], sigma, mode, xt, ut, eps_out);
  }
  {   // last iteration: elements past the end of the tensor are skipped (their normals are drawn and dropped, as ATen does)
    const uint4 r = curand_Philox4x32_10(make_uint4((unsigned)c01, (unsigned)(c01 >> 32), idx, 0u), key);
    const float2 n0 = _curand_box_muller(r.x, r.y), n1 = _curand_box_muller(r.z, r.w);
    const float eps[4] = {n0.x, n0.y, n1.x, n1.y};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const unsigned e = base + S * ii;
      if (e < n) philox_elem<DM>(e, __ldcs(x0 + e), __ldcs(x1 + e), __ldg(t + row_of(e)), eps[ii], sigma, mode, xt, ut, eps_out);
    }
  }
}

static int philox_plan(long numel, int dev, int* grid, unsigned long long* inc) {
  int sms = 0, tpsm = 0;
  IDFF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDFF_CUDA(cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
  const unsigned long long n = (unsigned long long)numel;
  unsigned long long g = (n + 255ull) / 256ull;
  const unsigned long long cap = (unsigned long long)sms * (unsigned long long)(tpsm / 256);
  if (g > cap) g = cap;
  *grid = (int)g;
  *inc = ((n - 1ull) / (256ull * g * 4ull) + 1ull) * 4ull;
  return IDFF_OK;
}

}  // namespace idff

using namespace idff;

extern "C" int idff_philox_plan(int64_t numel, int32_t device, int32_t* h_grid, uint64_t* h_offset_increment) {
  IDFF_CHECK_ARG(h_grid && h_offset_increment, "idff_philox_plan: null pointer");
  IDFF_CHECK_ARG(numel >= 1 && numel < (1ll << 31), "idff_philox_plan: numel %lld outside [1, 2^31)", (long long)numel);
  unsigned long long inc = 0;
  int grid = 0;
  int rc = philox_plan((long)numel, device, &grid, &inc);
  *h_grid = grid;
  *h_offset_increment = (uint64_t)inc;
  return rc;
}

extern "C" int idff_path_sample_philox(const float* d_x0, const float* d_x1, const float* d_t, uint64_t seed,
                                       uint64_t philox_offset, float sigma, int32_t sigma_mode, float* d_xt, float* d_ut,
                                       float* d_eps_out, int64_t B, int64_t D, uint64_t* h_offset_increment, void* stream) {
  IDFF_CHECK_ARG(B >= 0 && D >= 1, "idff_path_sample_philox: bad shape B=%lld D=%lld", (long long)B, (long long)D);
  if (h_offset_increment) *h_offset_increment = 0;
  if (B == 0) return IDFF_OK;   // randn_like of an empty tensor draws nothing
  IDFF_CHECK_ARG(d_x0 && d_x1 && d_t && d_xt, "idff_path_sample_philox: null pointer");
  IDFF_CHECK_ARG(sigma_mode == 0 || sigma_mode == 1, "idff_path_sample_philox: sigma_mode %d", sigma_mode);
  IDFF_CHECK_ARG(philox_offset % 4 == 0, "idff_path_sample_philox: the Philox offset must be a multiple of 4 (torch's always is)");
  const long long numel = (long long)B * (long long)D;
  if (numel >= (1ll << 31)) {
    set_error("idff_path_sample_philox: %lld elements: ATen splits tensors beyond 32-bit indexing into several RNG launches; "
              "draw eps with torch and call idff_path_sample", numel);
    return IDFF_EUNSUPPORTED;
  }
  PtrDeviceGuard guard(d_xt);
  int grid = 0;
  unsigned long long inc = 0;
  int rc = philox_plan((long)numel, guard.dev, &grid, &inc);
  if (rc) return rc;
  if (h_offset_increment) *h_offset_increment = (uint64_t)inc;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned n = (unsigned)numel, Du = (unsigned)D;
  if (D == 1)
    k_path_sample_philox<1><<<grid, 256, 0, s>>>(d_x0, d_x1, d_t, seed, philox_offset, sigma, sigma_mode, d_xt, d_ut, d_eps_out, n, Du);
  else if (D == 2)
    k_path_sample_philox<2><<<grid, 256, 0, s>>>(d_x0, d_x1, d_t, seed, philox_offset, sigma, sigma_mode, d_xt, d_ut, d_eps_out, n, Du);
  else
    k_path_sample_philox<0><<<grid, 256, 0, s>>>(d_x0, d_x1, d_t, seed, philox_offset, sigma, sigma_mode, d_xt, d_ut, d_eps_out, n, Du);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}
