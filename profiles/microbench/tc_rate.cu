// pure tcgen05 tf32 MMA issue/execute rate for M=128, N in {64,128,256}, SS operands resident in smem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int N>
__global__ void __launch_bounds__(128) k(long long* out, int reps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 2 * N * 1024) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (tid == 0) {
    uint32_t parity = 0;
    const uint32_t a0 = s32(smem), b0 = s32(smem + 65536);
    for (int r = 0; r < reps + 1; ++r) {
      long long t0 = clock64();
      for (int t = 0; t < 16; ++t) {             // 16 (m-tile, k-block) tiles, 12 MMAs each, like one layer pass
        const uint32_t wa = a0 + (t & 1) * 32768;
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ahi = make_desc(wa + ks * 32), alo = make_desc(wa + 16384 + ks * 32);
          const uint64_t bhi = make_desc(b0 + (t & 7) * (N * 128) + ks * 32), blo = make_desc(b0 + N * 1024 + (t & 7) * (N * 128) + ks * 32);
          const uint32_t dm = tmem + (t >> 3) * 256, dc = dm + N > tmem + 512 - N ? tmem : dm + N;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(dm), "l"(ahi), "l"(bhi), "r"(idesc), "r"(1u) : "memory");
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(dc), "l"(alo), "l"(bhi), "r"(idesc), "r"(1u) : "memory");
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(dc), "l"(ahi), "l"(blo), "r"(idesc), "r"(1u) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(s32(&bar)), "r"(parity) : "memory");
      parity ^= 1;
      long long t1 = clock64();
      if (r == reps && blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N> void run(long long* d) {
  const int smem = 65536 + 2 * N * 1024 + 2048;   // 2 weight tiles (hi|lo) + B hi/lo (N x 128 floats each... N*512 B per copy-k-block x 8 / shared)
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    k<N><<<grid, 128, smem>>>(d, 3);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d grid=%3d: 192 MMAs (128xNx8) in %lld cycles = %.1f cyc/MMA (ideal %d) %s\n", N, grid, c, c / 192.0, N / 2, cudaGetErrorString(e));
  }
}
int main() { long long* d; cudaMalloc(&d, 8); run<64>(d); run<32>(d); run<48>(d); return 0; }
