// Which (lane, column) of tensor memory lands in which thread / register for tcgen05.ld.16x128b.x4 (and .16x256b.x2):
// fill TMEM through the 32x32b shape (thread = lane, register i = column i) with value lane*100 + column, read back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k(int* out) {
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(32)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_s + ((uint32_t)(warp * 32) << 16);
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = (warp * 32 + lane) * 100 + c;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(tmem) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t a[8], b[8];
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "r"(tmem));
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(tmem + (16u << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) { out[tid * 16 + i] = a[i]; out[tid * 16 + 8 + i] = b[i]; }
  // write through 16x128b, read through 32x32b (checks the st shape is the mirror image)
  for (int i = 0; i < 8; ++i) { a[i] = 100000 + tid * 16 + i; b[i] = 100000 + tid * 16 + 8 + i; }
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(tmem + 16) : "memory");
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(tmem + (16u << 16) + 16) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(tmem + 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 16; ++c) out[128 * 16 + tid * 16 + c] = v[c];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(32));
}
int main() {
  int* d; cudaMalloc(&d, 2 * 128 * 16 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  static int h[2 * 128 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  // expectation: register r = 2*rep + hh of load A (lane offset 0) holds lane (T/4) + 8*hh, column 4*rep + T%4; load B the same +16 lanes
  int bad = 0;
  for (int t = 0; t < 128; ++t) for (int i = 0; i < 16; ++i) {
    const int w = t / 32, T = t % 32, ld = i / 8, r = i % 8, rep = r / 2, hh = r % 2;
    const int lane = w * 32 + ld * 16 + T / 4 + 8 * hh, col = 4 * rep + T % 4;
    if (h[t * 16 + i] != lane * 100 + col) { if (bad < 8) printf("ld mismatch thread %d reg %d: got %d expected %d\n", t, i, h[t * 16 + i], lane * 100 + col); ++bad; }
  }
  printf("16x128b.x4 load layout as expected: %s (%d mismatches)\n", bad ? "NO" : "yes", bad);
  for (int T = 0; T < 8; ++T) { printf("thread %d:", T); for (int i = 0; i < 16; ++i) printf(" %d", h[T * 16 + i]); printf("\n"); }
  int bad2 = 0;
  for (int t = 0; t < 128; ++t) for (int c = 0; c < 16; ++c) {
    // lane t (32x32b), column c  <-  written by thread T with T%4 == c%4, register 2*(c/4) + hh, load ld
    const int w = t / 32, L = t % 32, ld = L / 16, hh = (L % 16) / 8, T = (L % 8) * 4 + c % 4, r = 2 * (c / 4) + hh;
    const int exp = 100000 + (w * 32 + T) * 16 + ld * 8 + r;
    if (h[128 * 16 + t * 16 + c] != exp) { if (bad2 < 8) printf("st mismatch lane %d col %d: got %d expected %d\n", t, c, h[128 * 16 + t * 16 + c], exp); ++bad2; }
  }
  printf("16x128b.x4 store layout mirrors the load: %s (%d mismatches)\n", bad2 ? "NO" : "yes", bad2);
  return 0;
}
