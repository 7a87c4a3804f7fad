// Feasibility test of a hand-written tcgen05 tf32 MMA (SS operands, K-major, 128B swizzle), 3xTF32 split accuracy.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
constexpr int M = 128, N = 64, K = 256;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                 // version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t swz_off(int rows, int r, int k) {   // byte offset of element (r,k) in a [rows x K] operand
  const int kb = k >> 5, kk = k & 31, chunk = kk >> 2, w = kk & 3;
  return (uint32_t)(kb * rows * 128 + r * 128 + ((chunk ^ (r & 7)) << 4) + w * 4);
}
__global__ void __launch_bounds__(128) k(const float* __restrict__ Ahi, const float* __restrict__ Alo,
                                         const float* __restrict__ Bhi, const float* __restrict__ Blo,
                                         float* __restrict__ D, int nterms) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sAhi = smem;                       // 128 x 256 x 4 = 128 KB   (only hi resident; lo streamed per test)
  unsigned char* sB = smem + 131072;                // 64 x 256 x 4 = 64 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t parity = 0;
  for (int term = 0; term < nterms; ++term) {
    // term 0: Ahi x Bhi, term 1: Alo x Bhi, term 2: Ahi x Blo
    const float* Ag = term == 1 ? Alo : Ahi;
    const float* Bg = term == 2 ? Blo : Bhi;
    __syncthreads();
    for (int i = tid; i < M * K; i += 128) { const int r = i / K, kx = i % K; *(float*)(sAhi + swz_off(M, r, kx)) = Ag[i]; }
    for (int i = tid; i < N * K; i += 128) { const int r = i / K, kx = i % K; *(float*)(sB + swz_off(N, r, kx)) = Bg[i]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int kb = 0; kb < K / 32; ++kb)
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da = make_desc(s32(sAhi) + kb * (M * 128) + ks * 32);
          const uint64_t db = make_desc(s32(sB) + kb * (N * 128) + ks * 32);
          const uint32_t acc = (term | kb | ks) != 0;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    // wait for the MMAs (they read smem) before the next term overwrites the operands
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(s32(&bar)), "r"(parity) : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  uint32_t v[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int c = 0; c < 64; ++c) D[(warp * 32 + lane) * N + c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}
static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
int main() {
  std::vector<float> A(M * K), B(N * K), Ahi(M * K), Alo(M * K), Bhi(N * K), Blo(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
  for (int i = 0; i < M * K; ++i) { Ahi[i] = trunc_tf32(A[i]); Alo[i] = A[i] - Ahi[i]; }
  for (int i = 0; i < N * K; ++i) { Bhi[i] = trunc_tf32(B[i]); Blo[i] = B[i] - Bhi[i]; }
  float *dA, *dAl, *dB, *dBl, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dAl, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dBl, N * K * 4); cudaMalloc(&dD, M * N * 4);
  const int smem = 131072 + 65536 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  auto run = [&](const std::vector<float>& a, const std::vector<float>& al, const std::vector<float>& b, const std::vector<float>& bl, int nterms) {
    cudaMemcpy(dA, a.data(), M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dAl, al.data(), M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, b.data(), N * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBl, bl.data(), N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    k<<<1, 128, smem>>>(dA, dAl, dB, dBl, dD, nterms);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  };
  auto err_vs = [&](const std::vector<float>& a, const std::vector<float>& b, const char* what) {
    double maxabs = 0, maxrel = 0, scale = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double s = 0, sa = 0; for (int kx = 0; kx < K; ++kx) { s += (double)a[m * K + kx] * b[n * K + kx]; sa += fabs((double)a[m * K + kx] * b[n * K + kx]); }
      maxabs = fmax(maxabs, fabs(D[m * N + n] - s)); scale = fmax(scale, fabs(s)); maxrel = fmax(maxrel, fabs(D[m * N + n] - s) / sa);
    }
    printf("%-40s max|err| %.3e  max|err|/sum|terms| %.3e  (max|D| %.3f)\n", what, maxabs, maxrel, scale);
  };
  run(Ahi, Alo, Bhi, Blo, 1); err_vs(Ahi, Bhi, "tf32-exact inputs, 1 term vs fp64");
  run(A, Alo, B, Blo, 1);
  { std::vector<float> ar(M * K), br(N * K); for (int i = 0; i < M * K; ++i) ar[i] = round_tf32(A[i]); for (int i = 0; i < N * K; ++i) br[i] = round_tf32(B[i]);
    err_vs(Ahi, Bhi, "raw fp32 inputs vs TRUNCATED model"); err_vs(ar, br, "raw fp32 inputs vs ROUNDED model"); }
  run(Ahi, Alo, Bhi, Blo, 3); err_vs(A, B, "3xTF32 trunc split vs fp64");
  for (int i = 0; i < M * K; ++i) { Ahi[i] = round_tf32(A[i]); Alo[i] = round_tf32(A[i] - Ahi[i]); }
  for (int i = 0; i < N * K; ++i) { Bhi[i] = round_tf32(B[i]); Blo[i] = round_tf32(B[i] - Bhi[i]); }
  run(Ahi, Alo, Bhi, Blo, 3); err_vs(A, B, "3xTF32 round-to-nearest split vs fp64");
  for (int i = 0; i < M * K; ++i) { Alo[i] = A[i] - Ahi[i]; }
  for (int i = 0; i < N * K; ++i) { Blo[i] = B[i] - Bhi[i]; }
  run(Ahi, Alo, Bhi, Blo, 3); err_vs(A, B, "3xTF32 rn hi, exact lo (HW truncs lo)");
  // fp32 FMA reference error for comparison
  { double maxrel = 0; for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; double sd = 0, sa = 0; for (int kx = 0; kx < K; ++kx) { s = fmaf(A[m * K + kx], B[n * K + kx], s); sd += (double)A[m * K + kx] * B[n * K + kx]; sa += fabs((double)A[m * K + kx] * B[n * K + kx]); } maxrel = fmax(maxrel, fabs(s - sd) / sa); }
    printf("%-40s max|err|/sum|terms| %.3e\n", "fp32 FMA chain vs fp64", maxrel); }
  return 0;
}
