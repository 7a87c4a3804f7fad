// Device helpers shared by the 3xFP16 tcgen05 samplers (idff_sampler_tc16.cu, idff_sampler_tc16o3.cu): mbarrier / TMA /
// tcgen05 wrappers, shared-memory matrix descriptors, the fp16 hi / lo' split and the SELU jets.  sm_100a only.
#pragma once
#include <cuda_fp16.h>

#include "idff_common.cuh"

namespace idff {
namespace tcg {

constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;   // v = hi + 2^-11 * lo'

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  while (!mbar_try_wait(b, parity)) {
  }
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* b, uint32_t parity) {
  while (!mbar_try_wait(b, parity)) __nanosleep(128);
}
__device__ __forceinline__ void tma_load_1d_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)),
      "l"(src), "r"(bytes), "r"(s32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ebar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // epilogue warps only
__device__ __forceinline__ void ebar576() { asm volatile("bar.sync 1, 576;" ::: "memory"); }   // epilogue + integrator warps (k_tc16)

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// K-major SWIZZLE_64B (64-byte rows, 8-row groups 512 B apart): the activation operand
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per instruction on a 64-bit register pair)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t bc2(float v) { return pk2(v, v); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// v0, v1 -> packed fp16 pairs: h2 = {lo half: hi(v0), hi half: hi(v1)}, l2 likewise for the scaled remainders.
// The remainder (v - hi) * 2^11 of both values is one FADD2 + one FMUL2 (both exact or correctly rounded like the scalar ops).
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& h2, uint32_t& l2) {
  float f0, f1, d0, d1;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(v1), "f"(v0));
  asm("{\n.reg .b16 a, b;\nmov.b32 {a, b}, %2;\ncvt.f32.f16 %0, a;\ncvt.f32.f16 %1, b;\n}\n" : "=f"(f0), "=f"(f1) : "r"(h2));
  upk2(mul2(sub2(pk2(v0, v1), pk2(f0, f1)), pk2(LO_SCALE, LO_SCALE)), d0, d1);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
}
// the same on a packed pair
__device__ __forceinline__ void split2p(uint64_t v, uint32_t& h2, uint32_t& l2) {
  float v0, v1, f0, f1, d0, d1;
  upk2(v, v0, v1);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(v1), "f"(v0));
  asm("{\n.reg .b16 a, b;\nmov.b32 {a, b}, %2;\ncvt.f32.f16 %0, a;\ncvt.f32.f16 %1, b;\n}\n" : "=f"(f0), "=f"(f1) : "r"(h2));
  upk2(mul2(sub2(v, pk2(f0, f1)), bc2(LO_SCALE)), d0, d1);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(d1), "f"(d0));
}
// tcgen05.ld/st.16x128b.x4: 16 TMEM lanes x 16 columns; thread T: register 2r + h <-> lane T/4 + 8h, column 4r + T%4
// [layout verified on the part: profiles/microbench/tmem_shapes.cu]
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(taddr)
               : "memory");
}
// tcgen05.ld/st.16x128b.x2: 16 TMEM lanes x 8 columns; thread T: register 2r + h <-> lane T/4 + 8h, column 4r + T%4, r < 2
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%4], {%0,%1,%2,%3};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(taddr)
               : "memory");
}
template <int OFF>
__device__ __forceinline__ void st64(uint32_t base, uint32_t w0, uint32_t w1) {
  asm volatile("st.shared.v2.b32 [%0+%3], {%1, %2};" ::"r"(base), "r"(w0), "r"(w1), "n"(OFF) : "memory");
}

// SELU value a, first derivative s1 and E = every higher derivative at z (ATen: z <= 0 -> (exp(z) - 1) * LA, derivative
// LA * exp(z); z > 0 -> lambda * z, lambda, 0).  exp through ex2.approx on z * log2(e): the same MUFU.EX2 CUDA's expf is
// built on (2 ulp), without its range-reduction scaffolding; the extra input rounding is |z| * 2^-24 relative.
__device__ __forceinline__ void selu_jet(float z, float& a, float& s1, float& E) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(z, 0.0f) * 1.4426950408889634f));
  const bool pos = !(z <= 0.0f);   // keeps NaN on the linear branch
  const float t = kSeluLA * e;
  a = pos ? z * kSeluLambda : (e - 1.0f) * kSeluLA;
  s1 = pos ? kSeluLambda : t;
  E = pos ? 0.0f : t;
}

// same, returning the derivative code kept in the stash: LA * exp(z) (= s1 = E) on the exp branch, -1 on the linear one
__device__ __forceinline__ void selu_jet_code(float z, float& a, float& s1, float& code) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(z, 0.0f) * 1.4426950408889634f));
  const bool pos = !(z <= 0.0f);
  const float t = kSeluLA * e;
  a = pos ? z * kSeluLambda : (e - 1.0f) * kSeluLA;
  s1 = pos ? kSeluLambda : t;
  code = pos ? -1.0f : t;
}

// SELU jets of a packed pair of pre-activations (the scalar formulas above, two lanes per FMUL2 / FADD2): a = s(z), s1 = s'(z),
// E = every higher derivative.  MUFU.EX2, the min and the selects stay scalar.
__device__ __forceinline__ void selu_jet2(uint64_t z, uint64_t& a, uint64_t& s1, uint64_t& E) {
  float z0, z1, g0, g1, e0, e1, t0, t1, n0, n1, p0, p1;
  upk2(z, z0, z1);
  upk2(mul2(pk2(fminf(z0, 0.0f), fminf(z1, 0.0f)), bc2(1.4426950408889634f)), g0, g1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const uint64_t e = pk2(e0, e1);
  upk2(mul2(e, bc2(kSeluLA)), t0, t1);
  upk2(mul2(add2(e, bc2(-1.0f)), bc2(kSeluLA)), n0, n1);
  upk2(mul2(z, bc2(kSeluLambda)), p0, p1);
  const bool q0 = !(z0 <= 0.0f), q1 = !(z1 <= 0.0f);   // keeps NaN on the linear branch
  a = pk2(q0 ? p0 : n0, q1 ? p1 : n1);
  s1 = pk2(q0 ? kSeluLambda : t0, q1 ? kSeluLambda : t1);
  E = pk2(q0 ? 0.0f : t0, q1 ? 0.0f : t1);
}
// same, with the derivative code kept in the stash: LA * exp(z) (= s1 = E) on the exp branch, -1 on the linear one
__device__ __forceinline__ void selu_jet_code2(uint64_t z, uint64_t& a, uint64_t& s1, float& code0, float& code1) {
  float z0, z1, g0, g1, e0, e1, t0, t1, n0, n1, p0, p1;
  upk2(z, z0, z1);
  upk2(mul2(pk2(fminf(z0, 0.0f), fminf(z1, 0.0f)), bc2(1.4426950408889634f)), g0, g1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const uint64_t e = pk2(e0, e1);
  upk2(mul2(e, bc2(kSeluLA)), t0, t1);
  upk2(mul2(add2(e, bc2(-1.0f)), bc2(kSeluLA)), n0, n1);
  upk2(mul2(z, bc2(kSeluLambda)), p0, p1);
  const bool q0 = !(z0 <= 0.0f), q1 = !(z1 <= 0.0f);
  a = pk2(q0 ? p0 : n0, q1 ? p1 : n1);
  s1 = pk2(q0 ? kSeluLambda : t0, q1 ? kSeluLambda : t1);
  code0 = q0 ? -1.0f : t0;
  code1 = q1 ? -1.0f : t1;
}

}  // namespace tcg
}  // namespace idff
