// Tensor-core IDFF sampler for sm_100a (IDFF_IMPL_TENSOR): hidden width 256, state dimension 2, up to two jets.
//
// Every w x w layer of the evaluation (forward jets and reverse sweep) runs on tcgen05.mma kind::tf32 with the 3xTF32
// split  D = W_hi*X_hi + (W_lo*X_hi + W_hi*X_lo)  and fp32 accumulation in tensor memory, which keeps fp32-grade
// accuracy (hi = rn_tf32(v), lo = v - hi; the hardware truncates raw fp32 operands to tf32 [measured]).
//
// Formulation ("units on M"): D[unit][col] = sum_k W[unit][k] * X[col][k], col = jet*32 + particle.
//   * A operand = weights, 128-unit M-tiles, streamed from L2 as pre-swizzled 32 KB (hi|lo) tile images by a TMA
//     producer warp (cp.async.bulk + mbarrier) -- 2 m-tiles x 8 k-blocks per layer pass;
//   * B operand = the activations of the CTA's 32 particles x 2 jets (N = 64), K-major SWIZZLE_128B, hi and lo copies
//     resident in shared memory (128 KB), rewritten in place by the epilogue of every layer;
//   * accumulators in TMEM: per m-tile a main (hi*hi) and a correction (lo*hi + hi*lo) block of 64 columns each, so the
//     2^-11-sized terms do not disturb the rounding of the main sum; the layer-2 derivative stash of the reverse sweep
//     lives in spare TMEM columns.
// A TMEM lane is one hidden unit, so the thread that reads lane n with tcgen05.ld owns every jet of every particle of
// unit n: the SELU jet / adjoint algebra is thread-local, per-unit constants are scalars.  Reductions over units
// (x1hat = W3 a3, grad_x = W0^T P1) are a shuffle reduce-scatter.  Warp roles: 4 epilogue warps, 1 TMA producer, 1 MMA
// issuer (single elected thread); mbarriers between them, named barriers among the epilogue warps.
#include <cstdlib>

#include "idff_common.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);
// After the first launch on a handle: synchronise once and fail loudly if the kernel found its dynamic shared memory
// misaligned for SWIZZLE_128B operands (it then returned without computing).
__device__ int g_tc_misaligned = 0;
static int verify_tc_launch(idff_weights_t* w, void* stream, int bit) {
  if (w->tc_verified & bit) return IDFF_OK;
  int flag = 0;
  IDFF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  IDFF_CUDA(cudaMemcpyFromSymbol(&flag, g_tc_misaligned, sizeof(int)));
  if (flag) {
    set_error("tcgen05 kernel: dynamic shared memory is not 1024-byte aligned on this driver; use impl = IDFF_IMPL_FUSED");
    return IDFF_EUNSUPPORTED;
  }
  w->tc_verified |= bit;
  return IDFF_OK;
}

namespace tc {

constexpr int W = 256;
constexpr int NP = 32;                   // particles per CTA
constexpr int NCOL = 64;                 // MMA N: 2 jets x 32 particles
constexpr int NSTAGE = 3;                // weight ring depth (bytes in flight hide the ~1000-cycle L2->SM latency)
constexpr int TILE_BYTES = 32768;        // hi (16 KB) + lo (16 KB) image of one [128 x 32] weight tile
constexpr int KB_BYTES = 2 * NCOL * 128;  // B operand of one k-block: rows 0..63 = hi copy, rows 64..127 = lo copy (16 KB)
constexpr int ACT_BYTES = 8 * KB_BYTES;   // whole B operand: 8 k-blocks
constexpr int NE = 512;                  // epilogue threads: warp w -> lane quarter w%4, m-tile (w/4)%2, particle half w/8
constexpr int NEW = NE / 32;             // epilogue warps
constexpr int NTHREADS = NE + 64;        // + producer warp 16 + MMA warp 17
constexpr int TMEM_COLS = 512;
// accumulator sets (shared by the two m-tiles, which run back to back): set s = k-block parity, [main 64 | corr 64];
// two sets halve the number of truncating accumulation steps per accumulator.  Stash: 64 columns per m-tile.
constexpr int COL_SET = 128, COL_CORR = 64, COL_STASH = 256;

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
  while (!mbar_try_wait(b, parity)) __nanosleep(256);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
// each CTA of a pair loads one half (hi or lo) of every weight tile and multicasts it to both CTAs: halves the L2 reads
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

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}

#define IDFF_R8(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define IDFF_I8(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : IDFF_R8(v, 0), IDFF_R8(v, 8), IDFF_R8(v, 16), IDFF_R8(v, 24)
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" ::IDFF_I8(v, 0),
      IDFF_I8(v, 8), IDFF_I8(v, 16), IDFF_I8(v, 24), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : IDFF_R8(v, 0), IDFF_R8(v, 8)
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::IDFF_I8(v, 0),
               IDFF_I8(v, 8), "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// phase timestamps (clock64) of CTA 0, first tile, evaluation 1 (an interior one): diagnostic only
__device__ long long g_tc_stamps[32];

struct TParams {
  NetView net;
  int variant, D, tidx, reverse, mode;   // mode 0 eval, 1 rk4
  long N;
  const float* x_in;
  const float* x0_cfm;
  float* out;
  float* traj;
};

struct Smem {   // byte offsets from the dynamic base, which must be 1024-byte aligned (checked at kernel entry)
  static constexpr int act = 0;
  static constexpr int ring = ACT_BYTES;
  static constexpr int misc = ring + NSTAGE * TILE_BYTES;   // floats: red [16 warps][32], xs [64]
  static constexpr int f_red = 0, f_xs = NEW * 32, f_end = NEW * 32 + 64;
  static constexpr int bars = misc + f_end * 4;
  static constexpr int total = bars + 128;
};
static_assert(Smem::total <= 232448, "shared memory budget");

// element (col, k) of the B operand: byte offset inside one copy
__device__ __forceinline__ uint32_t act_off(int col, int k) {
  return (uint32_t)((k >> 5) * (NCOL * 128) + col * 128 + ((((k >> 2) & 7) ^ (col & 7)) << 4) + ((k & 3) << 2));
}

// reduce-scatter of 64 per-thread partial sums over the 32 lanes of a warp: lane L ends with the totals of v[2L], v[2L+1]
__device__ __forceinline__ void warp_reduce_scatter64(float (&v)[64], int lane) {
#pragma unroll
  for (int off = 16, cnt = 32; off >= 1; off >>= 1, cnt >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const float send = up ? v[i] : v[i + cnt];
      const float keep = up ? v[i + cnt] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// same for 32 partial sums: lane L ends with the total of v[L]
__device__ __forceinline__ void warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, cnt = 16; off >= 1; off >>= 1, cnt >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const float send = up ? v[i] : v[i + cnt];
      const float keep = up ? v[i + cnt] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_tc(const __grid_constant__ TParams P, const __grid_constant__ Rk4Table rk, const __grid_constant__ Stage one) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((s32(smem) & 1023u) != 0u) {   // SWIZZLE_128B operands need a 1024-byte aligned base: report and do nothing
    if (threadIdx.x == 0) g_tc_misaligned = 1;
    return;
  }   // SWIZZLE_128B operands need a 1024-byte aligned base (host checks the result)
  float* misc = reinterpret_cast<float*>(smem + Smem::misc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;                 // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;       // [NSTAGE]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * NSTAGE;   // [2] MMA -> epilogue: the accumulators of m-tile mt are complete
  // epilogue -> MMA.  b_ready_lo: k-blocks 0..3 of the next layer pass are written (m-tile-0 threads) AND the m-tile-1
  // threads have read their accumulators out; b_ready_hi: k-blocks 4..7 are written (m-tile-1 threads).  The next pass
  // starts on k-blocks 0..3 while the m-tile-1 epilogue is still computing.
  uint64_t* b_ready_lo = acc_full + 2;
  uint64_t* b_ready_hi = acc_full + 3;
  uint64_t* tmem_free = acc_full + 4;       // m-tile-0 epilogue threads -> MMA: accumulators read, m-tile 1 may overwrite
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);   // both CTAs of the pair must have consumed a slot before either half is overwritten
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(b_ready_lo, NE);
    mbar_init(b_ready_hi, NE / 2);
    mbar_init(tmem_free, NE / 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NEW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // jet-1 columns are read by the MMAs of boundary evaluations too: start from finite values
  for (int i = tid; i < ACT_BYTES / 4; i += NTHREADS) reinterpret_cast<float*>(smem)[i] = 0.0f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barrier inits visible to the peer before any remote complete_tx / commit arrives
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t crank = cluster_ctarank();
  // both CTAs of a pair walk the same number of tiles (the ring couples them); surplus tiles are all-padding
  const long n_tiles = (P.N + NP - 1) / NP;
  const long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;

  const int n_eval = P.mode == 0 ? 1 : 4 * rk.n_iv;
  auto stage_of = [&](int e) -> const Stage& { return P.mode == 0 ? one : rk.st[e]; };
  auto heavy_of = [&](const Stage& st) { return st.kind == 2 && P.reverse; };

  if (warp == NEW) {
    // ===== TMA producer: weight tile images of every layer pass, in consumption order
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long it = 0; it < n_iter; ++it)
        for (int e = 0; e < n_eval; ++e) {
          const int ng = heavy_of(stage_of(e)) ? 4 : 2;
          for (int g = 0; g < ng; ++g)
            for (int t = 0; t < 16; ++t) {   // (m-tile, k-block) pairs of gemm g are contiguous
              mbar_wait_sleep(&empty[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
              tma_load_1d_mc(smem + Smem::ring + stage * TILE_BYTES + crank * (TILE_BYTES / 2),
                             reinterpret_cast<const unsigned char*>(P.net.tc_tiles) + ((size_t)g * 16 + t) * TILE_BYTES +
                                 crank * (TILE_BYTES / 2),
                             TILE_BYTES / 2, &full[stage], (uint16_t)3);
              if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
            }
        }
    }
  } else if (warp == NEW + 1) {
    // ===== MMA issuer: the whole warp walks the schedule in uniform control flow (descriptors and TMEM addresses in
    // uniform registers), one elected lane issues -- see idff_sampler_tc16.cu
    {
      const uint32_t idesc_w = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bop = s32(smem + Smem::act), ring0 = s32(smem + Smem::ring);
      uint32_t el;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(el));
      const bool leader = el != 0;
      int stage = 0;
      uint32_t phase = 0, bphase = 0, hphase = 0, fphase = 0;
      for (long it = 0; it < n_iter; ++it)
        for (int e = 0; e < n_eval; ++e) {
          const int ng = heavy_of(stage_of(e)) ? 4 : 2;
          for (int g = 0; g < ng; ++g) {
            mbar_wait(b_ready_lo, bphase);
            bphase ^= 1u;
            tc_fence_after();
            const bool mstamp = blockIdx.x == 0 && it == 0 && e == 1 && lane == 0;
            if (mstamp) g_tc_stamps[16 + 2 * g] = clock64();
            for (int mt = 0; mt < 2; ++mt) {
              if (mt == 1) {   // the accumulator sets are reused: m-tile 0 must have been read out
                mbar_wait(tmem_free, fphase);
                fphase ^= 1u;
                tc_fence_after();
              }
              for (int kb = 0; kb < 8; ++kb) {
                if (mt == 0 && kb == 4) {   // the upper half of K comes from the m-tile-1 epilogue threads
                  mbar_wait(b_ready_hi, hphase);
                  hphase ^= 1u;
                  tc_fence_after();
                }
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t a_hi0 = make_desc(ring0 + (uint32_t)(stage * TILE_BYTES));
                const uint64_t a_lo0 = make_desc(ring0 + (uint32_t)(stage * TILE_BYTES + 16384));
                const uint64_t b_0 = make_desc(bop + (uint32_t)(kb * KB_BYTES));   // 128 rows: hi copy then lo copy
                const uint32_t dset = tmem + (kb & 1) * COL_SET;
                if (leader) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    // [main | corr] += W_hi * [X_hi ; X_lo]   (one N = 128 MMA reads W_hi once), then corr += W_lo * X_hi
                    mma_tf32(dset, a_hi0 + (uint64_t)(2 * ks), b_0 + (uint64_t)(2 * ks), idesc_w, ((kb >> 1) | ks) != 0);
                    mma_tf32(dset + COL_CORR, a_lo0 + (uint64_t)(2 * ks), b_0 + (uint64_t)(2 * ks), idesc_n, 1u);
                  }
                  mma_commit_mc(&empty[stage], (uint16_t)3);   // frees the slot in both CTAs once these MMAs have read it
                }
                __syncwarp();
                if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
              }
              if (leader) mma_commit(&acc_full[mt]);
              __syncwarp();
            }
            if (mstamp) g_tc_stamps[17 + 2 * g] = clock64();
          }
        }
    }
  } else {
    // ===== epilogue warps: thread <-> (hidden unit n, half h of the CTA's particles)
    const int mt = (warp >> 2) & 1, h = warp >> 3;
    const int n = mt * 128 + (warp & 3) * 32 + lane;
    constexpr int HP = NP / 2;             // particles per thread
    const int p0 = h * HP;
    float* red = misc + Smem::f_red;       // [16 warps][32]
    float* sxs = misc + Smem::f_xs;        // stage input of the 32 particles, [p][2]
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + p0;
    const uint32_t stash_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + COL_STASH + mt * 64 + p0;
    uint32_t aphase0 = 0, aphase1 = 0;
    const float third = (float)(1.0 / 3.0);
    const float wx0 = __ldg(P.net.wt0 + 0 * W + n), wx1 = __ldg(P.net.wt0 + 1 * W + n), wtt = __ldg(P.net.wt0 + 2 * W + n),
                bb0 = __ldg(P.net.b0 + (size_t)P.tidx * W + n), bb1 = __ldg(P.net.b1 + n), bb2 = __ldg(P.net.b2 + n),
                uu3 = __ldg(P.net.u3 + n), zz1 = __ldg(P.net.zd1 + n), w30 = __ldg(P.net.w3 + 0 * W + n),
                w31 = __ldg(P.net.w3 + 1 * W + n);
    // integrator state of particle `lane`, held by warp 0 (y, k1, k2, k3, x0 of the CFM field: 2 components each)
    float sy[2] = {0.f, 0.f}, sk1[2] = {0.f, 0.f}, sk2[2] = {0.f, 0.f}, sk3[2] = {0.f, 0.f}, sx0[2] = {0.f, 0.f};
    float pr_tot[2] = {0.f, 0.f};
    // this unit's k position inside the B operand: k-block, 16-byte chunk, word
    unsigned char* ahi = smem + Smem::act + (n >> 5) * KB_BYTES + ((n & 3) << 2);
    unsigned char* alo = ahi + NCOL * 128;
    const int kchunk = (n >> 2) & 7;
    auto put = [&](int col, float v) {
      const uint32_t o = (uint32_t)(col * 128 + ((kchunk ^ (col & 7)) << 4));
      const float hv = rn_tf32(v);
      *reinterpret_cast<float*>(ahi + o) = hv;
      *reinterpret_cast<float*>(alo + o) = v - hv;
    };
    // after writing the B operand: make it visible to the tensor core and hand over to the MMA thread
    auto publish = [&]() {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(mt == 0 ? b_ready_lo : b_ready_hi);
    };
    // z = sum of the two accumulator sets (main + corr each) of this thread's unit: jet j, particles p0 .. p0+15
    auto load_acc = [&](int j, float (&z)[HP]) {
      uint32_t m0[HP], c0[HP], m1[HP], c1[HP];
      tmem_ld16(lane_base + j * NP, m0);
      tmem_ld16(lane_base + COL_CORR + j * NP, c0);
      tmem_ld16(lane_base + COL_SET + j * NP, m1);
      tmem_ld16(lane_base + COL_SET + COL_CORR + j * NP, c1);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < HP; ++i)
        z[i] = (__uint_as_float(m0[i]) + __uint_as_float(m1[i])) + (__uint_as_float(c0[i]) + __uint_as_float(c1[i]));
    };
    // accumulators of this thread's m-tile complete
    auto wait_acc = [&]() {
      if (mt == 0) { mbar_wait_sleep(&acc_full[0], aphase0); aphase0 ^= 1u; }
      else { mbar_wait(&acc_full[1], aphase1); aphase1 ^= 1u; }
      tc_fence_after();
    };
    // accumulators are in registers: m-tile 0 -> m-tile 1 may overwrite them; m-tile 1 -> the next layer pass may
    auto release_acc = [&]() {
      tc_fence_before();
      mbar_arrive(mt == 0 ? tmem_free : b_ready_lo);
    };
    // the B operand may be rewritten once every MMA of the layer pass has read it (= m-tile 1 complete)
    auto wait_operand_free = [&]() {
      if (mt == 0) { mbar_wait(&acc_full[1], aphase1); aphase1 ^= 1u; }   // on the critical path: no back-off
    };

    if (mt == 1) mbar_arrive(b_ready_lo);   // nothing to read out before the very first layer pass
    for (long it = 0; it < n_iter; ++it) {
      const long base = (blockIdx.x + it * gridDim.x) * NP;   // may lie beyond N: an all-padding tile
      if (warp == 0) {
        const long row = base + lane;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const float v = row < P.N ? P.x_in[row * 2 + d] : 0.0f;
          sy[d] = v;
          sxs[2 * lane + d] = v;
          if (P.variant == IDFF_FIELD_CFM && P.x0_cfm) sx0[d] = row < P.N ? P.x0_cfm[row * 2 + d] : 0.0f;
        }
      }
      ebar();
#pragma unroll 1
      for (int ev = 0; ev < n_eval; ++ev) {
        const Stage& st = stage_of(ev);
        const bool heavy = heavy_of(st);
        const bool stamp = blockIdx.x == 0 && it == 0 && ev == 1 && tid == 0;
        if (stamp) g_tc_stamps[0] = clock64();
        // ---- layer 1 (K = 3, computed directly): a1 and, for interior evaluations, a1' = s'(z1) * (W0 d)
        {
          const float zb = fmaf(wtt, st.t, bb0);
#pragma unroll
          for (int i = 0; i < HP; ++i) {
            const int p = p0 + i;
            const float z = fmaf(wx1, sxs[2 * p + 1], fmaf(wx0, sxs[2 * p], zb));
            float code, s1, E;
            const float a = selu_fwd(z, code);
            selu_decode(code, s1, E);
            put(p, a);
            if (heavy) put(NP + p, s1 * zz1);
          }
        }
        publish();
        if (stamp) g_tc_stamps[1] = clock64();
        float part[32];
        // ---- gemm 0: layer 2 forward
        wait_acc();
        if (stamp) g_tc_stamps[2] = clock64();
        {
          float z[HP], zd[HP], o0[HP], o1[HP];
          load_acc(0, z);
          if (heavy) load_acc(1, zd);
          release_acc();
          uint32_t st_code[HP];
#pragma unroll
          for (int i = 0; i < HP; ++i) {
            float code, s1, E;
            o0[i] = selu_fwd(z[i] + bb1, code);
            selu_decode(code, s1, E);
            st_code[i] = __float_as_uint(code);
            o1[i] = heavy ? s1 * zd[i] : 0.0f;
          }
          if (heavy) {
            uint32_t st_zd[HP];
#pragma unroll
            for (int i = 0; i < HP; ++i) st_zd[i] = __float_as_uint(zd[i]);
            tmem_st16(stash_base, st_code);
            tmem_st16(stash_base + NP, st_zd);
            tmem_wait_st();
          }
          wait_operand_free();
#pragma unroll
          for (int i = 0; i < HP; ++i) {
            put(p0 + i, o0[i]);
            if (heavy) put(NP + p0 + i, o1[i]);
          }
        }
        publish();
        if (stamp) g_tc_stamps[3] = clock64();
        // ---- gemm 1: layer 3 forward; x1hat partials; adjoint seeds
        wait_acc();
        if (stamp) g_tc_stamps[4] = clock64();
        {
          float z[HP], zd[HP], o0[HP], o1[HP];
          load_acc(0, z);
          if (heavy) load_acc(1, zd);
          release_acc();
          const float A0 = st.c[0] * uu3, A1 = st.c[1] * uu3;
#pragma unroll
          for (int i = 0; i < HP; ++i) {
            float code, s1, E;
            const float a = selu_fwd(z[i] + bb2, code);
            selu_decode(code, s1, E);
            part[2 * i] = a * w30;
            part[2 * i + 1] = a * w31;
            o0[i] = heavy ? fmaf(A1 * E, zd[i], A0 * s1) : 0.0f;   // adjoint of z3
            o1[i] = A1 * s1;                                        // adjoint of z3'
          }
          wait_operand_free();
          if (heavy) {
#pragma unroll
            for (int i = 0; i < HP; ++i) {
              put(p0 + i, o0[i]);
              put(NP + p0 + i, o1[i]);
            }
          }
        }
        if (heavy) publish();
        if (stamp) g_tc_stamps[5] = clock64();
        warp_reduce_scatter32(part, lane);
        red[warp * 32 + lane] = part[0];
        ebar();
        if (warp == 0) {   // x1hat of particle `lane`: fixed summation order over the 8 warps of its half
          const int hh = lane >> 4, q = (lane & 15) * 2;
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            float acc = __ldg(P.net.b3 + d);
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) acc += red[(hh * 8 + w8) * 32 + q + d];
            pr_tot[d] = acc;
          }
        }
        if (stamp) g_tc_stamps[6] = clock64();
        float gg[2] = {0.f, 0.f};
        if (heavy) {
          // ---- gemm 2: reverse, layer 3 -> 2
          wait_acc();
          if (stamp) g_tc_stamps[7] = clock64();
          {
            float a0[HP], a1[HP], o0[HP], o1[HP];
            load_acc(0, a0);
            load_acc(1, a1);
            release_acc();
            uint32_t cd[HP], zdu[HP];
            tmem_ld16(stash_base, cd);
            tmem_ld16(stash_base + NP, zdu);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < HP; ++i) {
              float s1, E;
              selu_decode(__uint_as_float(cd[i]), s1, E);
              o0[i] = fmaf(a1[i] * E, __uint_as_float(zdu[i]), a0[i] * s1);
              o1[i] = a1[i] * s1;
            }
            wait_operand_free();
#pragma unroll
            for (int i = 0; i < HP; ++i) {
              put(p0 + i, o0[i]);
              put(NP + p0 + i, o1[i]);
            }
          }
          publish();
          if (stamp) g_tc_stamps[8] = clock64();
          // ---- gemm 3: reverse, layer 2 -> 1 -> x
          wait_acc();
          if (stamp) g_tc_stamps[9] = clock64();
          {
            float a0[HP], a1[HP];
            load_acc(0, a0);
            load_acc(1, a1);
            release_acc();
            const float zb = fmaf(wtt, st.t, bb0);
#pragma unroll
            for (int i = 0; i < HP; ++i) {
              const int p = p0 + i;
              const float z = fmaf(wx1, sxs[2 * p + 1], fmaf(wx0, sxs[2 * p], zb));
              float code, s1, E;
              (void)selu_fwd(z, code);
              selu_decode(code, s1, E);
              const float P1 = fmaf(a1[i] * E, zz1, a0[i] * s1);
              part[2 * i] = P1 * wx0;
              part[2 * i + 1] = P1 * wx1;
            }
            wait_operand_free();   // keeps the barrier phases of both m-tile groups in step
          }
          warp_reduce_scatter32(part, lane);
          red[warp * 32 + lane] = part[0];
          if (stamp) g_tc_stamps[10] = clock64();
          ebar();
          if (warp == 0) {
            const int hh = lane >> 4, q = (lane & 15) * 2;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
              float acc = 0.0f;
#pragma unroll
              for (int w8 = 0; w8 < 8; ++w8) acc += red[(hh * 8 + w8) * 32 + q + d];
              gg[d] = acc;
            }
          }
        }
        if (stamp) g_tc_stamps[11] = clock64();
        // ---- combine + integrator stage update: warp 0, lane = particle
        if (warp == 0) {
          const long row = base + lane;
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const float pr = pr_tot[d];
            const float x = sxs[2 * lane + d];
            float f;
            if (P.variant == IDFF_FIELD_CFM) {
              if (st.kind == 0) sx0[d] = x;
              f = __fsub_rn(pr, sx0[d]);
            } else {
              const float diff = __fsub_rn(pr, x);
              if (st.kind == 0) f = diff;
              else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
              else {
                f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
                if (heavy) f = __fadd_rn(f, gg[d]);
              }
            }
            if (P.mode == 0) {
              if (row < P.N) P.out[row * 2 + d] = f;
            } else {
              const int iv = ev >> 2, sg = ev & 3;
              const float dt = rk.dt[iv], y = sy[d];
              float xn;
              if (sg == 0) {
                sk1[d] = f;
                xn = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, f), third));
              } else if (sg == 1) {
                sk2[d] = f;
                xn = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(f, __fmul_rn(sk1[d], third))));
              } else if (sg == 2) {
                sk3[d] = f;
                xn = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(sk1[d], sk2[d]), f)));
              } else {
                const float sum = __fadd_rn(__fadd_rn(sk1[d], __fmul_rn(3.0f, __fadd_rn(sk2[d], sk3[d]))), f);
                xn = __fadd_rn(y, __fmul_rn(__fmul_rn(sum, dt), 0.125f));
                sy[d] = xn;
                if (P.traj != nullptr && row < P.N) P.traj[((size_t)(iv + 1) * P.N + row) * 2 + d] = xn;
                if (ev == n_eval - 1 && row < P.N) P.out[row * 2 + d] = xn;
              }
              sxs[2 * lane + d] = xn;
            }
          }
        }
        ebar();
        if (stamp) g_tc_stamps[12] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be multicasting into this CTA's ring / arriving on its barriers
  if (warp == NEW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace tc

int tensor_debug_stamps(long long* h_out, int n) {
  IDFF_CUDA(cudaMemcpyFromSymbol(h_out, tc::g_tc_stamps, sizeof(long long) * (n < 32 ? n : 32)));
  return IDFF_OK;
}

bool tensor_supports(const idff_weights_t* w, const idff_field_params_t* p) {
  const NetView& v = w->v;
  if (v.w != 256 || v.tc_tiles == nullptr || p->dim != 2 || v.din != 3 || v.dout != 2) return false;
  if (!(p->variant == IDFF_FIELD_MOMENTUM || p->variant == IDFF_FIELD_CFM)) return false;
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  return nj <= 2;
}

static int launch_tc(const idff_weights_t* w, const idff_field_params_t* p, tc::TParams& P, const Rk4Table* rk,
                     const Stage* one, void* stream) {
  if (!tensor_supports(w, p)) {
    set_error("tensor-core sampler: needs a 3-256-256-256-2 network, dim 2, order <= 2 (momentum / CFM field)");
    return IDFF_EUNSUPPORTED;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long tiles = (P.N + tc::NP - 1) / tc::NP;
  int grid = (int)(tiles < sms ? tiles : sms);
  grid = (grid + 1) & ~1;   // CTA pairs
  static thread_local Rk4Table rk0;
  static thread_local Stage one0;
  const Rk4Table& rkr = rk ? *rk : rk0;
  const Stage& oner = one ? *one : one0;
  const int smem = tc::Smem::total;
  IDFF_CUDA(cudaFuncSetAttribute(tc::k_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IDFF_CUDA(cudaLaunchKernelEx(&cfg, tc::k_tc, P, rkr, oner));
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return verify_tc_launch(const_cast<idff_weights_t*>(w), stream, 1);
}

static void fill_tparams(const idff_weights_t* w, const idff_field_params_t* p, int mode, int64_t N, tc::TParams* P) {
  memset(P, 0, sizeof(*P));
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  P->net = w->v;
  P->variant = p->variant; P->D = p->dim; P->tidx = p->t_idx; P->reverse = rev; P->mode = mode;
  P->N = (long)N;
}

int tensor_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, const float* d_x0,
                      float* d_out, int64_t N, void* stream) {
  tc::TParams P;
  fill_tparams(w, p, 0, N, &P);
  P.x_in = d_x; P.x0_cfm = d_x0; P.out = d_out;
  Stage st;
  make_stage(*p, t, &st);
  return launch_tc(w, p, P, nullptr, &st, stream);
}

int tensor_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                      const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream) {
  static thread_local Rk4Table tab;
  const float* src = d_x0;
  for (int done = 0; done < n_iv; done += kMaxIntervals) {
    const int cnt = n_iv - done < kMaxIntervals ? n_iv - done : kMaxIntervals;
    tab.n_iv = cnt;
    memcpy(tab.dt, dts + done, sizeof(float) * cnt);
    memcpy(tab.st, stages + 4 * (size_t)done, sizeof(Stage) * 4 * cnt);
    tc::TParams P;
    fill_tparams(w, p, 1, N, &P);
    P.x_in = src; P.out = d_out_final;
    P.x0_cfm = done > 0 ? d_x0 : nullptr;   // CFMModel keeps the x captured at t == 0 for the whole solve
    P.traj = d_out_traj ? d_out_traj + (size_t)done * N * p->dim : nullptr;
    int rc = launch_tc(w, p, P, &tab, nullptr, stream);
    if (rc) return rc;
    src = d_out_final;
  }
  return IDFF_OK;
}

}  // namespace idff
