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
#include "idff_common.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);

namespace tc {

constexpr int W = 256;
constexpr int NP = 32;                   // particles per CTA
constexpr int NCOL = 64;                 // MMA N: 2 jets x 32 particles
constexpr int NSTAGE = 2;                // weight ring depth
constexpr int TILE_BYTES = 32768;        // hi (16 KB) + lo (16 KB) image of one [128 x 32] weight tile
constexpr int ACT_BYTES = NCOL * W * 4;  // one copy (hi or lo) of the B operand: 8 k-blocks x [64 x 128 B]
constexpr int NE = 256;                  // epilogue threads: warps 0..3 own m-tile 0, warps 4..7 m-tile 1 (lane quarter = warp % 4)
constexpr int NTHREADS = 320;            // + producer warp 8 + MMA warp 9
constexpr int TMEM_COLS = 512;
constexpr int COL_MAIN = 0, COL_CORR = 64, COL_MT = 128, COL_STASH = 256;   // + mt * COL_MT / + mt * 64

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
  while (!mbar_try_wait(b, parity)) __nanosleep(64);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ebar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // epilogue warps only

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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

struct TParams {
  NetView net;
  int variant, D, tidx, reverse, mode;   // mode 0 eval, 1 rk4
  long N;
  const float* x_in;
  const float* x0_cfm;
  float* out;
  float* traj;
};

struct Smem {   // byte offsets from the 1024-aligned dynamic base
  static constexpr int act_hi = 0;
  static constexpr int act_lo = ACT_BYTES;
  static constexpr int ring = 2 * ACT_BYTES;
  static constexpr int misc = ring + NSTAGE * TILE_BYTES;   // floats from here
  // misc floats: w0 [256][4] (w0x0, w0x1, w0t, b0), vecs b1,b2,u3,zd1 [4][256], w3 [2][256], red [2][8][64], state [6][64]
  static constexpr int f_w0 = 0, f_vec = 1024, f_w3 = 2048, f_red = 2560, f_state = 3584, f_end = 3584 + 12 * 32;
  static constexpr int bars = misc + f_end * 4;
  static constexpr int total = bars + 128;
};

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

__global__ void __launch_bounds__(NTHREADS, 1)
k_tc(const __grid_constant__ TParams P, const __grid_constant__ Rk4Table rk, const __grid_constant__ Stage one) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // SWIZZLE_128B operands need a 1024-byte aligned base: align by hand (the launch reserves the slack)
  unsigned char* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  float* misc = reinterpret_cast<float*>(smem + Smem::misc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;                 // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;       // [NSTAGE]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * NSTAGE;   // MMA -> epilogue: the accumulators of one layer pass are complete
  uint64_t* b_ready = acc_full + 1;         // epilogue -> MMA: the B operand of the next layer pass is written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(b_ready, NE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // resident small operands
  for (int n = tid; n < W; n += NTHREADS) {
    misc[Smem::f_w0 + 4 * n + 0] = P.net.wt0[0 * W + n];
    misc[Smem::f_w0 + 4 * n + 1] = P.net.wt0[1 * W + n];
    misc[Smem::f_w0 + 4 * n + 2] = P.net.wt0[2 * W + n];
    misc[Smem::f_w0 + 4 * n + 3] = P.net.b0[(size_t)P.tidx * W + n];
    misc[Smem::f_vec + 0 * W + n] = P.net.b1[n];
    misc[Smem::f_vec + 1 * W + n] = P.net.b2[n];
    misc[Smem::f_vec + 2 * W + n] = P.net.u3[n];
    misc[Smem::f_vec + 3 * W + n] = P.net.zd1[n];
    misc[Smem::f_w3 + 0 * W + n] = P.net.w3[0 * W + n];
    misc[Smem::f_w3 + 1 * W + n] = P.net.w3[1 * W + n];
  }
  // jet-1 columns are read by the MMAs of boundary evaluations too: start from finite values
  for (int i = tid; i < 2 * ACT_BYTES / 4; i += NTHREADS) reinterpret_cast<float*>(smem)[i] = 0.0f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_eval = P.mode == 0 ? 1 : 4 * rk.n_iv;
  auto stage_of = [&](int e) -> const Stage& { return P.mode == 0 ? one : rk.st[e]; };
  auto heavy_of = [&](const Stage& st) { return st.kind == 2 && P.reverse; };

  if (warp == 8) {
    // ===== TMA producer: weight tile images of every layer pass, in consumption order
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long tile = blockIdx.x; tile * NP < P.N; tile += gridDim.x)
        for (int e = 0; e < n_eval; ++e) {
          const int ng = heavy_of(stage_of(e)) ? 4 : 2;
          for (int g = 0; g < ng; ++g)
            for (int t = 0; t < 16; ++t) {   // (m-tile, k-block) pairs of gemm g are contiguous
              mbar_wait_sleep(&empty[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
              tma_load_1d(smem + Smem::ring + stage * TILE_BYTES,
                          reinterpret_cast<const unsigned char*>(P.net.tc_tiles) + ((size_t)g * 16 + t) * TILE_BYTES, TILE_BYTES,
                          &full[stage]);
              if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
            }
        }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bhi = s32(smem + Smem::act_hi), blo = s32(smem + Smem::act_lo);
      int stage = 0;
      uint32_t phase = 0, bphase = 0;
      for (long tile = blockIdx.x; tile * NP < P.N; tile += gridDim.x)
        for (int e = 0; e < n_eval; ++e) {
          const int ng = heavy_of(stage_of(e)) ? 4 : 2;
          for (int g = 0; g < ng; ++g) {
            mbar_wait(b_ready, bphase);
            bphase ^= 1u;
            tc_fence_after();
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t d_main = tmem + mt * COL_MT + COL_MAIN, d_corr = tmem + mt * COL_MT + COL_CORR;
              for (int kb = 0; kb < 8; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t whi = s32(smem + Smem::ring + stage * TILE_BYTES), wlo = whi + 16384;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t a_hi = make_desc(whi + ks * 32), a_lo = make_desc(wlo + ks * 32);
                  const uint64_t b_hi = make_desc(bhi + kb * (NCOL * 128) + ks * 32), b_lo = make_desc(blo + kb * (NCOL * 128) + ks * 32);
                  mma_tf32(d_main, a_hi, b_hi, idesc, (kb | ks) != 0);
                  mma_tf32(d_corr, a_lo, b_hi, idesc, (kb | ks) != 0);
                  mma_tf32(d_corr, a_hi, b_lo, idesc, 1u);
                }
                mma_commit(&empty[stage]);   // the ring slot is free once these MMAs have read it
                if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
              }
            }
            mma_commit(acc_full);
          }
        }
    }
  } else {
    // ===== epilogue warps: warp w owns m-tile w / 4 and TMEM lanes 32 (w % 4) ..: thread <-> one hidden unit n
    const int mt = warp >> 2;
    const int n = mt * 128 + (warp & 3) * 32 + lane;
    float* red = misc + Smem::f_red;       // [2][8 warps][64]
    float* stt = misc + Smem::f_state;     // [6][64]: y, k1, k2, k3, xs, x0c
    float *sy = stt, *sk1 = stt + 64, *sk2 = stt + 128, *sk3 = stt + 192, *sxs = stt + 256, *sx0 = stt + 320;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + mt * COL_MT;
    const uint32_t stash_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + COL_STASH + mt * 64;
    uint32_t aphase = 0;
    const float third = (float)(1.0 / 3.0);
    const float wx0 = misc[Smem::f_w0 + 4 * n + 0], wx1 = misc[Smem::f_w0 + 4 * n + 1], wtt = misc[Smem::f_w0 + 4 * n + 2],
                bb0 = misc[Smem::f_w0 + 4 * n + 3], bb1 = misc[Smem::f_vec + 0 * W + n], bb2 = misc[Smem::f_vec + 1 * W + n],
                uu3 = misc[Smem::f_vec + 2 * W + n], zz1 = misc[Smem::f_vec + 3 * W + n], w30 = misc[Smem::f_w3 + 0 * W + n],
                w31 = misc[Smem::f_w3 + 1 * W + n];
    // this unit's k position inside the B operand: k-block, 16-byte chunk, word
    unsigned char* ahi = smem + Smem::act_hi + (n >> 5) * (NCOL * 128) + ((n & 3) << 2);
    unsigned char* alo = ahi + ACT_BYTES;
    const int kchunk = (n >> 2) & 7;
    auto put = [&](int col, float v) {
      const uint32_t o = (uint32_t)(col * 128 + ((kchunk ^ (col & 7)) << 4));
      const float h = rn_tf32(v);
      *reinterpret_cast<float*>(ahi + o) = h;
      *reinterpret_cast<float*>(alo + o) = v - h;
    };
    // after writing the B operand: make it visible to the tensor core and hand over to the MMA thread
    auto publish = [&]() {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(b_ready);
    };
    // z = main + corr, columns [c0, c0+32) of this thread's lane
    auto load_acc = [&](int c0, float (&z)[32]) {
      uint32_t m[32], c[32];
      tmem_ld32(lane_base + COL_MAIN + c0, m);
      tmem_ld32(lane_base + COL_CORR + c0, c);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = __uint_as_float(m[i]) + __uint_as_float(c[i]);
    };

    for (long tile = blockIdx.x; tile * NP < P.N; tile += gridDim.x) {
      const long base = tile * NP;
      if (tid < 64) {
        const int p = tid >> 1, d = tid & 1;
        const long row = base + p;
        const float v = row < P.N ? P.x_in[row * 2 + d] : 0.0f;
        sy[tid] = v;
        sxs[tid] = v;
        if (P.mode == 0 && P.variant == IDFF_FIELD_CFM && P.x0_cfm) sx0[tid] = row < P.N ? P.x0_cfm[row * 2 + d] : 0.0f;
      }
      ebar();
#pragma unroll 1
      for (int ev = 0; ev < n_eval; ++ev) {
        const Stage& st = stage_of(ev);
        const bool heavy = heavy_of(st);
        // ---- layer 1 (K = 3, computed directly): a1 and, for interior evaluations, a1' = s'(z1) * (W0 d)
        {
          const float zb = fmaf(wtt, st.t, bb0);
#pragma unroll 8
          for (int p = 0; p < NP; ++p) {
            const float z = fmaf(wx1, sxs[2 * p + 1], fmaf(wx0, sxs[2 * p], zb));
            float code, s1, E;
            const float a = selu_fwd(z, code);
            selu_decode(code, s1, E);
            put(p, a);
            if (heavy) put(NP + p, s1 * zz1);
          }
        }
        publish();
        float part[64];
        // ---- gemm 0: layer 2 forward
        mbar_wait(acc_full, aphase); aphase ^= 1u; tc_fence_after();
        {
          float z[32], zd[32];
          load_acc(0, z);
          if (heavy) load_acc(32, zd);
          uint32_t st_code[32];
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            float code, s1, E;
            const float a = selu_fwd(z[p] + bb1, code);
            selu_decode(code, s1, E);
            st_code[p] = __float_as_uint(code);
            put(p, a);
            if (heavy) put(NP + p, s1 * zd[p]);
          }
          if (heavy) {
            uint32_t st_zd[32];
#pragma unroll
            for (int p = 0; p < NP; ++p) st_zd[p] = __float_as_uint(zd[p]);
            tmem_st32(stash_base, st_code);
            tmem_st32(stash_base + 32, st_zd);
            tmem_wait_st();
          }
        }
        publish();
        // ---- gemm 1: layer 3 forward; x1hat partials; adjoint seeds
        mbar_wait(acc_full, aphase); aphase ^= 1u; tc_fence_after();
        {
          float z[32], zd[32];
          load_acc(0, z);
          if (heavy) load_acc(32, zd);
          const float A0 = st.c[0] * uu3, A1 = st.c[1] * uu3;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            float code, s1, E;
            const float a = selu_fwd(z[p] + bb2, code);
            selu_decode(code, s1, E);
            part[2 * p] = a * w30;
            part[2 * p + 1] = a * w31;
            if (heavy) {
              put(p, fmaf(A1 * E, zd[p], A0 * s1));   // adjoint of z3
              put(NP + p, A1 * s1);                   // adjoint of z3'
            }
          }
        }
        if (heavy) publish();
        warp_reduce_scatter64(part, lane);
        red[warp * 64 + 2 * lane] = part[0];
        red[warp * 64 + 2 * lane + 1] = part[1];
        if (heavy) {
          // ---- gemm 2: reverse, layer 3 -> 2
          mbar_wait(acc_full, aphase); aphase ^= 1u; tc_fence_after();
          {
            float a0[32], a1[32];
            load_acc(0, a0);
            load_acc(32, a1);
            uint32_t cd[32], zdu[32];
            tmem_ld32(stash_base, cd);
            tmem_ld32(stash_base + 32, zdu);
            tmem_wait_ld();
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              float s1, E;
              selu_decode(__uint_as_float(cd[p]), s1, E);
              put(p, fmaf(a1[p] * E, __uint_as_float(zdu[p]), a0[p] * s1));
              put(NP + p, a1[p] * s1);
            }
          }
          publish();
          // ---- gemm 3: reverse, layer 2 -> 1 -> x
          mbar_wait(acc_full, aphase); aphase ^= 1u; tc_fence_after();
          {
            float a0[32], a1[32];
            load_acc(0, a0);
            load_acc(32, a1);
            const float zb = fmaf(wtt, st.t, bb0);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const float z = fmaf(wx1, sxs[2 * p + 1], fmaf(wx0, sxs[2 * p], zb));
              float code, s1, E;
              (void)selu_fwd(z, code);
              selu_decode(code, s1, E);
              const float P1 = fmaf(a1[p] * E, zz1, a0[p] * s1);
              part[2 * p] = P1 * wx0;
              part[2 * p + 1] = P1 * wx1;
            }
          }
          warp_reduce_scatter64(part, lane);
          red[512 + warp * 64 + 2 * lane] = part[0];
          red[512 + warp * 64 + 2 * lane + 1] = part[1];
        }
        ebar();
        // ---- combine + integrator stage update: warp 0, lane = particle
        if (warp == 0) {
          const int p = lane;
          const long row = base + p;
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const int o = 2 * p + d;
            float pr = __ldg(P.net.b3 + d), gg = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) pr += red[q * 64 + o];
            if (heavy) {
#pragma unroll
              for (int q = 0; q < 8; ++q) gg += red[512 + q * 64 + o];
            }
            const float x = sxs[o];
            float f;
            if (P.variant == IDFF_FIELD_CFM) {
              if (st.kind == 0) sx0[o] = x;
              f = __fsub_rn(pr, sx0[o]);
            } else {
              const float diff = __fsub_rn(pr, x);
              if (st.kind == 0) f = diff;
              else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
              else {
                f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
                if (heavy) f = __fadd_rn(f, gg);
              }
            }
            if (P.mode == 0) {
              if (row < P.N) P.out[row * 2 + d] = f;
            } else {
              const int iv = ev >> 2, sg = ev & 3;
              const float dt = rk.dt[iv], y = sy[o];
              float xn;
              if (sg == 0) {
                sk1[o] = f;
                xn = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, f), third));
              } else if (sg == 1) {
                sk2[o] = f;
                xn = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(f, __fmul_rn(sk1[o], third))));
              } else if (sg == 2) {
                sk3[o] = f;
                xn = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(sk1[o], sk2[o]), f)));
              } else {
                const float sum = __fadd_rn(__fadd_rn(sk1[o], __fmul_rn(3.0f, __fadd_rn(sk2[o], sk3[o]))), f);
                xn = __fadd_rn(y, __fmul_rn(__fmul_rn(sum, dt), 0.125f));
                sy[o] = xn;
                if (P.traj != nullptr && row < P.N) P.traj[((size_t)(iv + 1) * P.N + row) * 2 + d] = xn;
                if (ev == n_eval - 1 && row < P.N) P.out[row * 2 + d] = xn;
              }
              sxs[o] = xn;
            }
          }
        }
        ebar();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace tc

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
  const int grid = (int)(tiles < sms ? tiles : sms);
  static thread_local Rk4Table rk0;
  static thread_local Stage one0;
  const Rk4Table& rkr = rk ? *rk : rk0;
  const Stage& oner = one ? *one : one0;
  const int smem = tc::Smem::total + 1024;
  IDFF_CUDA(cudaFuncSetAttribute(tc::k_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::k_tc<<<grid, tc::NTHREADS, smem, (cudaStream_t)stream>>>(P, rkr, oner);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
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
    P.traj = d_out_traj ? d_out_traj + (size_t)done * N * p->dim : nullptr;
    int rc = launch_tc(w, p, P, &tab, nullptr, stream);
    if (rc) return rc;
    src = d_out_final;
  }
  return IDFF_OK;
}

}  // namespace idff
