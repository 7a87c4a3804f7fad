// Tensor-core PolyALA sampler (SDE_cal, timeseris_example/MD_simulation.py:481-529): the MD drift field on the
// MLP_Embedding network (47-128-128-128-46 with the embedding folded into the layer-0 bias, models.py:23-42), state dimension
// D <= 48, integrated by torchsde's srk / euler with injected Brownian increments -- the throughput path for many
// trajectories (K >= a few thousand; the latency path for the reference's 5 trajectories is k_md_chain).
//
// Same 3xFP16 tcgen05 formulation as k_tc16 (v = hi + 2^-11 lo', D = W_hi X_hi + 2^-11 (W_hi X_lo' + W_lo' X_hi), fp32
// accumulators in tensor memory, units on M), re-proportioned for a network whose first and last layers are real
// contractions (K = 46 and 46 outputs x K = 128), not the K = 3 / 2-output trifles of the 2-D toys:
//   * hidden width 128 = ONE m-tile; 32 trajectories per CTA.
//   * SEVEN GEMMs per interior evaluation, all on the tensor core:
//        g0  z1   = W0[:, :D] x            (K = 64: D real columns)        -> a1, a1' = s'(z1) (W0 d)
//        g1  z2   = W1 [a1 | a1']          g2  z3 = W2 [a2 | a2']
//        g3  x1hat = W3 a3                 (46 outputs on the M axis)
//        g4  [A0 | A1] = W2^T [P3 | Q3]    g5  [A0 | A1] = W1^T [P2 | Q2]  (reverse sweep, adjoint chains of both jets)
//        g6  grad_x = W0[:, :D]^T P1       (D outputs on the M axis)
//     g3 and g4 are issued together (both follow the layer-3 epilogue) into separate accumulator regions.
//   * B operand (48 KB): per 32-wide k-block 192 rows x 64 B, K-major SWIZZLE_64B: rows 0..127 = [jet0 hi | jet1 hi | jet0 lo' |
//     jet1 lo'] (the two-jet operand: N = 128 / N = 64 MMAs), rows 128..191 = [hi | lo'] of a one-jet operand (x, a3, P1 and
//     every operand of a boundary evaluation: N = 64 / N = 32 MMAs).
//   * weights: 13 tile images [128 rows x 64 k] hi | lo' (32 KB), streamed by a TMA producer warp through a 4-stage ring,
//     multicast inside a CTA pair.
//   * TMEM: two-jet accumulators [main 64 | corr 64] at columns 0..127, one-jet accumulators [main 32 | corr 32] at 128..191,
//     layer-2 stash (code, z2') at 192..255, layer-1 stash (code) at 256..287.
//   * epilogue thread = 4 consecutive hidden units x 2 trajectories x 2 jets on packed pairs; the SRK / Euler update, the
//     drift's closed form and the fp16 split of the next stage input are elementwise passes of the same 512 threads over
//     the tile's 32 x D state, which lives in shared memory.
// Values beyond the fp16 range make that trajectory's output NaN (as in k_tc16); IDFF_IMPL_AUTO runs the fp32 safety net.
#include <cstdlib>
#include <vector>

#include "idff_tcgen05.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);
__device__ int g_md16_misaligned = 0;

namespace tcm {

constexpr int W = 128;
constexpr int NP = 32;                    // trajectories per CTA
constexpr int DMAX = 48;                  // state dimension limit = row pitch of the shared-memory state (layer-0 K is padded to 64)
constexpr int NROW = 6 * NP;              // 192 operand rows per k-block
constexpr int NSTAGE = 4;
constexpr int TILE_BYTES = 32768;
constexpr int KB_BYTES = NROW * 64;       // 12 KB per 32-wide operand k-block
constexpr int ACT_BYTES = (W / 32) * KB_BYTES;
constexpr int NE = 512;
constexpr int NEW = NE / 32;
constexpr int NTHREADS = NE + 128;        // + TMA warp 16, MMA warp 17, two spare warps (register pool)
constexpr int HP = 8;
constexpr int TMEM_COLS = 512;
constexpr int COL_CORR = 2 * NP, COL_E = 128, COL_ST2 = 192, COL_ST1 = 256;
constexpr int OFF_J1 = NP * 64, OFF_LO = 2 * NP * 64, OFF_E = 4 * NP * 64, OFF_ELO = 5 * NP * 64;   // row-group byte offsets
constexpr int NTILES = 13;

using namespace tcg;

struct MParams {
  NetView net;
  int D, tidx, reverse, mode;   // mode 0: one drift evaluation, 2: sdeint
  long N;
  const float* x_in;
  float* out;
  const float* Ik;
  const float* Ik0;
  unsigned long long* counters;
};

struct Smem {   // byte offsets from the 1024-byte aligned dynamic base
  static constexpr int act = 0;
  static constexpr int ring = ACT_BYTES;
  static constexpr int state = ring + NSTAGE * TILE_BYTES;           // floats [NP][DMAX] each
  static constexpr int n_state = 7;                                  // y, xs, k1, k2, yprev, pred, gout
  static constexpr int bars = state + n_state * NP * DMAX * 4;
  static constexpr int total = bars + 128;
};
static_assert(Smem::total <= 232448, "shared memory budget");

template <int HI, int LO>
__device__ __forceinline__ void store_grp(const uint32_t (&sa)[2], const uint32_t (&hw)[4], const uint32_t (&lw)[4]) {
  st64<HI + 0 * 256>(sa[0], hw[0], hw[1]);
  st64<LO + 0 * 256>(sa[0], lw[0], lw[1]);
  st64<HI + 1 * 256>(sa[1], hw[2], hw[3]);
  st64<LO + 1 * 256>(sa[1], lw[2], lw[3]);
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_md16(const __grid_constant__ MParams P, const __grid_constant__ SdeTable tab, const __grid_constant__ Stage one) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((s32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) g_md16_misaligned = 1;
    return;
  }
  float* st_f = reinterpret_cast<float*>(smem + Smem::state);
  float* s_y = st_f, *s_xs = st_f + NP * DMAX, *s_k1 = st_f + 2 * NP * DMAX, *s_k2 = st_f + 3 * NP * DMAX;
  float* s_yp = st_f + 4 * NP * DMAX, *s_pred = st_f + 5 * NP * DMAX, *s_gout = st_f + 6 * NP * DMAX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;                    // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]  MMA -> TMA (both CTAs of the pair)
  uint64_t* acc_h = bars + 2 * NSTAGE;      // MMA -> epilogue: two-jet accumulators complete
  uint64_t* acc_e = acc_h + 1;              // MMA -> epilogue: one-jet accumulators complete
  uint64_t* b_ready = acc_h + 2;            // epilogue -> MMA: the operand of the next MMA phase is written (and every
                                            // accumulator of the previous phase has been read out)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);
    }
    mbar_init(acc_h, 1);
    mbar_init(acc_e, 1);
    mbar_init(b_ready, NE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NEW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < ACT_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t crank = cluster_ctarank();
  const long n_tiles = (P.N + NP - 1) / NP;
  const long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;   // both CTAs of a pair walk the same number of tiles
  const int n_stage = P.mode == 0 ? 1 : (tab.method == 1 ? 1 : 3);
  const int n_eval = P.mode == 0 ? 1 : tab.n_steps * n_stage;
  auto kind_of = [&](int e) -> int { return P.mode == 0 ? one.kind : tab.st[e / n_stage][e % n_stage].kind; };
  auto stage_of = [&](int e) -> Stage { return P.mode == 0 ? one : tab.st[e / n_stage][e % n_stage]; };
  auto heavy_of = [&](int kind) { return kind == 2 && P.reverse != 0; };

  if (warp == NEW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ===== TMA producer: the 13 (interior) / 7 (boundary) weight tiles of every evaluation, in consumption order
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const unsigned char* tiles = reinterpret_cast<const unsigned char*>(P.net.md16_tiles) + crank * (TILE_BYTES / 2);
    unsigned char* const ring = smem + Smem::ring + crank * (TILE_BYTES / 2);
    for (long it = 0; it < n_iter; ++it)
      for (int e = 0; e < n_eval; ++e) {
        const int nt = heavy_of(kind_of(e)) ? NTILES : 7;
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (leader) {
            mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
            tma_load_1d_mc(ring + stage * TILE_BYTES, tiles + (size_t)t * TILE_BYTES, TILE_BYTES / 2, &full[stage], (uint16_t)3);
          }
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
  } else if (warp == NEW + 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ===== MMA issuer (whole warp in uniform control flow, one elected lane issues)
    const uint32_t idesc_128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t bop = s32(smem + Smem::act);
    const uint32_t ring0 = s32(smem + Smem::ring);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0, bphase = 0;
    // one GEMM: n_kb weight tiles (k-blocks of 64) against the two-jet operand (two = true) or the one-jet operand
    auto gemm = [&](int n_kb, bool two) {
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t d1 = tmem + (two ? 0u : (uint32_t)COL_E), d2 = d1 + (two ? (uint32_t)COL_CORR : (uint32_t)NP);
        const uint32_t i1 = two ? idesc_128 : idesc_64, i2 = two ? idesc_64 : idesc_32;
        const uint64_t a_hi0 = make_desc(ring0 + (uint32_t)(stage * TILE_BYTES));
        const uint64_t a_lo0 = make_desc(ring0 + (uint32_t)(stage * TILE_BYTES + TILE_BYTES / 2));
        const uint64_t b_0 = make_desc64(bop + (uint32_t)(2 * kb * KB_BYTES) + (two ? 0u : (uint32_t)OFF_E));
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t boff = (uint64_t)(((ks >> 1) * KB_BYTES + (ks & 1) * 32) >> 4);
            mma_f16(d1, a_hi0 + (uint64_t)(2 * ks), b_0 + boff, i1, (uint32_t)(kb | ks) != 0u);
            mma_f16(d2, a_lo0 + (uint64_t)(2 * ks), b_0 + boff, i2, 1u);
          }
          mma_commit_mc(&empty[stage], (uint16_t)3);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
    };
    auto wait_operand = [&]() {
      mbar_wait(b_ready, bphase);
      bphase ^= 1u;
      tc_fence_after();
    };
    auto done = [&](uint64_t* bar) {
      if (leader) mma_commit(bar);
      __syncwarp();
    };
    for (long it = 0; it < n_iter; ++it)
      for (int e = 0; e < n_eval; ++e) {
        const bool heavy = heavy_of(kind_of(e));
        wait_operand(); gemm(1, false); done(acc_e);             // g0: layer 1 (K = 64)
        wait_operand(); gemm(2, heavy); done(heavy ? acc_h : acc_e);   // g1: layer 2
        wait_operand(); gemm(2, heavy); done(heavy ? acc_h : acc_e);   // g2: layer 3
        wait_operand();
        gemm(2, false); done(acc_e);                             // g3: output layer
        if (heavy) {
          gemm(2, true); done(acc_h);                            // g4: reverse 3 -> 2 (same operand hand-over as g3)
          wait_operand(); gemm(2, true); done(acc_h);            // g5: reverse 2 -> 1
          wait_operand(); gemm(2, false); done(acc_e);           // g6: gradient wrt x
        }
      }
  } else if (warp >= NEW + 2) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");   // spare warps: only return their registers to the pool
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ===== epilogue / integrator warps
    const int q = warp & 3, pq = warp >> 2;
    const int p0 = pq * 8;
    const int g = lane >> 2, c4 = lane & 3;
    const int slot = (g & 1) | (((g >> 1) & 1) << 2) | ((g >> 2) << 1);
    const int n0 = q * 32 + 4 * slot;                       // this thread's 4 consecutive hidden units / output dims
    const int D = P.D;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + p0;
    uint32_t hphase = 0, ephase = 0;
    uint32_t sa0[2];
    {
      const uint32_t b0 = s32(smem + Smem::act) + (uint32_t)(q * KB_BYTES + (p0 + c4) * 64 + (slot & 1) * 8);
      const uint32_t ce = (uint32_t)((slot >> 1) ^ (c4 >> 1));
      sa0[0] = b0 + (ce << 4);
      sa0[1] = b0 + ((ce ^ 2u) << 4);
    }
    auto publish = [&]() {   // this thread's operand writes and accumulator reads are done: hand over to the MMA warp
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(b_ready);
    };
    auto wait_h = [&]() { mbar_wait(acc_h, hphase); hphase ^= 1u; tc_fence_after(); };
    auto wait_e = [&]() { mbar_wait(acc_e, ephase); ephase ^= 1u; tc_fence_after(); };
    // accumulators -> packed pairs (pair i = 2*ld + r: units n0 + 2 ld + {0, 1}, trajectory p0 + 4 r + c4): main + 2^-11 corr
    auto load_pairs = [&](uint32_t col_main, uint32_t col_corr, uint64_t (&z)[4]) {
      uint32_t m[HP], c[HP];
      tmem_ld4(lane_base + col_main, m);
      tmem_ld4(lane_base + (16u << 16) + col_main, m + 4);
      tmem_ld4(lane_base + col_corr, c);
      tmem_ld4(lane_base + (16u << 16) + col_corr, c + 4);
      tmem_wait_ld();
      const uint64_t inv2 = bc2(LO_INV);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        z[i] = fma2(pk2(__uint_as_float(c[2 * i]), __uint_as_float(c[2 * i + 1])), inv2,
                    pk2(__uint_as_float(m[2 * i]), __uint_as_float(m[2 * i + 1])));
    };
    auto load_h = [&](int j, uint64_t (&z)[4]) { load_pairs((uint32_t)(j * NP), (uint32_t)(COL_CORR + j * NP), z); };
    auto load_e = [&](uint64_t (&z)[4]) { load_pairs((uint32_t)COL_E, (uint32_t)(COL_E + NP), z); };
    auto pack_pairs = [&](const uint64_t (&v)[4], uint32_t (&hw)[4], uint32_t (&lw)[4]) {
#pragma unroll
      for (int ld = 0; ld < 2; ++ld)
#pragma unroll
        for (int r = 0; r < 2; ++r) split2p(v[2 * ld + r], hw[2 * r + ld], lw[2 * r + ld]);
    };
    auto put_jet0 = [&](const uint64_t (&v)[4]) { uint32_t hw[4], lw[4]; pack_pairs(v, hw, lw); store_grp<0, OFF_LO>(sa0, hw, lw); };
    auto put_jet1 = [&](const uint64_t (&v)[4]) { uint32_t hw[4], lw[4]; pack_pairs(v, hw, lw); store_grp<OFF_J1, OFF_LO + OFF_J1>(sa0, hw, lw); };
    auto put_one = [&](const uint64_t (&v)[4]) { uint32_t hw[4], lw[4]; pack_pairs(v, hw, lw); store_grp<OFF_E, OFF_ELO>(sa0, hw, lw); };
    auto stash_put = [&](uint32_t col, const uint64_t (&v)[4]) {
      uint32_t u[HP];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        upk2(v[i], lo, hi);
        u[2 * i] = __float_as_uint(lo);
        u[2 * i + 1] = __float_as_uint(hi);
      }
      tmem_st4(lane_base + col, u);
      tmem_st4(lane_base + (16u << 16) + col, u + 4);
    };
    auto stash_put_f = [&](uint32_t col, const float (&c)[HP]) {
      uint32_t u[HP];
#pragma unroll
      for (int i = 0; i < HP; ++i) u[i] = __float_as_uint(c[i]);
      tmem_st4(lane_base + col, u);
      tmem_st4(lane_base + (16u << 16) + col, u + 4);
    };
    auto stash_get = [&](uint32_t col, uint32_t (&u)[HP]) {
      tmem_ld4(lane_base + col, u);
      tmem_ld4(lane_base + (16u << 16) + col, u + 4);
    };
    auto pair4 = [&](const float* base) -> float4 { return __ldg(reinterpret_cast<const float4*>(base + n0)); };
    // accumulators of an output GEMM (lanes = output dims n0..n0+3) -> dst[trajectory][dim] (+ bias)
    auto take_outputs = [&](float* dst, const float* bias) {
      uint64_t z[4];
      load_e(z);
#pragma unroll
      for (int ld = 0; ld < 2; ++ld)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float v0, v1;
          upk2(z[2 * ld + r], v0, v1);
          const int d0 = n0 + 2 * ld, p = p0 + 4 * r + c4;
          if (d0 < D) dst[p * DMAX + d0] = v0 + (bias ? __ldg(bias + d0) : 0.0f);
          if (d0 + 1 < D) dst[p * DMAX + d0 + 1] = v1 + (bias ? __ldg(bias + d0 + 1) : 0.0f);
        }
    };
    // the stage input xs (fp32, [NP][DMAX]) -> one-jet operand rows, k = state dimension (zero beyond D): 16-byte chunks
    auto write_x_operand = [&]() {
      const int item = tid & 255, hl = tid >> 8;           // 32 trajectories x 8 chunks of 8 k, hi | lo'
      const int p = item >> 3, c = item & 7;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int d = 8 * c + i;
        v[i] = d < D ? s_xs[p * DMAX + d] : 0.0f;
      }
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
      const int row = 4 * NP + hl * NP + p;                // rows 128.. (hi), 160.. (lo')
      const uint32_t a = s32(smem + Smem::act) + (uint32_t)((c >> 2) * KB_BYTES + row * 64 + (((c & 3) ^ ((row >> 1) & 3)) << 4));
      if (hl == 0) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
      else asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
    };

    for (long it = 0; it < n_iter; ++it) {
      const long tile = blockIdx.x + it * gridDim.x;
      const long base = tile * NP;
      // ---- tile start: y0 -> y, xs (and ys[0] = y0 for sdeint)
      for (int o = tid; o < NP * D; o += NE) {
        const int p = o / D, d = o - p * D;
        const long row = base + p;
        const float v = row < P.N ? P.x_in[row * D + d] : 0.0f;
        s_y[p * DMAX + d] = v;
        s_xs[p * DMAX + d] = v;
        if (P.mode == 2 && row < P.N) P.out[row * D + d] = v;
      }
      ebar();
      write_x_operand();
      publish();
      int emit = 0;
      if (P.mode == 2)
        while (emit < tab.n_emit && tab.emit_step[emit] < 0) ++emit;
#pragma unroll 1
      for (int ev = 0; ev < n_eval; ++ev) {
        const Stage st = stage_of(ev);
        const bool heavy = heavy_of(st.kind);
        // ---- g0 -> layer 1
        {
          wait_e();
          uint64_t z[4];
          load_e(z);
          const float4 bb0 = __ldg(reinterpret_cast<const float4*>(P.net.b0 + (size_t)P.tidx * W + n0));
          const float4 wtt = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + (size_t)D * W + n0));   // W0[:, D]: the time column
          const uint64_t t2 = bc2(st.t);
          const uint64_t zb[2] = {fma2(pk2(wtt.x, wtt.y), t2, pk2(bb0.x, bb0.y)), fma2(pk2(wtt.z, wtt.w), t2, pk2(bb0.z, bb0.w))};
          if (heavy) {
            const float4 zz1 = pair4(P.net.zd1);
            const uint64_t zz1p[2] = {pk2(zz1.x, zz1.y), pk2(zz1.z, zz1.w)};
            uint64_t zd[4];
            float code[HP];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1;
              selu_jet_code2(add2(z[i], zb[i >> 1]), z[i], s1, code[2 * i], code[2 * i + 1]);
              zd[i] = mul2(s1, zz1p[i >> 1]);
            }
            stash_put_f((uint32_t)COL_ST1, code);
            tmem_wait_st();
            put_jet0(z);
            put_jet1(zd);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1, E;
              selu_jet2(add2(z[i], zb[i >> 1]), z[i], s1, E);
            }
            put_one(z);
          }
          publish();
        }
        // ---- g1 -> layer 2
        {
          const float4 bb1 = pair4(P.net.b1);
          const uint64_t bb1p[2] = {pk2(bb1.x, bb1.y), pk2(bb1.z, bb1.w)};
          if (heavy) {
            wait_h();
            uint64_t z[4], zd[4];
            load_h(0, z);
            load_h(1, zd);
            stash_put((uint32_t)(COL_ST2 + NP), zd);
            float code[HP];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1;
              selu_jet_code2(add2(z[i], bb1p[i >> 1]), z[i], s1, code[2 * i], code[2 * i + 1]);
              zd[i] = mul2(s1, zd[i]);
            }
            stash_put_f((uint32_t)COL_ST2, code);
            tmem_wait_st();
            put_jet0(z);
            put_jet1(zd);
          } else {
            wait_e();
            uint64_t z[4];
            load_e(z);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1, E;
              selu_jet2(add2(z[i], bb1p[i >> 1]), z[i], s1, E);
            }
            put_one(z);
          }
          publish();
        }
        // ---- g2 -> layer 3: a3 (operand of the output layer) and, for interior evaluations, the adjoints (P3, Q3)
        {
          const float4 bb2 = pair4(P.net.b2);
          const uint64_t bb2p[2] = {pk2(bb2.x, bb2.y), pk2(bb2.z, bb2.w)};
          if (heavy) {
            wait_h();
            uint64_t z[4], zd[4], a3[4];
            load_h(0, z);
            load_h(1, zd);
            const float4 uu3 = pair4(P.net.u3);
            const uint64_t u3p[2] = {pk2(uu3.x, uu3.y), pk2(uu3.z, uu3.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint64_t A0 = mul2(bc2(st.c[0]), u3p[i >> 1]), A1 = mul2(bc2(st.c[1]), u3p[i >> 1]);
              uint64_t s1, E;
              selu_jet2(add2(z[i], bb2p[i >> 1]), a3[i], s1, E);
              z[i] = fma2(mul2(A1, E), zd[i], mul2(A0, s1));
              zd[i] = mul2(A1, s1);
            }
            put_one(a3);
            put_jet0(z);
            put_jet1(zd);
          } else {
            wait_e();
            uint64_t z[4];
            load_e(z);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1, E;
              selu_jet2(add2(z[i], bb2p[i >> 1]), z[i], s1, E);
            }
            put_one(z);
          }
          publish();
        }
        // ---- g3 -> x1hat
        wait_e();
        take_outputs(s_pred, P.net.b3);
        if (heavy) {
          // ---- g4 -> adjoints of (a2, a2') -> (P2, Q2)
          {
            wait_h();
            uint64_t a0[4], a1[4];
            load_h(0, a0);
            load_h(1, a1);
            uint32_t cd[HP], zu[HP];
            stash_get((uint32_t)COL_ST2, cd);
            stash_get((uint32_t)(COL_ST2 + NP), zu);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float c0 = __uint_as_float(cd[2 * i]), c1 = __uint_as_float(cd[2 * i + 1]);
              const uint64_t s1 = pk2(c0 < 0.0f ? kSeluLambda : c0, c1 < 0.0f ? kSeluLambda : c1);
              const uint64_t E = pk2(c0 < 0.0f ? 0.0f : c0, c1 < 0.0f ? 0.0f : c1);
              a0[i] = fma2(mul2(a1[i], E), pk2(__uint_as_float(zu[2 * i]), __uint_as_float(zu[2 * i + 1])), mul2(a0[i], s1));
              a1[i] = mul2(a1[i], s1);
            }
            put_jet0(a0);
            put_jet1(a1);
            publish();
          }
          // ---- g5 -> adjoints of (a1, a1') -> P1 = A0 s' + A1 E z1'   (z1' = W0 d per unit)
          {
            wait_h();
            uint64_t a0[4], a1[4];
            load_h(0, a0);
            load_h(1, a1);
            uint32_t cd[HP];
            stash_get((uint32_t)COL_ST1, cd);
            tmem_wait_ld();
            const float4 zz1 = pair4(P.net.zd1);
            const uint64_t zz1p[2] = {pk2(zz1.x, zz1.y), pk2(zz1.z, zz1.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float c0 = __uint_as_float(cd[2 * i]), c1 = __uint_as_float(cd[2 * i + 1]);
              const uint64_t s1 = pk2(c0 < 0.0f ? kSeluLambda : c0, c1 < 0.0f ? kSeluLambda : c1);
              const uint64_t E = pk2(c0 < 0.0f ? 0.0f : c0, c1 < 0.0f ? 0.0f : c1);
              a0[i] = fma2(mul2(a1[i], E), zz1p[i >> 1], mul2(a0[i], s1));
            }
            put_one(a0);
            publish();
          }
          // ---- g6 -> gradient wrt x
          wait_e();
          take_outputs(s_gout, nullptr);
        }
        ebar();   // pred / gout of all 32 trajectories are in shared memory
        // ---- drift value and integrator stage update, elementwise over the tile's 32 x D state
        {
          const int stp = P.mode == 2 ? ev / n_stage : 0, sg = P.mode == 2 ? ev - stp * n_stage : 0;
          const float h = P.mode == 2 ? tab.h[stp] : 0.0f;
          const float rdt = __fdiv_rn(1.0f, h), sq = __fsqrt_rn(h);
          const size_t noff = (size_t)stp * P.N * D;
          for (int o = tid; o < NP * D; o += NE) {
            const int p = o / D, d = o - p * D, si = p * DMAX + d;
            const long row = base + p;
            const float x = s_xs[si];
            const float diff = __fsub_rn(s_pred[si], x);
            float f;
            if (st.kind == 0) f = diff;
            else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
            else {
              f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
              if (heavy) f = __fadd_rn(f, s_gout[si]);
            }
            if (row < P.N && !(fabsf(f) <= 3.0e38f)) atomicAdd(P.counters, 1ull);
            if (P.mode == 0) {
              if (row < P.N) P.out[row * D + d] = f;
              continue;
            }
            const size_t gi = noff + (size_t)row * D + d;
            const float y = s_y[si];
            if (tab.method == 1) {
              // euler: y + f(t0,y) h + g(t0) I_k
              const float ik = row < P.N ? P.Ik[gi] : 0.0f;
              const float yn = __fadd_rn(__fadd_rn(y, __fmul_rn(f, h)), __fmul_rn(tab.gval[stp][0], ik));
              s_yp[si] = y; s_y[si] = yn; s_xs[si] = yn;
            } else if (sg == 0) {
              // H0_1 = y + A0[1][0] F0 h + B0[1][0] G0 Ik0/h,  A0 = 1, B0 = 0     (torchsde srk, as in idff_generic_sde.cu)
              const float ik0 = row < P.N ? P.Ik0[gi] : 0.0f;
              s_k1[si] = f;
              const float t1 = __fmul_rn(__fmul_rn(1.0f, f), h);
              const float t2 = __fmul_rn(__fmul_rn(__fmul_rn(0.0f, tab.gval[stp][0]), ik0), rdt);
              s_xs[si] = __fadd_rn(__fadd_rn(y, t1), t2);
            } else if (sg == 1) {
              // H0_2 = y + (1/4 F0 h + 1 G0 Ik0/h) + (1/4 F1 h + 1/2 G1 Ik0/h)
              const float ik0 = row < P.N ? P.Ik0[gi] : 0.0f;
              s_k2[si] = f;
              float v = y;
              v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, s_k1[si]), h)),
                            __fmul_rn(__fmul_rn(__fmul_rn(1.0f, tab.gval[stp][0]), ik0), rdt));
              v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, f), h)),
                            __fmul_rn(__fmul_rn(__fmul_rn(0.5f, tab.gval[stp][1]), ik0), rdt));
              s_xs[si] = v;
            } else {
              // y1 = y + sum_s alpha_s F_s h + G_s * g_weight_s
              const float ik = row < P.N ? P.Ik[gi] : 0.0f;
              const float ik0 = row < P.N ? P.Ik0[gi] : 0.0f;
              const float ikk = __fmul_rn(__fsub_rn(__fmul_rn(ik, ik), h), 0.5f);
              const float ikkk = __fdiv_rn(__fsub_rn(__fmul_rn(__fmul_rn(ik, ik), ik), __fmul_rn(__fmul_rn(3.0f, h), ik)), 6.0f);
              const float b1[4] = {-1.0f, (float)(4.0 / 3.0), (float)(2.0 / 3.0), 0.0f};
              const float b2[4] = {1.0f, (float)(-4.0 / 3.0), (float)(1.0 / 3.0), 0.0f};
              const float b3[4] = {2.0f, (float)(-4.0 / 3.0), (float)(-2.0 / 3.0), 0.0f};
              const float b4[4] = {-2.0f, (float)(5.0 / 3.0), (float)(-2.0 / 3.0), 1.0f};
              const float al[4] = {(float)(1.0 / 6.0), (float)(1.0 / 6.0), (float)(2.0 / 3.0), 0.0f};
              const float F[4] = {s_k1[si], s_k2[si], f, 0.0f};
              float v = y;
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                float gw = __fmul_rn(b1[qq], ik);
                gw = __fadd_rn(gw, __fdiv_rn(__fmul_rn(b2[qq], ikk), sq));
                gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b3[qq], ik0), rdt));
                gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b4[qq], ikkk), rdt));
                v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(al[qq], F[qq]), h)), __fmul_rn(tab.gval[stp][qq], gw));
              }
              s_yp[si] = y; s_y[si] = v; s_xs[si] = v;
            }
          }
          if (P.mode == 2 && sg == n_stage - 1) {
            // outputs requested inside this step: linear interpolation between the previous and the new state (own elements)
            while (emit < tab.n_emit && tab.emit_step[emit] == stp) {
              const float w0 = tab.emit_w0[emit], w1 = tab.emit_w1[emit];
              const size_t ooff = (size_t)tab.emit_out[emit] * P.N * D;
              for (int o = tid; o < NP * D; o += NE) {
                const int p = o / D, d = o - p * D, si = p * DMAX + d;
                const long row = base + p;
                if (row < P.N) P.out[ooff + row * D + d] = __fadd_rn(__fmul_rn(w0, s_yp[si]), __fmul_rn(w1, s_y[si]));
              }
              ++emit;
            }
          }
        }
        ebar();   // xs of all trajectories is updated
        if (ev + 1 < n_eval) {
          write_x_operand();
          publish();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == NEW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace tcm

bool md16_supports(const idff_weights_t* w, const idff_field_params_t* p) {
  const NetView& v = w->v;
  return v.w == 128 && v.md16_tiles != nullptr && p->variant == IDFF_FIELD_MD && p->dim == v.din - 1 && p->dim == v.dout &&
         p->dim <= tcm::DMAX;
}

static int verify_launch(idff_weights_t* w, void* stream) {
  if (w->tc_verified & 16) return IDFF_OK;
  int flag = 0;
  IDFF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  IDFF_CUDA(cudaMemcpyFromSymbol(&flag, g_md16_misaligned, sizeof(int)));
  if (flag) {
    set_error("tcgen05 kernel: dynamic shared memory is not 1024-byte aligned on this driver; use impl = IDFF_IMPL_FUSED");
    return IDFF_EUNSUPPORTED;
  }
  w->tc_verified |= 16;
  return IDFF_OK;
}

static int launch_md16(const idff_weights_t* w, const idff_field_params_t* p, tcm::MParams& P, const SdeTable* tab, const Stage* one,
                       void* stream) {
  if (!md16_supports(w, p)) {
    set_error("3xFP16 PolyALA sampler: needs a (D+1)-128-128-128-D network with D <= %d and the MD field", tcm::DMAX);
    return IDFF_EUNSUPPORTED;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long tiles = (P.N + tcm::NP - 1) / tcm::NP;
  int grid = (int)(tiles < sms ? tiles : sms);
  grid = (grid + 1) & ~1;   // CTA pairs
  static thread_local SdeTable tab0;
  static thread_local Stage one0;
  const SdeTable& tabr = tab ? *tab : tab0;
  const Stage& oner = one ? *one : one0;
  const int smem = tcm::Smem::total;
  IDFF_CUDA(cudaFuncSetAttribute(tcm::k_md16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tcm::NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IDFF_CUDA(cudaLaunchKernelEx(&cfg, tcm::k_md16, P, tabr, oner));
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return verify_launch(const_cast<idff_weights_t*>(w), stream);
}

static void fill_params(const idff_weights_t* w, const idff_field_params_t* p, int mode, int64_t N, tcm::MParams* P) {
  memset(P, 0, sizeof(*P));
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  P->net = w->v;
  P->D = p->dim; P->tidx = p->t_idx; P->reverse = rev; P->mode = mode;
  P->N = (long)N;
  P->counters = w->d_counters;
}

int md16_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, float* d_out, int64_t N,
                    void* stream) {
  tcm::MParams P;
  fill_params(w, p, 0, N, &P);
  P.x_in = d_x; P.out = d_out;
  Stage st;
  make_stage(*p, t, &st);
  return launch_md16(w, p, P, nullptr, &st, stream);
}

int md16_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, const float* d_y0,
                    const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream) {
  tcm::MParams P;
  fill_params(w, p, 2, N, &P);
  P.x_in = d_y0; P.out = d_out; P.Ik = d_Ik; P.Ik0 = d_Ik0;
  return launch_md16(w, p, P, &tab, nullptr, stream);
}

}  // namespace idff
