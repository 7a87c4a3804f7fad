// Tensor-core IDFF sampler for the ORDER-3 field (Autograd_example_order3.py:34-77: gamma1, gamma2, gamma3 momentum terms),
// 3xFP16 split, hidden width 256, state dimension 2: the three-jet sibling of k_tc16 (idff_sampler_tc16.cu) -- same
// formulation (units on M, activations = B operand resident in shared memory, weights streamed by TMA multicast in CTA
// pairs, fp32 accumulators in tensor memory, whole RK4 solve in one launch), re-proportioned for three Taylor jets:
//
//   * 32 particles per CTA.  B operand per 32-wide k-block: 192 rows x 64 B, row = hl*96 + jet*32 + particle (hl 0 = hi,
//     1 = lo'), K-major SWIZZLE_64B; 8 k-blocks = 96 KB.  The 32 KB this saves against k_tc16 buys a FOURTH weight stage.
//   * per 16-wide k-step   [main 96 | corr 96] += W_hi * [X_hi ; X_lo']   (N = 192)   and   corr += W_lo' * X_hi   (N = 96).
//   * TMEM: accumulators at columns 0..191; the layer-2 stash of the reverse sweep -- derivative code, z2', z2'' -- at
//     columns 256 + 96 mt + {0, 32, 64} (192 columns): 384 of 512 columns in use.
//   * boundary evaluations (t = 0, t = 1: one jet) keep [X_hi ; X_lo'] in rows 0..63, run N = 64 / N = 32 MMAs and split the
//     K sum over two accumulator sets (columns 0 and 128) by k-block parity, as k_tc16 does.
//   * epilogue thread = 4 consecutive hidden units x 2 particles x 3 jets (tcgen05.ld.16x128b.x2); 16 epilogue warps = 4 lane
//     quarters x 4 groups of 8 particles; jets as packed pairs (FADD2 / FMUL2 / FFMA2).  One integrator warp (a lane per
//     particle) holds the RK4 state.  Register budgets per role as in k_tc16 (setmaxnreg 32 / 112).
//
// Jet algebra per hidden unit (tests/jet_model.py), E = every SELU derivative beyond the first:
//     a = s(z)      a' = s'(z) z'      a'' = E z'^2 + s'(z) z''
//     P = A0 s' + A1 E z' + A2 E (z'^2 + z'')     Q = A1 s' + 2 A2 E z'     R = A2 s'        (adjoints of z, z', z'')
// with seeds A_i = c_i (W3^T 1) at the last hidden layer, z1' = W0 d a per-unit constant and z1'' = 0.
// Values beyond the fp16 range make that particle's output NaN, as in k_tc16; IDFF_IMPL_AUTO adds the fp32 safety net.
#include <cstdlib>
#include <vector>

#include "idff_tcgen05.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);
int fixup_nonfinite(const idff_weights_t* w, const idff_field_params_t* p, int mode, int n_sets, int n_iv, const Rk4Table* rk,
                    const SweepTab* d_tabs, const Stage* g_st, const float* g_dt, const float* d_x_in, const float* d_x0_cfm,
                    float* d_out, float* d_traj, int64_t N, void* stream);
int keep_input(idff_weights_t* wm, const float** src, const float* d_out, bool force, size_t bytes, void* stream);
__device__ int g_tc16o3_misaligned = 0;

namespace tc3 {

constexpr int W = 256;
constexpr int NP = 32;                    // particles per CTA
constexpr int NJ = 3;
constexpr int NROW = 2 * NJ * NP;         // 192 rows of the B operand: hl*96 + jet*32 + particle
constexpr int NSTAGE = 4;                 // weight ring depth
constexpr int TILE_BYTES = 32768;         // hi (16 KB) + lo' (16 KB) image of one [128 units x 64 k] fp16 weight tile
constexpr int KB_BYTES = NROW * 64;       // B operand of one 32-wide k-block: 192 rows x 64 B (12 KB)
constexpr int NKB = W / 64;               // 4 weight k-blocks of 64 (= 8 operand k-blocks of 32)
constexpr int ACT_BYTES = (W / 32) * KB_BYTES;
constexpr int NE = 512;                   // epilogue threads: warp w -> lane quarter w & 3, particle group w >> 2
constexpr int NEW = NE / 32;
constexpr int NTHREADS = NE + 128;        // + producer warp 16, MMA warp 17, integrator warp 18, spare warp 19
constexpr int HP = 8;                     // values per jet and epilogue thread: 4 units x 2 particles
constexpr int TMEM_COLS = 512;
constexpr int COL_CORR = NJ * NP, COL_STASH = 256, STASH_MT = NJ * NP;
constexpr int LO_ROWS_BYTES = NJ * NP * 64;   // byte offset of the lo' rows inside a k-block (interior evaluations)
constexpr int JET_BYTES = NP * 64;            // byte offset between the jets' row groups

using namespace tcg;

struct TParams {
  NetView net;
  int variant, D, tidx, reverse, mode;   // mode 0 eval, 1 rk4, 2 rk4 sweep over n_sets parameter sets (same x0, same grid)
  int n_sets, sweep_n_iv;
  long tiles_per_set;                    // sweep: even, so both CTAs of a pair always work on the same set
  const SweepTab* tabs;                  // sweep: [n_sets] in global memory
  long N;
  const float* x_in;
  float* out;
  float* traj;
  unsigned long long* counters;
};

struct Smem {   // byte offsets from the dynamic base, which must be 1024-byte aligned (checked at kernel entry)
  static constexpr int act = 0;
  static constexpr int ring = ACT_BYTES;
  static constexpr int misc = ring + NSTAGE * TILE_BYTES;   // floats: red [16 warps][32], xs [32][2]
  static constexpr int f_red = 0, f_xs = NEW * 32, f_end = NEW * 32 + 2 * NP;
  static constexpr int bars = misc + f_end * 4;
  static constexpr int total = bars + 128;
};
static_assert(Smem::total <= 232448, "shared memory budget");

__device__ __forceinline__ void ebar544() { asm volatile("bar.sync 1, 544;" ::: "memory"); }   // epilogue + integrator warps

// 8-byte stores of one jet into the thread's two particle rows: hi rows at +0, lo' rows LO bytes further.
// sa[0] / sa[1] = swizzled address of the thread's 8-byte group for r = 0 / 1, jet 0, hi.  hw[2r + ld]: units (2 ld, 2 ld + 1).
template <int JET, int LO>
__device__ __forceinline__ void store_jet(const uint32_t (&sa)[2], const uint32_t (&hw)[4], const uint32_t (&lw)[4]) {
  st64<JET * JET_BYTES + 0 * 256>(sa[0], hw[0], hw[1]);
  st64<LO + JET * JET_BYTES + 0 * 256>(sa[0], lw[0], lw[1]);
  st64<JET * JET_BYTES + 1 * 256>(sa[1], hw[2], hw[3]);
  st64<LO + JET * JET_BYTES + 1 * 256>(sa[1], lw[2], lw[3]);
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_tc16o3(const __grid_constant__ TParams P, const __grid_constant__ Rk4Table rk, const __grid_constant__ Stage one) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((s32(smem) & 1023u) != 0u) {   // swizzled operands need a 1024-byte aligned base: report and do nothing
    if (threadIdx.x == 0) g_tc16o3_misaligned = 1;
    return;
  }
  float* misc = reinterpret_cast<float*>(smem + Smem::misc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;                    // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]  MMA -> TMA (both CTAs of the pair)
  uint64_t* acc_full = bars + 2 * NSTAGE;   // MMA -> epilogue: the accumulators of the current m-tile are complete
  uint64_t* acc_free = acc_full + 1;        // epilogue -> MMA: every thread has read its accumulators out
  uint64_t* b_ready = acc_full + 2;         // [2] epilogue -> MMA: k-blocks 0-1 / 2-3 of the next pass are written
  uint64_t* kfree = acc_full + 4;           // MMA -> epilogue: k-blocks 0-1 have been read by both m-tiles of this pass
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);
    }
    mbar_init(acc_full, 1);
    mbar_init(kfree, 1);
    mbar_init(acc_free, NE);
    mbar_init(&b_ready[0], NE);
    mbar_init(&b_ready[1], NE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NEW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows of jets 1, 2 are read by the MMAs of boundary evaluations too: start from finite values
  for (int i = tid; i < ACT_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barrier inits visible to the peer before any remote complete_tx / commit arrives
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t crank = cluster_ctarank();
  // both CTAs of a pair walk the same number of tiles (the ring couples them); surplus tiles are all-padding
  const long tiles_per_set = P.mode == 2 ? P.tiles_per_set : (P.N + NP - 1) / NP;
  const long n_tiles = P.mode == 2 ? tiles_per_set * P.n_sets : tiles_per_set;
  const long n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;

  const int n_eval = P.mode == 0 ? 1 : (P.mode == 1 ? 4 * rk.n_iv : 4 * P.sweep_n_iv);
  auto set_of = [&](long it) -> int {
    if (P.mode != 2) return 0;
    const long g = (blockIdx.x + it * gridDim.x) / tiles_per_set;
    return (int)(g < P.n_sets ? g : P.n_sets - 1);
  };
  auto stage_of = [&](int e, int g) -> Stage {   // by value: see k_tc16
    if (P.mode == 0) return one;
    if (P.mode == 1) return rk.st[e];
    return P.tabs[g].st[e];
  };
  auto kind_of = [&](int e, int g) -> int { return P.mode == 0 ? one.kind : (P.mode == 1 ? rk.st[e].kind : P.tabs[g].st[e].kind); };
  auto heavy_of = [&](int kind, int g) { return kind == 2 && (P.mode == 2 ? P.tabs[g].reverse : P.reverse) != 0; };

  if (warp == NEW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ===== TMA producer: weight tile images of every layer pass, in consumption order (whole warp in uniform control flow,
    // one elected lane issues)
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      const unsigned char* tiles = reinterpret_cast<const unsigned char*>(P.net.tc16_tiles) + crank * (TILE_BYTES / 2);
      unsigned char* const ring = smem + Smem::ring + crank * (TILE_BYTES / 2);
      for (long it = 0; it < n_iter; ++it)
        for (int e = 0; e < n_eval; ++e) {
          const int gs = set_of(it);
          const int ng = heavy_of(kind_of(e, gs), gs) ? 4 : 2;
          for (int g = 0; g < ng; ++g)
            for (int t = 0; t < 2 * NKB; ++t) {   // (m-tile, k-block) tiles of gemm g are contiguous
              mbar_wait(&empty[stage], phase ^ 1u);
              if (leader) {
                mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
                tma_load_1d_mc(ring + stage * TILE_BYTES, tiles + ((size_t)g * 2 * NKB + t) * TILE_BYTES, TILE_BYTES / 2,
                               &full[stage], (uint16_t)3);
              }
              __syncwarp();
              if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
            }
        }
    }
  } else if (warp == NEW + 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ===== MMA issuer: the whole warp walks the schedule in uniform control flow, one elected lane issues
    {
      // kind::f16: D fp32 (bit 4), A = B = fp16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc_192 = (1u << 4) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_96 = (1u << 4) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bop = s32(smem + Smem::act);
      const uint32_t ring0 = s32(smem + Smem::ring);
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0, fphase = 0, bphase0 = 0, bphase1 = 0;
      for (long it = 0; it < n_iter; ++it)
        for (int e = 0; e < n_eval; ++e) {
          const int gs = set_of(it);
          const bool heavy = heavy_of(kind_of(e, gs), gs);
          const int ng = heavy ? 4 : 2;
          for (int g = 0; g < ng; ++g) {
            for (int mt = 0; mt < 2; ++mt) {
              mbar_wait(acc_free, fphase);   // the previous m-tile's accumulators have been read out
              fphase ^= 1u;
              tc_fence_after();
              for (int kb = 0; kb < NKB; ++kb) {
                if (mt == 0 && kb == 0) {
                  mbar_wait(&b_ready[0], bphase0);
                  bphase0 ^= 1u;
                  tc_fence_after();
                }
                if (mt == 0 && kb == 2) {
                  mbar_wait(&b_ready[1], bphase1);
                  bphase1 ^= 1u;
                  tc_fence_after();
                }
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                // interior: [main 96 | corr 96] += W_hi * [X_hi ; X_lo'] (N = 192), corr += W_lo' * X_hi (N = 96).
                // boundary (one jet): rows 0..31 X_hi, rows 32..63 X_lo': N = 64 and N = 32; accumulator set (kb & 1):
                // [main 32 | corr 32] at columns 128 * (kb & 1)
                const uint32_t d1 = tmem + (heavy ? 0u : (uint32_t)((kb & 1) * 128)), d2 = d1 + (heavy ? (uint32_t)COL_CORR : (uint32_t)NP);
                const uint32_t i1 = heavy ? idesc_192 : idesc_64, i2 = heavy ? idesc_96 : idesc_32;
                const uint32_t first = (uint32_t)(heavy ? kb : (kb >> 1));
                const uint64_t a_hi0 = make_desc(ring0 + (uint32_t)(stage * TILE_BYTES));
                const uint64_t a_lo0 = make_desc(ring0 + (uint32_t)(stage * TILE_BYTES + TILE_BYTES / 2));
                const uint64_t b_0 = make_desc64(bop + (uint32_t)(2 * kb * KB_BYTES));
                if (leader) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t boff = (uint64_t)(((ks >> 1) * KB_BYTES + (ks & 1) * 32) >> 4);
                    mma_f16(d1, a_hi0 + (uint64_t)(2 * ks), b_0 + boff, i1, (first | (uint32_t)ks) != 0u);
                    mma_f16(d2, a_lo0 + (uint64_t)(2 * ks), b_0 + boff, i2, 1u);
                  }
                  mma_commit_mc(&empty[stage], (uint16_t)3);   // frees the slot in both CTAs once these MMAs have read it
                  if (mt == 1 && kb == 1) mma_commit(kfree);   // the m-tile-0 epilogue may overwrite k-blocks 0-1
                }
                __syncwarp();
                if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
              }
              if (leader) mma_commit(acc_full);
              __syncwarp();
            }
          }
        }
    }
  } else if (warp == NEW + 2) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ===== integrator warp: one particle per lane (see k_tc16)
    const float* red = misc + Smem::f_red;
    float* sxs = misc + Smem::f_xs;
    const float third = (float)(1.0 / 3.0);
    float sy[2] = {0.f, 0.f}, sk1[2] = {0.f, 0.f}, sk2[2] = {0.f, 0.f}, sk3[2] = {0.f, 0.f};
    const int myp = lane;
    // red[] holds the totals of particle myp in warps 4*(myp/8) .. +3, lane 16 r + 8 d + c4 with r = (myp%8)/4, c4 = myp%4
    const int red_w = (myp >> 3) * 4, red_l = ((myp & 7) >> 2) * 16 + (myp & 3);
    for (long it = 0; it < n_iter; ++it) {
      const long tile = blockIdx.x + it * gridDim.x;
      const int gs = set_of(it);
      const long base = (P.mode == 2 ? (tile < n_tiles ? tile - (long)gs * tiles_per_set : tiles_per_set) : tile) * NP;
      float* const out_set = P.out + (P.mode == 2 ? (size_t)gs * P.N * 2 : (size_t)0);
      const long row = base + myp;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const float v = row < P.N ? P.x_in[row * 2 + d] : 0.0f;
        sy[d] = v;
        sxs[2 * myp + d] = v;
      }
      ebar544();
#pragma unroll 1
      for (int ev = 0; ev < n_eval; ++ev) {
        const Stage st = stage_of(ev, gs);
        const bool heavy = heavy_of(st.kind, gs);
        ebar544();   // x1hat partial sums are in red[]
        float pr_tot[2], gg[2] = {0.f, 0.f};
#pragma unroll
        for (int d = 0; d < 2; ++d) {   // fixed summation order over the 4 lane-quarter warps of the particle's group
          float acc = __ldg(P.net.b3 + d);
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) acc += red[(red_w + w4) * 32 + red_l + 8 * d];
          pr_tot[d] = acc;
        }
        if (heavy) {
          ebar544();   // gradient partial sums are in red[]
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            float acc = 0.0f;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) acc += red[(red_w + w4) * 32 + red_l + 8 * d];
            gg[d] = acc;
          }
        }
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const float diff = __fsub_rn(pr_tot[d], sxs[2 * myp + d]);
          float f;
          if (st.kind == 0) f = diff;
          else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
          else {
            f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
            if (heavy) f = __fadd_rn(f, gg[d]);
          }
          if (row < P.N && !(fabsf(f) <= 3.0e38f)) atomicAdd(P.counters, 1ull);
          if (P.mode == 0) {
            if (row < P.N) out_set[row * 2 + d] = f;
          } else {
            const int iv = ev >> 2, sg = ev & 3;
            const float dt = P.mode == 2 ? P.tabs[gs].dt[iv] : rk.dt[iv], y = sy[d];
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
              if (ev == n_eval - 1 && row < P.N) out_set[row * 2 + d] = xn;
            }
            sxs[2 * myp + d] = xn;
          }
        }
        ebar544();   // the next stage's input is in sxs
      }
    }
  } else if (warp == NEW + 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");   // spare warp: only returns its registers to the pool
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ===== epilogue warps: thread <-> 4 consecutive hidden units x 2 particles x 3 jets
    const int q = warp & 3, pq = warp >> 2;
    const int p0 = pq * 8;
    const int g = lane >> 2, c4 = lane & 3;
    const int slot = (g & 1) | (((g >> 1) & 1) << 2) | ((g >> 2) << 1);   // 8-byte slot of the thread's 4 units (see k_tc16)
    float* red = misc + Smem::f_red;       // [16 warps][32]
    float* sxs = misc + Smem::f_xs;        // stage input of the 32 particles, [p][2]
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + p0;   // + (16 << 16) for the upper 16 lanes
    uint32_t aphase = 0, kphase = 0;
    // this thread's 8-byte group inside the operand rows of its particles p0 + 4r + c4 (r = 0, 1): operand k-block mt*4 + q,
    // 16-byte chunk (slot >> 1) ^ ((row >> 1) & 3) [SWIZZLE_64B], half slot & 1; (row >> 1) & 3 = (r << 1) | (c4 >> 1)
    uint32_t sa0[2];
    {
      const uint32_t b0 = s32(smem + Smem::act) + (uint32_t)(q * KB_BYTES + (p0 + c4) * 64 + (slot & 1) * 8);
      const uint32_t ce = (uint32_t)((slot >> 1) ^ (c4 >> 1));
      sa0[0] = b0 + (ce << 4);
      sa0[1] = b0 + ((ce ^ 2u) << 4);
    }
    auto make_sa = [&](int mt, uint32_t (&sa)[2]) {
      sa[0] = sa0[0] + (uint32_t)(mt * 4 * KB_BYTES);
      sa[1] = sa0[1] + (uint32_t)(mt * 4 * KB_BYTES);
    };
    auto publish = [&](int mt) {
      fence_proxy_async();
      mbar_arrive(&b_ready[mt]);
    };
    // A thread's 8 values of one jet are 4 packed pairs: pair i = 2*ld + r holds units (2 ld, 2 ld + 1) of its 4 consecutive
    // units for particle r of its 2 (tcgen05.ld.16x128b.x2: register 2r + h <-> lane g + 8h, column 4r + c4).
    // z = main + 2^-11 * corr, jet j
    auto load_acc = [&](int j, uint64_t (&z)[4]) {
      uint32_t m[HP], c[HP];
      tmem_ld4(lane_base + j * NP, m);
      tmem_ld4(lane_base + (16u << 16) + j * NP, m + 4);
      tmem_ld4(lane_base + COL_CORR + j * NP, c);
      tmem_ld4(lane_base + (16u << 16) + COL_CORR + j * NP, c + 4);
      tmem_wait_ld();
      const uint64_t inv2 = bc2(LO_INV);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        z[i] = fma2(pk2(__uint_as_float(c[2 * i]), __uint_as_float(c[2 * i + 1])), inv2,
                    pk2(__uint_as_float(m[2 * i]), __uint_as_float(m[2 * i + 1])));
    };
    // boundary evaluations: jet 0 summed over the two accumulator sets [main 32 | corr 32] at columns 0 and 128
    auto load_acc2 = [&](uint64_t (&z)[4]) {
      uint32_t m0[HP], c0[HP], m1[HP], c1[HP];
      tmem_ld4(lane_base, m0);
      tmem_ld4(lane_base + (16u << 16), m0 + 4);
      tmem_ld4(lane_base + NP, c0);
      tmem_ld4(lane_base + (16u << 16) + NP, c0 + 4);
      tmem_ld4(lane_base + 128, m1);
      tmem_ld4(lane_base + (16u << 16) + 128, m1 + 4);
      tmem_ld4(lane_base + 128 + NP, c1);
      tmem_ld4(lane_base + (16u << 16) + 128 + NP, c1 + 4);
      tmem_wait_ld();
      const uint64_t inv2 = bc2(LO_INV);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        z[i] = fma2(add2(pk2(__uint_as_float(c0[2 * i]), __uint_as_float(c0[2 * i + 1])),
                         pk2(__uint_as_float(c1[2 * i]), __uint_as_float(c1[2 * i + 1]))),
                    inv2,
                    add2(pk2(__uint_as_float(m0[2 * i]), __uint_as_float(m0[2 * i + 1])),
                         pk2(__uint_as_float(m1[2 * i]), __uint_as_float(m1[2 * i + 1]))));
    };
    // hi / lo' fp16 words of the 4 pairs: hw[2r + ld] = units (2 ld, 2 ld + 1) of particle r
    auto pack_pairs = [&](const uint64_t (&v)[4], uint32_t (&hw)[4], uint32_t (&lw)[4]) {
#pragma unroll
      for (int ld = 0; ld < 2; ++ld)
#pragma unroll
        for (int r = 0; r < 2; ++r) split2p(v[2 * ld + r], hw[2 * r + ld], lw[2 * r + ld]);
    };
    // pack and store one jet of an interior evaluation
    auto put_jet = [&](int jet, const uint32_t (&sa)[2], const uint64_t (&v)[4]) {
      uint32_t hw[4], lw[4];
      pack_pairs(v, hw, lw);
      if (jet == 0) store_jet<0, LO_ROWS_BYTES>(sa, hw, lw);
      else if (jet == 1) store_jet<1, LO_ROWS_BYTES>(sa, hw, lw);
      else store_jet<2, LO_ROWS_BYTES>(sa, hw, lw);
    };
    auto put_light = [&](const uint32_t (&sa)[2], const uint64_t (&v)[4]) {   // boundary: X_lo' rows directly behind X_hi
      uint32_t hw[4], lw[4];
      pack_pairs(v, hw, lw);
      store_jet<0, JET_BYTES>(sa, hw, lw);
    };
    // raw bits of 4 pairs -> one tcgen05.st.x2 per lane half
    auto stash_put = [&](uint32_t addr, const uint64_t (&v)[4]) {
      uint32_t u[HP];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        upk2(v[i], lo, hi);
        u[2 * i] = __float_as_uint(lo);
        u[2 * i + 1] = __float_as_uint(hi);
      }
      tmem_st4(addr, u);
      tmem_st4(addr + (16u << 16), u + 4);
    };
    auto stash_get = [&](uint32_t addr, uint32_t (&u)[HP]) {
      tmem_ld4(addr, u);
      tmem_ld4(addr + (16u << 16), u + 4);
    };
    auto wait_acc = [&]() {
      mbar_wait(acc_full, aphase);
      aphase ^= 1u;
      tc_fence_after();
    };
    auto wait_kfree = [&]() {
      mbar_wait(kfree, kphase);
      kphase ^= 1u;
    };
    auto release_acc = [&]() {
      tc_fence_before();
      mbar_arrive(acc_free);
    };
    // sum over the 8 threads that share the thread's particles: v[2r + d] -> lane g keeps index 2 (g >> 2) + ((g >> 1) & 1)
    auto reduce_units = [&](float (&v)[4]) -> float {
      {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = up ? v[i] : v[i + 2];
          const float keep = up ? v[i + 2] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool up = (lane & 8) != 0;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
      return v[0];
    };
    auto fold_pairs = [&](const uint64_t (&px)[2], const uint64_t (&py)[2], float (&v)[4]) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float lo, hi;
        upk2(px[r], lo, hi);
        v[2 * r] = lo + hi;
        upk2(py[r], lo, hi);
        v[2 * r + 1] = lo + hi;
      }
    };
    // layer 1 of the thread's 4 units x 2 particles: z1 = W0 [x; t] + b0 (K = 3, computed directly), per pair
    auto layer1_pre = [&](int n0, float tcur, uint64_t (&z1)[4]) {
      const float4 wx0 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 0 * W + n0));
      const float4 wx1 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 1 * W + n0));
      const float4 wtt = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 2 * W + n0));
      const float4 bb0 = __ldg(reinterpret_cast<const float4*>(P.net.b0 + (size_t)P.tidx * W + n0));
      const uint64_t t2 = bc2(tcur);
      const uint64_t zb[2] = {fma2(pk2(wtt.x, wtt.y), t2, pk2(bb0.x, bb0.y)), fma2(pk2(wtt.z, wtt.w), t2, pk2(bb0.z, bb0.w))};
      const uint64_t w0p[2] = {pk2(wx0.x, wx0.y), pk2(wx0.z, wx0.w)}, w1p[2] = {pk2(wx1.x, wx1.y), pk2(wx1.z, wx1.w)};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float2 xp = *reinterpret_cast<const float2*>(sxs + 2 * (p0 + 4 * r + c4));
        const uint64_t xx = bc2(xp.x), xy = bc2(xp.y);
#pragma unroll
        for (int ld = 0; ld < 2; ++ld) z1[2 * ld + r] = fma2(w1p[ld], xy, fma2(w0p[ld], xx, zb[ld]));
      }
    };

    mbar_arrive(acc_free);   // nothing to read out before the very first m-tile
    for (long it = 0; it < n_iter; ++it) {
      const int gs = set_of(it);
      ebar544();   // the integrator warp has written the tile's start points to sxs
#pragma unroll 1
      for (int ev = 0; ev < n_eval; ++ev) {
        const Stage st = stage_of(ev, gs);
        const bool heavy = heavy_of(st.kind, gs);
        // ---- layer 1: a1, a1' = s'(z1) z1', a1'' = E z1'^2   (z1' = W0 d per unit, z1'' = 0)
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
          const int n0 = mt * 128 + q * 32 + 4 * slot;
          uint64_t v0[4];
          layer1_pre(n0, st.t, v0);
          uint32_t sa[2];
          make_sa(mt, sa);
          if (heavy) {
            const float4 zz1 = __ldg(reinterpret_cast<const float4*>(P.net.zd1 + n0));
            const uint64_t zz1p[2] = {pk2(zz1.x, zz1.y), pk2(zz1.z, zz1.w)};
            uint64_t v1[4], v2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1, E;
              selu_jet2(v0[i], v0[i], s1, E);
              v1[i] = mul2(s1, zz1p[i >> 1]);
              v2[i] = mul2(mul2(E, zz1p[i >> 1]), zz1p[i >> 1]);
            }
            put_jet(0, sa, v0);
            put_jet(1, sa, v1);
            put_jet(2, sa, v2);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1, E;
              selu_jet2(v0[i], v0[i], s1, E);
            }
            put_light(sa, v0);
          }
          publish(mt);
        }
        // ---- gemm 0: layer 2 forward
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
          const int n0 = mt * 128 + q * 32 + 4 * slot;
          wait_acc();
          uint64_t z[4], zd[4], zdd[4];
          if (heavy) {
            load_acc(0, z);
            load_acc(1, zd);
            load_acc(2, zdd);
          } else {
            load_acc2(z);
          }
          release_acc();
          const float4 bb1 = __ldg(reinterpret_cast<const float4*>(P.net.b1 + n0));
          const uint64_t bb1p[2] = {pk2(bb1.x, bb1.y), pk2(bb1.z, bb1.w)};
          const uint32_t stash = lane_base + COL_STASH + mt * STASH_MT;
          if (heavy) {   // stash z2', z2'' (raw) and the derivative code of z2
            stash_put(stash + NP, zd);
            stash_put(stash + 2 * NP, zdd);
            uint32_t cu[HP];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1;
              float c0, c1;
              selu_jet_code2(add2(z[i], bb1p[i >> 1]), z[i], s1, c0, c1);
              cu[2 * i] = __float_as_uint(c0);
              cu[2 * i + 1] = __float_as_uint(c1);
              // E = code on the exp branch, 0 on the linear one
              const uint64_t E = pk2(c0 < 0.0f ? 0.0f : c0, c1 < 0.0f ? 0.0f : c1);
              zdd[i] = fma2(mul2(E, zd[i]), zd[i], mul2(s1, zdd[i]));   // a'' = E z'^2 + s' z''
              zd[i] = mul2(s1, zd[i]);                                  // a'  = s' z'
            }
            tmem_st4(stash, cu);
            tmem_st4(stash + (16u << 16), cu + 4);
            tmem_wait_st();
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t s1, E;
              selu_jet2(add2(z[i], bb1p[i >> 1]), z[i], s1, E);
            }
          }
          if (mt == 0) wait_kfree();   // k-blocks 0-1 have been read by every MMA of this pass
          uint32_t sa[2];
          make_sa(mt, sa);
          if (heavy) {
            put_jet(0, sa, z);
            put_jet(1, sa, zd);
            put_jet(2, sa, zdd);
          } else {
            put_light(sa, z);
          }
          publish(mt);
        }
        // ---- gemm 1: layer 3 forward; x1hat partial sums; adjoint seeds
        {
          float rsum = 0.0f;
#pragma unroll 1
          for (int mt = 0; mt < 2; ++mt) {
            const int n0 = mt * 128 + q * 32 + 4 * slot;
            wait_acc();
            uint64_t z[4], zd[4], zdd[4];
            if (heavy) {
              load_acc(0, z);
              load_acc(1, zd);
              load_acc(2, zdd);
            } else {
              load_acc2(z);
            }
            release_acc();
            const float4 bb2 = __ldg(reinterpret_cast<const float4*>(P.net.b2 + n0));
            const float4 uu3 = __ldg(reinterpret_cast<const float4*>(P.net.u3 + n0));
            const float4 w30 = __ldg(reinterpret_cast<const float4*>(P.net.w3 + 0 * W + n0));
            const float4 w31 = __ldg(reinterpret_cast<const float4*>(P.net.w3 + 1 * W + n0));
            const uint64_t bb2p[2] = {pk2(bb2.x, bb2.y), pk2(bb2.z, bb2.w)};
            const uint64_t w30p[2] = {pk2(w30.x, w30.y), pk2(w30.z, w30.w)}, w31p[2] = {pk2(w31.x, w31.y), pk2(w31.z, w31.w)};
            const uint64_t u3p[2] = {pk2(uu3.x, uu3.y), pk2(uu3.z, uu3.w)};
            uint64_t px[2] = {0ull, 0ull}, py[2] = {0ull, 0ull};
#pragma unroll
            for (int ld = 0; ld < 2; ++ld) {
              // adjoint seeds of a3, a3', a3'': A_i = c_i * (W3^T 1)
              const uint64_t A0 = mul2(bc2(st.c[0]), u3p[ld]), A1 = mul2(bc2(st.c[1]), u3p[ld]), A2 = mul2(bc2(st.c[2]), u3p[ld]);
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const int i = 2 * ld + r;
                uint64_t a, s1, E;
                selu_jet2(add2(z[i], bb2p[ld]), a, s1, E);
                px[r] = fma2(a, w30p[ld], px[r]);
                py[r] = fma2(a, w31p[ld], py[r]);
                if (heavy) {   // (P, Q, R) of layer 3, in place
                  const uint64_t A2E = mul2(A2, E);
                  z[i] = fma2(A2E, fma2(zd[i], zd[i], zdd[i]), fma2(mul2(A1, E), zd[i], mul2(A0, s1)));
                  zd[i] = fma2(add2(A2E, A2E), zd[i], mul2(A1, s1));
                  zdd[i] = mul2(A2, s1);
                }
              }
            }
            float part[4];
            fold_pairs(px, py, part);
            rsum += reduce_units(part);
            if (heavy) {
              if (mt == 0) wait_kfree();
              uint32_t sa[2];
              make_sa(mt, sa);
              put_jet(0, sa, z);
              put_jet(1, sa, zd);
              put_jet(2, sa, zdd);
              publish(mt);
            } else if (mt == 0) {
              wait_kfree();
            }
          }
          red[warp * 32 + lane] = rsum;
        }
        ebar544();   // x1hat partial sums are in red[]: the integrator warp adds them up
        if (heavy) {
          // ---- gemm 2: reverse, layer 3 -> 2: accumulators = adjoints (A0, A1, A2) of (a2, a2', a2'')
#pragma unroll 1
          for (int mt = 0; mt < 2; ++mt) {
            wait_acc();
            uint64_t a0[4], a1[4], a2[4];
            load_acc(0, a0);
            load_acc(1, a1);
            load_acc(2, a2);
            release_acc();
            {
              const uint32_t stash = lane_base + COL_STASH + mt * STASH_MT;
              uint32_t cd[HP], zu[HP], zzu[HP];
              stash_get(stash, cd);
              stash_get(stash + NP, zu);
              stash_get(stash + 2 * NP, zzu);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float c0 = __uint_as_float(cd[2 * i]), c1 = __uint_as_float(cd[2 * i + 1]);
                const uint64_t s1 = pk2(c0 < 0.0f ? kSeluLambda : c0, c1 < 0.0f ? kSeluLambda : c1);
                const uint64_t E = pk2(c0 < 0.0f ? 0.0f : c0, c1 < 0.0f ? 0.0f : c1);
                const uint64_t zd = pk2(__uint_as_float(zu[2 * i]), __uint_as_float(zu[2 * i + 1]));
                const uint64_t zdd = pk2(__uint_as_float(zzu[2 * i]), __uint_as_float(zzu[2 * i + 1]));
                const uint64_t A2E = mul2(a2[i], E);
                a0[i] = fma2(A2E, fma2(zd, zd, zdd), fma2(mul2(a1[i], E), zd, mul2(a0[i], s1)));
                a1[i] = fma2(add2(A2E, A2E), zd, mul2(a1[i], s1));
                a2[i] = mul2(a2[i], s1);
              }
            }
            if (mt == 0) wait_kfree();
            uint32_t sa[2];
            make_sa(mt, sa);
            put_jet(0, sa, a0);
            put_jet(1, sa, a1);
            put_jet(2, sa, a2);
            publish(mt);
          }
          // ---- gemm 3: reverse, layer 2 -> 1 -> x: P1 = A0 s' + A1 E z1' + A2 E z1'^2 (z1'' = 0); grad_x = W0^T P1
          float gsum = 0.0f;
#pragma unroll 1
          for (int mt = 0; mt < 2; ++mt) {
            const int n0 = mt * 128 + q * 32 + 4 * slot;
            wait_acc();
            uint64_t a0[4], a1[4], a2[4];
            load_acc(0, a0);
            load_acc(1, a1);
            load_acc(2, a2);
            release_acc();
            uint64_t z1[4];
            layer1_pre(n0, st.t, z1);   // layer 1 is recomputed from x (K = 3) rather than kept across the four passes
            const float4 wx0 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 0 * W + n0));
            const float4 wx1 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 1 * W + n0));
            const float4 zz1 = __ldg(reinterpret_cast<const float4*>(P.net.zd1 + n0));
            const uint64_t w0p[2] = {pk2(wx0.x, wx0.y), pk2(wx0.z, wx0.w)}, w1p[2] = {pk2(wx1.x, wx1.y), pk2(wx1.z, wx1.w)};
            const uint64_t zz1p[2] = {pk2(zz1.x, zz1.y), pk2(zz1.z, zz1.w)};
            uint64_t px[2] = {0ull, 0ull}, py[2] = {0ull, 0ull};
#pragma unroll
            for (int ld = 0; ld < 2; ++ld)
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const int i = 2 * ld + r;
                uint64_t a, s1, E;
                selu_jet2(z1[i], a, s1, E);
                const uint64_t Ez = mul2(E, zz1p[ld]);
                const uint64_t P1 = fma2(mul2(a2[i], Ez), zz1p[ld], fma2(a1[i], Ez, mul2(a0[i], s1)));
                px[r] = fma2(P1, w0p[ld], px[r]);
                py[r] = fma2(P1, w1p[ld], py[r]);
              }
            float part[4];
            fold_pairs(px, py, part);
            gsum += reduce_units(part);
            if (mt == 0) wait_kfree();   // keeps the barrier phase in step
          }
          red[warp * 32 + lane] = gsum;
          ebar544();   // gradient partial sums are in red[]
        }
        ebar544();   // the integrator warp has written the next stage's input to sxs
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be multicasting into this CTA's ring / arriving on its barriers
  if (warp == NEW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace tc3

// the fused three-jet kernel serves the order-3 momentum field on the 3-256-256-256-2 network
bool tensor16o3_supports(const idff_weights_t* w, const idff_field_params_t* p) {
  const NetView& v = w->v;
  if (v.w != 256 || v.tc16_tiles == nullptr || p->dim != 2 || v.din != 3 || v.dout != 2) return false;
  if (p->variant != IDFF_FIELD_MOMENTUM) return false;
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  return nj == 3;
}

// ... and can run any momentum field of that network: a jet whose coefficient is zero contributes nothing to the reverse sweep.
// idff_sample_rk4 uses this for SMALL order-1 / order-2 solves: 32 particles per CTA put 2048 particles on 64 SMs where the
// 64-particle tiles of k_tc16 reach 32 (measured: see idff_api.cu).
bool tensor16o3_can_run(const idff_weights_t* w, const idff_field_params_t* p) {
  const NetView& v = w->v;
  return v.w == 256 && v.tc16_tiles != nullptr && p->dim == 2 && v.din == 3 && v.dout == 2 && p->variant == IDFF_FIELD_MOMENTUM;
}

static int verify_launch(idff_weights_t* w, void* stream) {
  if (w->tc_verified & 8) return IDFF_OK;
  int flag = 0;
  IDFF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  IDFF_CUDA(cudaMemcpyFromSymbol(&flag, g_tc16o3_misaligned, sizeof(int)));
  if (flag) {
    set_error("tcgen05 kernel: dynamic shared memory is not 1024-byte aligned on this driver; use impl = IDFF_IMPL_FUSED");
    return IDFF_EUNSUPPORTED;
  }
  w->tc_verified |= 8;
  return IDFF_OK;
}

static int launch_o3(const idff_weights_t* w, const idff_field_params_t* p, tc3::TParams& P, const Rk4Table* rk, const Stage* one,
                     void* stream) {
  if (!tensor16o3_can_run(w, p)) {
    set_error("3xFP16 three-jet sampler: needs a 3-256-256-256-2 network, dim 2, the momentum field");
    return IDFF_EUNSUPPORTED;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long tiles = P.mode == 2 ? P.tiles_per_set * P.n_sets : (P.N + tc3::NP - 1) / tc3::NP;
  int grid = (int)(tiles < sms ? tiles : sms);
  grid = (grid + 1) & ~1;   // CTA pairs
  static thread_local Rk4Table rk0;
  static thread_local Stage one0;
  const Rk4Table& rkr = rk ? *rk : rk0;
  const Stage& oner = one ? *one : one0;
  const int smem = tc3::Smem::total;
  IDFF_CUDA(cudaFuncSetAttribute(tc3::k_tc16o3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc3::NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IDFF_CUDA(cudaLaunchKernelEx(&cfg, tc3::k_tc16o3, P, rkr, oner));
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return verify_launch(const_cast<idff_weights_t*>(w), stream);
}

static void fill_params(const idff_weights_t* w, const idff_field_params_t* p, int mode, int64_t N, tc3::TParams* P) {
  memset(P, 0, sizeof(*P));
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  P->net = w->v;
  P->variant = p->variant; P->D = p->dim; P->tidx = p->t_idx; P->reverse = rev; P->mode = mode;
  P->N = (long)N;
  P->counters = w->d_counters;
}

int tensor16o3_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, float* d_out,
                          int64_t N, bool safe, void* stream) {
  tc3::TParams P;
  fill_params(w, p, 0, N, &P);
  if (safe) {
    int rc = keep_input(const_cast<idff_weights_t*>(w), &d_x, d_out, false, (size_t)N * 2 * sizeof(float), stream);
    if (rc) return rc;
  }
  P.x_in = d_x; P.out = d_out;
  Stage st;
  make_stage(*p, t, &st);
  int rc = launch_o3(w, p, P, nullptr, &st, stream);
  if (rc || !safe) return rc;
  static thread_local Rk4Table tab;
  tab.n_iv = 0;
  tab.st[0] = st;
  return fixup_nonfinite(w, p, 0, 1, 0, &tab, nullptr, nullptr, nullptr, d_x, nullptr, d_out, nullptr, N, stream);
}

// One launch for n_sets parameter sets (all with three jets) over the same x0 and grid: the 10^3-tuple sweep of the order-3
// script (Autograd_example_order3.py:288-314).
int tensor16o3_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, const float* d_x0,
                                const Stage* stages, const float* dts, int n_iv, float* d_out, int64_t N, bool safe, void* stream) {
  if (n_iv > kMaxSweepIntervals) {
    set_error("3xFP16 sweep: %d intervals exceed the per-launch table (%d)", n_iv, kMaxSweepIntervals);
    return IDFF_EUNSUPPORTED;
  }
  idff_weights_t* wm = const_cast<idff_weights_t*>(w);
  const size_t bytes = sizeof(SweepTab) * (size_t)n_sets;
  if (wm->sweep_bytes < bytes) {
    if (wm->d_sweep) cudaFree(wm->d_sweep);
    wm->d_sweep = nullptr;
    wm->sweep_bytes = 0;
    IDFF_CUDA(cudaMalloc(&wm->d_sweep, bytes));
    wm->sweep_bytes = bytes;
  }
  static thread_local std::vector<SweepTab> host;
  host.assign((size_t)n_sets, SweepTab());
  for (int g = 0; g < n_sets; ++g) {
    if (!tensor16o3_supports(w, &params[g])) {
      set_error("3xFP16 three-jet sweep: parameter set %d is not served by this kernel", g);
      return IDFF_EUNSUPPORTED;
    }
    int nj, rev;
    jets_needed(params[g], &nj, &rev);
    memset(&host[g], 0, sizeof(SweepTab));
    host[g].reverse = rev;
    memcpy(host[g].dt, dts, sizeof(float) * n_iv);
    memcpy(host[g].st, stages + (size_t)g * 4 * n_iv, sizeof(Stage) * 4 * n_iv);
  }
  // pageable source: the runtime stages the copy before returning, `host` may be reused by the next call
  IDFF_CUDA(cudaMemcpyAsync(wm->d_sweep, host.data(), bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  tc3::TParams P;
  fill_params(w, &params[0], 2, N, &P);
  P.n_sets = n_sets;
  P.sweep_n_iv = n_iv;
  const long tiles = ((long)N + tc3::NP - 1) / tc3::NP;
  P.tiles_per_set = (tiles + 1) & ~1L;
  P.tabs = reinterpret_cast<const SweepTab*>(wm->d_sweep);
  P.x_in = d_x0; P.out = d_out;
  int rc = launch_o3(w, &params[0], P, nullptr, nullptr, stream);
  if (rc || !safe) return rc;
  return fixup_nonfinite(w, &params[0], 2, n_sets, n_iv, nullptr, reinterpret_cast<const SweepTab*>(wm->d_sweep), nullptr, nullptr,
                         d_x0, nullptr, d_out, nullptr, N, stream);
}

int tensor16o3_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                          const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, bool safe, void* stream) {
  static thread_local Rk4Table tab;
  const float* src = d_x0;
  for (int done = 0; done < n_iv; done += kMaxIntervals) {
    const int cnt = n_iv - done < kMaxIntervals ? n_iv - done : kMaxIntervals;
    tab.n_iv = cnt;
    memcpy(tab.dt, dts + done, sizeof(float) * cnt);
    memcpy(tab.st, stages + 4 * (size_t)done, sizeof(Stage) * 4 * cnt);
    tc3::TParams P;
    fill_params(w, p, 1, N, &P);
    if (safe) {
      int rc = keep_input(const_cast<idff_weights_t*>(w), &src, d_out_final, done > 0, (size_t)N * 2 * sizeof(float), stream);
      if (rc) return rc;
    }
    P.x_in = src; P.out = d_out_final;
    P.traj = d_out_traj ? d_out_traj + (size_t)done * N * p->dim : nullptr;
    int rc = launch_o3(w, p, P, &tab, nullptr, stream);
    if (rc) return rc;
    if (safe) {
      rc = fixup_nonfinite(w, p, 1, 1, cnt, &tab, nullptr, nullptr, nullptr, src, nullptr, d_out_final,
                           P.traj ? P.traj + (size_t)N * 2 : nullptr, N, stream);
      if (rc) return rc;
    }
    src = d_out_final;
  }
  return IDFF_OK;
}

}  // namespace idff
