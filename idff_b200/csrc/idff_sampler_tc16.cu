// Tensor-core IDFF sampler, 3xFP16 split (IDFF_IMPL_TENSOR16): hidden width 256, state dimension 2, up to two jets.
//
// Every w x w layer pass of the evaluation (forward jets and reverse sweep) runs on tcgen05.mma kind::f16 with fp32
// accumulation in tensor memory.  An fp32 value v is carried as two fp16 numbers
//     hi = fp16_rn(v),   lo' = fp16_rn((v - hi) * 2^11)
// (the remainder is scaled back into the normal fp16 range, so 22 significant bits survive even for tiny v), and
//     D = W_hi*X_hi  +  2^-11 * (W_hi*X_lo' + W_lo'*X_hi).
// Measured on the part (profiles/microbench/tc_f16.cu): error 1.4e-7 .. 8.9e-7 of sum|terms| over K = 256 -- the level
// of an fp32 FMA chain, better than 3xTF32 (1e-6) -- at twice the MACs per MMA instruction of kind::tf32.  fp16 has no
// headroom above 65504: an activation, jet or adjoint beyond that turns into inf/NaN in that particle's own columns
// only (columns never mix), so the particle's output is NaN -- loud, never silently wrong; such networks use
// IDFF_IMPL_TENSOR (3xTF32) or IDFF_IMPL_FUSED (fp32 FMA).
//
// Formulation ("units on M"): D[unit][col] = sum_k W[unit][k] * X[col][k].
//   * B operand = activations of the CTA's 64 particles, resident in shared memory (128 KB): per 32-wide k-block a
//     256-row K-major SWIZZLE_64B tile (64-byte rows), row = hl*128 + jet*64 + particle (hl 0 = hi, 1 = lo');
//     rewritten in place by the epilogue of every layer pass with conflict-free 64-bit stores (4 units each).
//   * A operand = weights: 32 KB tile images [128 units x 64 k] hi | lo', pre-swizzled at handle creation, streamed by
//     a TMA producer warp (cp.async.bulk multicast inside a CTA pair, each CTA loads one half), 3-stage ring.
//   * per 16-wide k-step two MMAs, both at a shape the tensor core runs at full rate [measured: 128xNx16 costs 72.6
//     cycles for any N <= 128 and 128.8 for N = 256]:
//         [main | corr] += W_hi * [X_hi ; X_lo']      N = 256
//                corr   += W_lo' * X_hi               N = 128
//   * TMEM (512 columns): accumulators [main 128 | corr 128] shared by the two m-tiles, which run back to back, and
//     the layer-2 derivative stash of the reverse sweep (2 x 128 columns).  Boundary evaluations (t = 0, t = 1: one jet,
//     no reverse sweep) keep [X_hi ; X_lo'] in operand rows 0..127, run N = 128 / N = 64 MMAs and split the K sum over
//     two accumulator sets by k-block parity: half as many truncating accumulation steps where t = 1 amplifies x100.
// A TMEM lane is one hidden unit.  The epilogue reads the accumulators with tcgen05.ld.16x128b: thread T of a warp gets
// lanes {g, g+8, g+16, g+24} (g = T/4) of its lane quarter and columns {T%4 + 4r}; the weight-tile rows are permuted at
// handle creation so that these four lanes are four CONSECUTIVE units 4s..4s+3 (s = slot(g), a fixed permutation of
// 0..7 chosen for bank-conflict-free stores), i.e. one 8-byte group of the next pass's K-major operand row.  A thread therefore owns every jet of 4 units x 4 particles: the SELU jet / adjoint
// algebra is thread-local, and sums over units are 3 in-thread adds + a 3-step shuffle.  16 epilogue warps = 4 lane
// quarters x 4 particle quarters; every thread serves m-tile 0 and then m-tile 1 of a pass.  The m-tile-0 epilogue computes while the m-tile-1 MMAs run, and the
// next pass starts on k-blocks 0-1 (written by the m-tile-0 epilogue) while the m-tile-1 epilogue is still at work.
#include <cstdlib>
#include <vector>

#include "idff_tcgen05.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);
__device__ int g_tc16_misaligned = 0;

namespace tch {

constexpr int W = 256;
constexpr int NP = 64;                    // particles per CTA
constexpr int NROW = 256;                 // rows of the B operand: hl*128 + jet*64 + particle
constexpr int NSTAGE = 3;                 // weight ring depth
constexpr int TILE_BYTES = 32768;         // hi (16 KB) + lo' (16 KB) image of one [128 units x 64 k] fp16 weight tile
constexpr int KB_BYTES = NROW * 64;       // B operand of one 32-wide k-block: 256 rows x 64 B (16 KB)
constexpr int NKB = W / 64;               // 4 weight k-blocks of 64 (= 8 operand k-blocks of 32)
constexpr int ACT_BYTES = (W / 32) * KB_BYTES;
constexpr int NE = 512;                   // epilogue threads: warp w -> lane quarter w & 3, particle quarter w >> 2
constexpr int NEW = NE / 32;
constexpr int NTHREADS = NE + 128;        // + producer warp 16 + MMA warp 17 + integrator warps 18, 19
constexpr int HP = 16;                    // particles per epilogue thread
constexpr int TMEM_COLS = 512;
constexpr int COL_CORR = 128, COL_STASH = 256;

using namespace tcg;

// One-jet passes leave TMEM columns 256..511 -- the stash of the reverse sweep -- unused: m-tile 1 can accumulate there with its
// own barrier pair, so the MMA warp does not wait for the read-out of m-tile 0 before it starts m-tile 1.  Used by the score-head
// instance (every pass is one-jet: +5..9 %, 987 -> 1079 M particle-steps/s at 2^21 particles).  For the boundary evaluations of the
// momentum instance the same switch measured SLOWER (415 vs 430 M particle-steps/s, profiles/r2_ab_variants.txt run 5): the extra
// branches raise the spills of the two-jet epilogue (40 -> 56 B) and that costs more than 10 % of the evaluations gain.
#ifdef IDFF_TC16_NO_LIGHT_DB
constexpr bool kLightDB = false;
#else
constexpr bool kLightDB = true;
#endif

// phase timestamps (clock64) of CTA 0, first tile, evaluation 1 (an interior one): diagnostic only, compiled in with
// -DIDFF_TC16_STAMPS (make STAMPS=1); all zero in the production build
__device__ long long g_tc16_stamps[32];

struct TParams {
  NetView net;
  int variant, D, tidx, reverse, mode;   // mode 0 eval, 1 rk4, 2 rk4 sweep over n_sets parameter sets (same x0, same grid)
  int n_sets, sweep_n_iv;
  long tiles_per_set;                    // sweep: even, so both CTAs of a pair always work on the same set
  const SweepTab* tabs;                  // sweep: [n_sets] in global memory
  long N;
  const float* x_in;
  const float* x0_cfm;
  float* out;
  float* traj;
  unsigned long long* counters;          // per handle: [0] += particle-evaluations whose field value was not finite
};

struct Smem {   // byte offsets from the dynamic base, which must be 1024-byte aligned (checked at kernel entry)
  static constexpr int act = 0;
  static constexpr int ring = ACT_BYTES;
  static constexpr int misc = ring + NSTAGE * TILE_BYTES;   // floats: red [16 warps][32], xs [64][2]
  static constexpr int f_red = 0, f_xs = NEW * 32, f_end = NEW * 32 + 2 * NP;
  static constexpr int bars = misc + f_end * 4;
  static constexpr int total = bars + 128;
};
static_assert(Smem::total <= 232448, "shared memory budget");

// A thread's 16 values of one jet are indexed e = 8*ld + 2*r + h: unit 2*ld + h (of its 4 consecutive units), particle
// r (of its 4).  pack_jet: hi / lo' fp16 words of the unit pairs (0,1) and (2,3) for each particle: hw[2r + ld].
__device__ __forceinline__ void pack_jet(const float (&v)[HP], uint32_t (&hw)[8], uint32_t (&lw)[8]) {
#pragma unroll
  for (int ld = 0; ld < 2; ++ld)
#pragma unroll
    for (int r = 0; r < 4; ++r) split2(v[8 * ld + 2 * r], v[8 * ld + 2 * r + 1], hw[2 * r + ld], lw[2 * r + ld]);
}
// 8-byte stores of one jet (JET) into the thread's four particle rows: hi rows at +0, lo' rows 128 rows further.
// sa[0] / sa[1] = swizzled address of the thread's 8-byte group for even / odd r, first particle, jet 0, hi.
// LO = byte offset of the lo' rows: 128 rows (8192 B) in interior evaluations; 64 rows in boundary evaluations, which
// carry one jet and keep [X_hi ; X_lo'] in rows 0..127 so that their MMAs are N = 128 / N = 64.
template <int JET, int LO = 8192>
__device__ __forceinline__ void store_jet(const uint32_t (&sa)[2], const uint32_t (&hw)[8], const uint32_t (&lw)[8]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    // (r & 1) is a compile-time constant after unrolling
    if (r == 0) { st64<JET * 4096 + 0 * 256>(sa[0], hw[0], hw[1]); st64<LO + JET * 4096 + 0 * 256>(sa[0], lw[0], lw[1]); }
    if (r == 1) { st64<JET * 4096 + 1 * 256>(sa[1], hw[2], hw[3]); st64<LO + JET * 4096 + 1 * 256>(sa[1], lw[2], lw[3]); }
    if (r == 2) { st64<JET * 4096 + 2 * 256>(sa[0], hw[4], hw[5]); st64<LO + JET * 4096 + 2 * 256>(sa[0], lw[4], lw[5]); }
    if (r == 3) { st64<JET * 4096 + 3 * 256>(sa[1], hw[6], hw[7]); st64<LO + JET * 4096 + 3 * 256>(sa[1], lw[6], lw[7]); }
  }
}
// per-unit constant c[0..3] of the thread's four units for element e
#define IDFF_UC(c, e) ((((e) >> 3) * 2 + ((e) & 1)) == 0 ? (c).x : (((e) >> 3) * 2 + ((e) & 1)) == 1 ? (c).y : (((e) >> 3) * 2 + ((e) & 1)) == 2 ? (c).z : (c).w)

// SH = true: the score-head field (example_order1_with_prediction_head.py:64-96) on its 5-256-256-256-4 network -- input
// cat[x, x, t], outputs x1hat = out[:, :2] and the score st = out[:, 2:], no autograd terms: every evaluation is a one-jet
// forward pass (the boundary-evaluation machinery), with a 5-term first layer, four output sums and the closed form
// a*(x1hat - x)/(1 - t) + c0*st + c1.  The second pair of output sums travels through operand rows 128.. of k-block 0, which
// no MMA of a one-jet pass reads.
template <bool SH>
__global__ void __launch_bounds__(NTHREADS, 1)
k_tc16(const __grid_constant__ TParams P, const __grid_constant__ Rk4Table rk, const __grid_constant__ Stage one) {
  constexpr bool DB = SH && kLightDB;   // second accumulator set for m-tile 1 of one-jet passes: the score-head instance only (see DB)
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((s32(smem) & 1023u) != 0u) {   // SWIZZLE_128B operands need a 1024-byte aligned base: report and do nothing
    if (threadIdx.x == 0) g_tc16_misaligned = 1;
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
  // SH: the score-head instance never uses the stash columns, so m-tile 1 accumulates in TMEM columns 256..511 with its own
  // barrier pair: the MMA warp does not wait for the read-out of m-tile 0 before it starts m-tile 1 (nor for m-tile 1's before the
  // next pass)
  uint64_t* acc_full1 = acc_full + 5;
  uint64_t* acc_free1 = acc_full + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 7);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);
    }
    mbar_init(acc_full, 1);
    mbar_init(kfree, 1);
    mbar_init(acc_free, NE);
    mbar_init(acc_full1, 1);
    mbar_init(acc_free1, NE);
    mbar_init(&b_ready[0], NE);
    mbar_init(&b_ready[1], NE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NEW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // jet-1 rows are read by the MMAs of boundary evaluations too: start from finite values
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
  // parameter set of iteration `it` of this CTA (0 outside sweeps); surplus tiles run the last set on padding
  auto set_of = [&](long it) -> int {
    if (P.mode != 2) return 0;
    const long g = (blockIdx.x + it * gridDim.x) / tiles_per_set;
    return (int)(g < P.n_sets ? g : P.n_sets - 1);
  };
  // by value: `one` / `rk` live in the constant bank, sweep tables in global memory -- a reference to "one of the three" makes
  // every field access a generic-address load (LD.E, long scoreboard) in the middle of the epilogue
  auto stage_of = [&](int e, int g) -> Stage {
    if (P.mode == 0) return one;
    if (P.mode == 1) return rk.st[e];
    return P.tabs[g].st[e];
  };
  auto kind_of = [&](int e, int g) -> int { return P.mode == 0 ? one.kind : (P.mode == 1 ? rk.st[e].kind : P.tabs[g].st[e].kind); };
  auto heavy_of = [&](int kind, int g) { return !SH && kind == 2 && (P.mode == 2 ? P.tabs[g].reverse : P.reverse) != 0; };
  float* const red2 = reinterpret_cast<float*>(smem + Smem::act + KB_BYTES / 2);   // SH: [16 warps][32] sums of outputs 2, 3

  // Register budget: 20 warps x 96 registers at launch.  The four service warps (TMA, MMA, 2 x integrator) give back 64 each,
  // the sixteen epilogue warps take 16 each: their jet / adjoint algebra holds 2 x 16 accumulators + 32 packed operand words.
  // (Each setmaxnreg sits at the top of the branch it governs: ptxas budgets the code a setmaxnreg DOMINATES -- placed in
  // an if / else ahead of the role branches, every role was compiled under the smaller limit.  Warps 16..19 are one
  // warpgroup and execute the same instruction.)
  if (warp >= NEW) {
  if (warp == NEW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ===== TMA producer: weight tile images of every layer pass, in consumption order.  Whole warp in uniform control
    // flow (addresses in uniform registers), one elected lane issues; no back-off on the slot wait: the ring is only
    // three stages deep and a late refill stalls the tensor pipe
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
    // ===== MMA issuer.  The whole warp walks the schedule in uniform control flow, so descriptors, TMEM addresses and
    // barrier phases live in uniform registers and each tcgen05.mma costs a handful of instructions; one elected lane
    // issues.  (Issued from an `if (lane == 0)` region every UTCHMMA was preceded by an ELECT / 7 x R2UR.BROADCAST /
    // BRA.U.ANY loop and the issuing thread, not the tensor pipe, set the pace of a layer pass.)
    {
      // kind::f16: D fp32 (bit 4), A = B = fp16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc_256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bop = s32(smem + Smem::act);
      const uint32_t ring0 = s32(smem + Smem::ring);
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0, fphase = 0, fphase1 = 0, bphase0 = 0, bphase1 = 0;
      for (long it = 0; it < n_iter; ++it)
        for (int e = 0; e < n_eval; ++e) {
          const int gs = set_of(it);
          const bool heavy = heavy_of(kind_of(e, gs), gs);
          const int ng = heavy ? 4 : 2;
#ifdef IDFF_TC16_STAMPS
          const bool mstamp = blockIdx.x == 0 && it == 0 && e == 1 && lane == 0;
          if (mstamp)
            for (int i = 24; i < 29; ++i) g_tc16_stamps[i] = 0;
#else
          constexpr bool mstamp = false;   // phase stamps cost the issuing warp registers it does not have: diagnostic builds only
#endif
          for (int g = 0; g < ng; ++g) {
            for (int mt = 0; mt < 2; ++mt) {
              long long c0 = mstamp ? clock64() : 0;
              if (DB && !heavy && mt == 1) {
                mbar_wait(acc_free1, fphase1);   // m-tile 1's own accumulator set (previous one-jet pass) has been read out
                fphase1 ^= 1u;
              } else {
                mbar_wait(acc_free, fphase);   // the previous m-tile's accumulators have been read out
                fphase ^= 1u;
              }
              tc_fence_after();
              if (mstamp) g_tc16_stamps[24 + mt] += clock64() - c0;
              for (int kb = 0; kb < NKB; ++kb) {
                if (mt == 0 && kb == 0) {
                  c0 = mstamp ? clock64() : 0;
                  mbar_wait(&b_ready[0], bphase0);
                  bphase0 ^= 1u;
                  tc_fence_after();
                  if (mstamp) g_tc16_stamps[16 + 2 * g] = clock64();
                  if (mstamp) g_tc16_stamps[26] += clock64() - c0;
                }
                if (mt == 0 && kb == 2) {
                  c0 = mstamp ? clock64() : 0;
                  mbar_wait(&b_ready[1], bphase1);
                  bphase1 ^= 1u;
                  tc_fence_after();
                  if (mstamp) g_tc16_stamps[27] += clock64() - c0;
                }
                c0 = mstamp ? clock64() : 0;
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (mstamp) g_tc16_stamps[28] += clock64() - c0;
                // interior evaluation: [main | corr] += W_hi * [X_hi ; X_lo'] (N = 256), corr += W_lo' * X_hi (N = 128).
                // boundary evaluation (one jet): rows 0..63 hold X_hi, rows 64..127 X_lo' (the jet-1 hi slot is free), so the
                // two MMAs are N = 128 and N = 64; accumulator set (kb & 1): [main 64 | corr 64] at columns 128 * (kb & 1)
                const uint32_t d1 = tmem + (heavy ? 0u : (uint32_t)((kb & 1) * 128)) + ((DB && !heavy && mt == 1) ? 256u : 0u),
                               d2 = d1 + (heavy ? 128u : 64u);
                const uint32_t i1 = heavy ? idesc_256 : idesc_128, i2 = heavy ? idesc_128 : idesc_64;
                const uint32_t first = (uint32_t)(heavy ? kb : (kb >> 1));
                // descriptors advance by 32 bytes (2 address units) per 16-wide k-step; the operand k-block (32 wide) is
                // 2*kb + ks/2, its 32-byte half ks%2
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
              if (leader) mma_commit((DB && !heavy && mt == 1) ? acc_full1 : acc_full);
              __syncwarp();
            }
            if (mstamp) g_tc16_stamps[17 + 2 * g] = clock64();
          }
        }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // ===== integrator warps (18, 19): one particle per lane.  x1hat / gradient totals from red[], the field's closed form
    // and the RK4 (3/8 rule) stage update in torchdiffeq's operation order; the state (y, k1..k3, x0) lives in this thread's
    // registers for the whole solve -- and no longer in those of the 512 epilogue threads, whose budget it strained.
    const float* red = misc + Smem::f_red;
    float* sxs = misc + Smem::f_xs;
    const float third = (float)(1.0 / 3.0);
    float sy[2] = {0.f, 0.f}, sk1[2] = {0.f, 0.f}, sk2[2] = {0.f, 0.f}, sk3[2] = {0.f, 0.f}, sx0[2] = {0.f, 0.f};
    const int myp = (warp - (NEW + 2)) * 32 + lane;
    // where red[] holds the totals of particle myp: warps 4*(myp/16) .. +3, lane (2*(c/4) + d)*4 + c%4 with c = myp%16
    const int red_w = (myp >> 4) * 4, red_l = ((myp & 15) >> 2) * 8 + (myp & 3);
    for (long it = 0; it < n_iter; ++it) {
      const long tile = blockIdx.x + it * gridDim.x;
      const int gs = set_of(it);
      // first particle of the tile; may lie beyond N: an all-padding tile
      const long base = (P.mode == 2 ? (tile < n_tiles ? tile - (long)gs * tiles_per_set : tiles_per_set) : tile) * NP;
      float* const out_set = P.out + (P.mode == 2 ? (size_t)gs * P.N * 2 : (size_t)0);
      const long row = base + myp;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const float v = row < P.N ? P.x_in[row * 2 + d] : 0.0f;
        sy[d] = v;
        sxs[2 * myp + d] = v;
        if (P.variant == IDFF_FIELD_CFM && P.x0_cfm) sx0[d] = row < P.N ? P.x0_cfm[row * 2 + d] : 0.0f;
      }
      ebar576();
#pragma unroll 1
      for (int ev = 0; ev < n_eval; ++ev) {
        const Stage st = stage_of(ev, gs);
        const bool heavy = heavy_of(st.kind, gs);
        ebar576();   // x1hat partial sums are in red[]
        float pr_tot[2], gg[2] = {0.f, 0.f};
#pragma unroll
        for (int d = 0; d < 2; ++d) {   // fixed summation order over the 4 lane-quarter warps of the particle's quarter
          float acc = __ldg(P.net.b3 + d);
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) acc += red[(red_w + w4) * 32 + red_l + 4 * d];
          pr_tot[d] = acc;
        }
        if (heavy) {
          ebar576();   // gradient partial sums are in red[]
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            float acc = 0.0f;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) acc += red[(red_w + w4) * 32 + red_l + 4 * d];
            gg[d] = acc;
          }
        }
        if (SH) {   // the score head's outputs: gg = st
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            float acc = __ldg(P.net.b3 + 2 + d);
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) acc += red2[(red_w + w4) * 32 + red_l + 4 * d];
            gg[d] = acc;
          }
        }
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const float pr = pr_tot[d];
          const float x = sxs[2 * myp + d];
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
              if (SH) f = __fadd_rn(__fadd_rn(f, __fmul_rn(st.c[0], gg[d])), st.c[1]);   // operation order of idff_field_generic.cuh
              if (heavy) f = __fadd_rn(f, gg[d]);
            }
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
        ebar576();   // the next stage's input is in sxs
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ===== epilogue warps: thread <-> 4 consecutive hidden units x 4 particles (see the header comment)
    const int q = warp & 3, pq = warp >> 2;
    const int p0 = pq * HP;
    const int g = lane >> 2, c4 = lane & 3;
    // 8-byte slot of this thread's 4 units inside the 64-byte operand row of its lane quarter.  Not g itself: a 64-bit
    // store is served per half-warp (g = 0..3 | 4..7, 4 particle rows each); rows c4 and c4 + 2 fall on the same banks
    // and the SWIZZLE_64B XOR between them is one 16-byte chunk, so a half-warp must own chunks {0, 2} or {1, 3}.
    const int slot = (g & 1) | (((g >> 1) & 1) << 2) | ((g >> 2) << 1);
    float* red = misc + Smem::f_red;       // [16 warps][32]
    float* sxs = misc + Smem::f_xs;        // stage input of the 64 particles, [p][2]
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + p0;   // + (16 << 16) for the upper 16 lanes
    uint32_t aphase = 0, aphase1 = 0, kphase = 0;
    // this thread's 8-byte group inside the operand rows of its particles p0 + 4r + c4: operand k-block mt*4 + q, 16-byte
    // chunk (slot >> 1) ^ ((row >> 1) & 3) [SWIZZLE_64B], half slot & 1.  (row >> 1) & 3 = ((r & 1) << 1) | (c4 >> 1).
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
    // after writing one half of the B operand: make it visible to the tensor core and hand over to the MMA thread
    auto publish = [&](int mt) {
      fence_proxy_async();
      mbar_arrive(&b_ready[mt]);
    };
    // z = main + 2^-11 * corr, jet j, element e = 8*ld + 2*r + h
    auto load_acc = [&](int j, float (&z)[HP]) {
      uint32_t m[HP], c[HP];
      tmem_ld8(lane_base + j * NP, m);
      tmem_ld8(lane_base + (16u << 16) + j * NP, m + 8);
      tmem_ld8(lane_base + COL_CORR + j * NP, c);
      tmem_ld8(lane_base + (16u << 16) + COL_CORR + j * NP, c + 8);
      tmem_wait_ld();
      const uint64_t inv2 = pk2(LO_INV, LO_INV);
#pragma unroll
      for (int i = 0; i < HP; i += 2)   // FFMA2: the loaded registers are consecutive pairs already
        upk2(fma2(pk2(__uint_as_float(c[i]), __uint_as_float(c[i + 1])), inv2, pk2(__uint_as_float(m[i]), __uint_as_float(m[i + 1]))),
             z[i], z[i + 1]);
    };
    // boundary evaluations: jet 0 summed over the two accumulator sets [main 64 | corr 64] at columns 0 and 128
    auto load_acc2 = [&](float (&z)[HP], int buf) {
      uint32_t m0[HP], c0[HP], m1[HP], c1[HP];
      const uint32_t lb = lane_base + (buf ? 256u : 0u);
      tmem_ld8(lb, m0);
      tmem_ld8(lb + (16u << 16), m0 + 8);
      tmem_ld8(lb + 64, c0);
      tmem_ld8(lb + (16u << 16) + 64, c0 + 8);
      tmem_ld8(lb + 128, m1);
      tmem_ld8(lb + (16u << 16) + 128, m1 + 8);
      tmem_ld8(lb + 192, c1);
      tmem_ld8(lb + (16u << 16) + 192, c1 + 8);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < HP; ++i)
        z[i] = fmaf(__uint_as_float(c0[i]) + __uint_as_float(c1[i]), LO_INV, __uint_as_float(m0[i]) + __uint_as_float(m1[i]));
    };
    auto wait_acc = [&](int buf) {
      if (DB && buf) {
        mbar_wait(acc_full1, aphase1);
        aphase1 ^= 1u;
      } else {
        mbar_wait(acc_full, aphase);
        aphase ^= 1u;
      }
      tc_fence_after();
    };
    auto wait_kfree = [&]() {
      mbar_wait(kfree, kphase);
      kphase ^= 1u;
    };
    auto release_acc = [&](int buf) {
      tc_fence_before();
      mbar_arrive((DB && buf) ? acc_free1 : acc_free);
    };
    // sum over the thread's 4 units, then over the 8 threads that share its particles: v[2r + d] -> lane g keeps v[g]
    auto reduce_units = [&](float (&v)[8]) -> float {
#pragma unroll
      for (int off = 16, cnt = 4; off >= 4; off >>= 1, cnt >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
          const float send = up ? v[i] : v[i + cnt];
          const float keep = up ? v[i + cnt] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      return v[0];
    };

    mbar_arrive(acc_free);   // nothing to read out before the very first m-tile
    if (DB) mbar_arrive(acc_free1);
    for (long it = 0; it < n_iter; ++it) {
      const int gs = set_of(it);
      ebar576();   // the integrator warps have written the tile's start points to sxs
#pragma unroll 1
      for (int ev = 0; ev < n_eval; ++ev) {
        const Stage st = stage_of(ev, gs);
        const bool heavy = heavy_of(st.kind, gs);
#ifdef IDFF_TC16_STAMPS
        const bool stamp = blockIdx.x == 0 && it == 0 && ev == 1 && tid == 0;
#else
        constexpr bool stamp = false;
#endif
        if (stamp) g_tc16_stamps[0] = clock64();
        // ---- layer 1 (K = 3, computed directly): a1 and, for interior evaluations, a1' = s'(z1) * (W0 d).
        // Every MMA of the previous evaluation is complete (this thread waited for its last accumulators).
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
          const int n0 = mt * 128 + q * 32 + 4 * slot;
          float2 xp[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) xp[r] = *reinterpret_cast<const float2*>(sxs + 2 * (p0 + 4 * r + c4));
          const float4 wx0 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 0 * W + n0));
          const float4 wx1 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 1 * W + n0));
          const float4 wtt = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + (SH ? 4 : 2) * W + n0));
          const float4 bb0 = __ldg(reinterpret_cast<const float4*>(P.net.b0 + (size_t)P.tidx * W + n0));
          const float4 zz1 = __ldg(reinterpret_cast<const float4*>(P.net.zd1 + n0));
          // SH: the network's input is cat[x, x, t] -- rows 2, 3 of W0^T meet the same x again
          const float4 wx2 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + (SH ? 2 : 0) * W + n0));
          const float4 wx3 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + (SH ? 3 : 1) * W + n0));
          float v0[HP], v1[HP];
#pragma unroll
          for (int e = 0; e < HP; ++e) {
            const int r = (e >> 1) & 3;
            const float zb = fmaf(IDFF_UC(wtt, e), st.t, IDFF_UC(bb0, e));
            float s1, E;
            float z1 = fmaf(IDFF_UC(wx1, e), xp[r].y, fmaf(IDFF_UC(wx0, e), xp[r].x, zb));
            if (SH) z1 = fmaf(IDFF_UC(wx3, e), xp[r].y, fmaf(IDFF_UC(wx2, e), xp[r].x, z1));
            selu_jet(z1, v0[e], s1, E);
            v1[e] = s1 * IDFF_UC(zz1, e);
          }
          uint32_t sa[2], hw[8], lw[8];
          make_sa(mt, sa);
          pack_jet(v0, hw, lw);
          if (heavy) {
            store_jet<0>(sa, hw, lw);
            pack_jet(v1, hw, lw);
            store_jet<1>(sa, hw, lw);
          } else {
            store_jet<0, 4096>(sa, hw, lw);
          }
          publish(mt);
        }
        if (stamp) g_tc16_stamps[1] = clock64();
        // ---- gemm 0: layer 2 forward
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
          const int n0 = mt * 128 + q * 32 + 4 * slot;
          const int abuf = (DB && !heavy && mt == 1) ? 1 : 0;   // one-jet passes: m-tile 1 has its own accumulator set
          wait_acc(abuf);
          if (stamp && mt == 0) g_tc16_stamps[2] = clock64();
          float z[HP], zd[HP];
          if (heavy) {
            load_acc(0, z);
            load_acc(1, zd);
          } else {
            load_acc2(z, abuf);
          }
          release_acc(abuf);
          const float4 bb1 = __ldg(reinterpret_cast<const float4*>(P.net.b1 + n0));
          const uint32_t stash = lane_base + COL_STASH + mt * 128;
          uint32_t h0[8], l0[8], h1[8], l1[8];
          if (heavy) {   // stash z2' (raw) and the derivative code of z2
            uint32_t u[HP];
#pragma unroll
            for (int e = 0; e < HP; ++e) u[e] = __float_as_uint(zd[e]);
            tmem_st8(stash + NP, u);
            tmem_st8(stash + (16u << 16) + NP, u + 8);
#pragma unroll
            for (int e = 0; e < HP; ++e) {
              float s1, code;
              selu_jet_code(z[e] + IDFF_UC(bb1, e), z[e], s1, code);
              u[e] = __float_as_uint(code);
              zd[e] = s1 * zd[e];
            }
            tmem_st8(stash, u);
            tmem_st8(stash + (16u << 16), u + 8);
            tmem_wait_st();
            pack_jet(z, h0, l0);
            pack_jet(zd, h1, l1);
          } else {
#pragma unroll
            for (int e = 0; e < HP; ++e) {
              float s1, E;
              selu_jet(z[e] + IDFF_UC(bb1, e), z[e], s1, E);
            }
            pack_jet(z, h0, l0);
          }
          if (mt == 0) wait_kfree();   // k-blocks 0-1 have been read by every MMA of this pass
          uint32_t sa[2];
          make_sa(mt, sa);
          if (heavy) {
            store_jet<0>(sa, h0, l0);
            store_jet<1>(sa, h1, l1);
          } else {
            store_jet<0, 4096>(sa, h0, l0);
          }
          publish(mt);
        }
        if (stamp) g_tc16_stamps[3] = clock64();
        // ---- gemm 1: layer 3 forward; x1hat partial sums; adjoint seeds
        {
          float rsum = 0.0f, rsum2 = 0.0f;
#pragma unroll 1
          for (int mt = 0; mt < 2; ++mt) {
            const int n0 = mt * 128 + q * 32 + 4 * slot;
            const int abuf = (DB && !heavy && mt == 1) ? 1 : 0;
            wait_acc(abuf);
            if (stamp && mt == 0) g_tc16_stamps[4] = clock64();
            float z[HP], zd[HP];
            if (heavy) {
              load_acc(0, z);
              load_acc(1, zd);
            } else {
              load_acc2(z, abuf);
            }
            release_acc(abuf);
            const float4 bb2 = __ldg(reinterpret_cast<const float4*>(P.net.b2 + n0));
            const float4 uu3 = __ldg(reinterpret_cast<const float4*>(P.net.u3 + n0));
            const float4 w30 = __ldg(reinterpret_cast<const float4*>(P.net.w3 + 0 * W + n0));
            const float4 w31 = __ldg(reinterpret_cast<const float4*>(P.net.w3 + 1 * W + n0));
            const float4 w32 = __ldg(reinterpret_cast<const float4*>(P.net.w3 + (SH ? 2 : 0) * W + n0));
            const float4 w33 = __ldg(reinterpret_cast<const float4*>(P.net.w3 + (SH ? 3 : 1) * W + n0));
            float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float part2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ld = 0; ld < 2; ++ld)       // fixed summation order over the thread's units 0, 1, 2, 3
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const int e = 8 * ld + 2 * r + h;
                  float a, s1, E;
                  selu_jet(z[e] + IDFF_UC(bb2, e), a, s1, E);
                  part[2 * r] += a * IDFF_UC(w30, e);
                  part[2 * r + 1] += a * IDFF_UC(w31, e);
                  if (SH) {
                    part2[2 * r] += a * IDFF_UC(w32, e);
                    part2[2 * r + 1] += a * IDFF_UC(w33, e);
                  }
                  if (heavy) {   // adjoints of z3 and z3', in place
                    const float A0 = st.c[0] * IDFF_UC(uu3, e), A1 = st.c[1] * IDFF_UC(uu3, e);
                    z[e] = fmaf(A1 * E, zd[e], A0 * s1);
                    zd[e] = A1 * s1;
                  }
                }
            rsum += reduce_units(part);
            if (SH) rsum2 += reduce_units(part2);
            if (heavy) {
              uint32_t h0[8], l0[8], h1[8], l1[8];
              pack_jet(z, h0, l0);
              pack_jet(zd, h1, l1);
              if (mt == 0) wait_kfree();
              uint32_t sa[2];
              make_sa(mt, sa);
              store_jet<0>(sa, h0, l0);
              store_jet<1>(sa, h1, l1);
              publish(mt);
            } else if (mt == 0) {
              wait_kfree();
            }
          }
          red[warp * 32 + lane] = rsum;
          if (SH) red2[warp * 32 + lane] = rsum2;
        }
        if (stamp) g_tc16_stamps[5] = clock64();
        ebar576();   // x1hat partial sums are in red[]: the integrator warps add them up
        if (stamp) g_tc16_stamps[6] = clock64();
        if (heavy) {
          // ---- gemm 2: reverse, layer 3 -> 2
#pragma unroll 1
          for (int mt = 0; mt < 2; ++mt) {
            wait_acc(0);
            if (stamp && mt == 0) g_tc16_stamps[7] = clock64();
            float a0[HP], a1[HP];
            load_acc(0, a0);
            load_acc(1, a1);
            release_acc(0);
            {
              const uint32_t stash = lane_base + COL_STASH + mt * 128;
              uint32_t cd[HP], zdu[HP];
              tmem_ld8(stash, cd);
              tmem_ld8(stash + (16u << 16), cd + 8);
              tmem_ld8(stash + NP, zdu);
              tmem_ld8(stash + (16u << 16) + NP, zdu + 8);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < HP; ++e) {
                const float code = __uint_as_float(cd[e]);
                const float s1 = code < 0.0f ? kSeluLambda : code, E = code < 0.0f ? 0.0f : code;
                a0[e] = fmaf(a1[e] * E, __uint_as_float(zdu[e]), a0[e] * s1);
                a1[e] = a1[e] * s1;
              }
            }
            uint32_t h0[8], l0[8], h1[8], l1[8];
            pack_jet(a0, h0, l0);
            pack_jet(a1, h1, l1);
            if (mt == 0) wait_kfree();
            uint32_t sa[2];
            make_sa(mt, sa);
            store_jet<0>(sa, h0, l0);
            store_jet<1>(sa, h1, l1);
            publish(mt);
          }
          if (stamp) g_tc16_stamps[8] = clock64();
          // ---- gemm 3: reverse, layer 2 -> 1 -> x
          float gsum = 0.0f;
#pragma unroll 1
          for (int mt = 0; mt < 2; ++mt) {
            const int n0 = mt * 128 + q * 32 + 4 * slot;
            wait_acc(0);
            if (stamp && mt == 0) g_tc16_stamps[9] = clock64();
            float a0[HP], a1[HP];
            load_acc(0, a0);
            load_acc(1, a1);
            release_acc(0);
            const float4 wx0 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 0 * W + n0));
            const float4 wx1 = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 1 * W + n0));
            const float4 wtt = __ldg(reinterpret_cast<const float4*>(P.net.wt0 + 2 * W + n0));
            const float4 bb0 = __ldg(reinterpret_cast<const float4*>(P.net.b0 + (size_t)P.tidx * W + n0));
            const float4 zz1 = __ldg(reinterpret_cast<const float4*>(P.net.zd1 + n0));
            float2 xp[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) xp[r] = *reinterpret_cast<const float2*>(sxs + 2 * (p0 + 4 * r + c4));
            float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ld = 0; ld < 2; ++ld)
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const int e = 8 * ld + 2 * r + h;
                  const float zb = fmaf(IDFF_UC(wtt, e), st.t, IDFF_UC(bb0, e));
                  float a, s1, E;
                  selu_jet(fmaf(IDFF_UC(wx1, e), xp[r].y, fmaf(IDFF_UC(wx0, e), xp[r].x, zb)), a, s1, E);
                  const float P1 = fmaf(a1[e] * E, IDFF_UC(zz1, e), a0[e] * s1);
                  part[2 * r] += P1 * IDFF_UC(wx0, e);
                  part[2 * r + 1] += P1 * IDFF_UC(wx1, e);
                }
            gsum += reduce_units(part);
            if (mt == 0) wait_kfree();   // keeps the barrier phase in step
          }
          red[warp * 32 + lane] = gsum;
          if (stamp) g_tc16_stamps[10] = clock64();
          ebar576();   // gradient partial sums are in red[]
        }
        if (stamp) g_tc16_stamps[11] = clock64();
        ebar576();   // the integrator warps have written the next stage's input to sxs
        if (stamp) g_tc16_stamps[12] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be multicasting into this CTA's ring / arriving on its barriers
  if (warp == NEW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace tch

int tensor16_debug_stamps(long long* h_out, int n) {
  IDFF_CUDA(cudaMemcpyFromSymbol(h_out, tch::g_tc16_stamps, sizeof(long long) * (n < 32 ? n : 32)));
  return IDFF_OK;
}

// idff_fixup.cu: the range safety net of IDFF_IMPL_AUTO
int fixup_nonfinite(const idff_weights_t* w, const idff_field_params_t* p, int mode, int n_sets, int n_iv, const Rk4Table* rk,
                    const SweepTab* d_tabs, const Stage* g_st, const float* g_dt, const float* d_x_in, const float* d_x0_cfm,
                    float* d_out, float* d_traj, int64_t N, void* stream);

int keep_input(idff_weights_t* wm, const float** src, const float* d_out, bool force, size_t bytes, void* stream);

bool tensor16_supports(const idff_weights_t* w, const idff_field_params_t* p) {
  const NetView& v = w->v;
  if (v.w != 256 || v.tc16_tiles == nullptr || p->dim != 2) return false;
  if (v.din == 5 && v.dout == 4)   // the score-head network, cat[x, x, t] -> (x1hat, st): its field and its CFM comparator
    return p->variant == IDFF_FIELD_SCOREHEAD || p->variant == IDFF_FIELD_CFM;
  if (v.din != 3 || v.dout != 2) return false;
  if (!(p->variant == IDFF_FIELD_MOMENTUM || p->variant == IDFF_FIELD_CFM)) return false;
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  return nj <= 2;
}

static int verify_tc16_launch(idff_weights_t* w, void* stream) {
  if (w->tc_verified & 4) return IDFF_OK;
  int flag = 0;
  IDFF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  IDFF_CUDA(cudaMemcpyFromSymbol(&flag, g_tc16_misaligned, sizeof(int)));
  if (flag) {
    set_error("tcgen05 kernel: dynamic shared memory is not 1024-byte aligned on this driver; use impl = IDFF_IMPL_FUSED");
    return IDFF_EUNSUPPORTED;
  }
  w->tc_verified |= 4;
  return IDFF_OK;
}

static int launch_tc16(const idff_weights_t* w, const idff_field_params_t* p, tch::TParams& P, const Rk4Table* rk,
                       const Stage* one, void* stream) {
  if (!tensor16_supports(w, p)) {
    set_error("3xFP16 tensor-core sampler: needs a 3-256-256-256-2 network, dim 2, order <= 2 (momentum / CFM field) or the "
              "5-256-256-256-4 score-head network");
    return IDFF_EUNSUPPORTED;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long tiles = P.mode == 2 ? P.tiles_per_set * P.n_sets : (P.N + tch::NP - 1) / tch::NP;
  int grid = (int)(tiles < sms ? tiles : sms);
  grid = (grid + 1) & ~1;   // CTA pairs
  static thread_local Rk4Table rk0;
  static thread_local Stage one0;
  const Rk4Table& rkr = rk ? *rk : rk0;
  const Stage& oner = one ? *one : one0;
  const int smem = tch::Smem::total;
  const bool sh = w->v.din == 5;   // score-head network (field or CFM comparator): the one-jet instance
  IDFF_CUDA(cudaFuncSetAttribute(sh ? tch::k_tc16<true> : tch::k_tc16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tch::NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (sh) IDFF_CUDA(cudaLaunchKernelEx(&cfg, tch::k_tc16<true>, P, rkr, oner));
  else IDFF_CUDA(cudaLaunchKernelEx(&cfg, tch::k_tc16<false>, P, rkr, oner));
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return verify_tc16_launch(const_cast<idff_weights_t*>(w), stream);
}

static void fill_tparams16(const idff_weights_t* w, const idff_field_params_t* p, int mode, int64_t N, tch::TParams* P) {
  memset(P, 0, sizeof(*P));
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  P->net = w->v;
  P->variant = p->variant; P->D = p->dim; P->tidx = p->t_idx; P->reverse = rev; P->mode = mode;
  P->N = (long)N;
  P->counters = w->d_counters;
}

int tensor16_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, const float* d_x0,
                        float* d_out, int64_t N, void* stream) {
  tch::TParams P;
  fill_tparams16(w, p, 0, N, &P);
  const bool safe = p->impl == IDFF_IMPL_AUTO;
  if (safe) {
    int rc = keep_input(const_cast<idff_weights_t*>(w), &d_x, d_out, false, (size_t)N * 2 * sizeof(float), stream);
    if (rc) return rc;
  }
  P.x_in = d_x; P.x0_cfm = d_x0; P.out = d_out;
  Stage st;
  make_stage(*p, t, &st);
  int rc = launch_tc16(w, p, P, nullptr, &st, stream);
  if (rc || !safe) return rc;
  static thread_local Rk4Table tab;
  tab.n_iv = 0;
  tab.st[0] = st;
  return fixup_nonfinite(w, p, 0, 1, 0, &tab, nullptr, nullptr, nullptr, d_x, d_x0, d_out, nullptr, N, stream);
}

// One launch for n_sets parameter sets over the same x0 and grid (the gamma sweeps of the 2-D scripts).
int tensor16_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, const float* d_x0,
                              const Stage* stages, const float* dts, int n_iv, float* d_out, int64_t N, void* stream) {
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
    if (!tensor16_supports(w, &params[g])) {
      set_error("3xFP16 sweep: parameter set %d is not served by this kernel", g);
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
  tch::TParams P;
  fill_tparams16(w, &params[0], 2, N, &P);
  P.n_sets = n_sets;
  P.sweep_n_iv = n_iv;
  const long tiles = ((long)N + tch::NP - 1) / tch::NP;
  P.tiles_per_set = (tiles + 1) & ~1L;
  P.tabs = reinterpret_cast<const SweepTab*>(wm->d_sweep);
  P.x_in = d_x0; P.out = d_out;
  int rc = launch_tc16(w, &params[0], P, nullptr, nullptr, stream);
  bool safe = false;   // range safe unless every set asks for the raw kernel
  for (int g = 0; g < n_sets; ++g) safe = safe || params[g].impl == IDFF_IMPL_AUTO;
  if (rc || !safe) return rc;
  return fixup_nonfinite(w, &params[0], 2, n_sets, n_iv, nullptr, reinterpret_cast<const SweepTab*>(wm->d_sweep), nullptr, nullptr,
                         d_x0, nullptr, d_out, nullptr, N, stream);
}

int tensor16_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                        const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream) {
  static thread_local Rk4Table tab;
  const float* src = d_x0;
  const bool safe = p->impl == IDFF_IMPL_AUTO;
  for (int done = 0; done < n_iv; done += kMaxIntervals) {
    const int cnt = n_iv - done < kMaxIntervals ? n_iv - done : kMaxIntervals;
    tab.n_iv = cnt;
    memcpy(tab.dt, dts + done, sizeof(float) * cnt);
    memcpy(tab.st, stages + 4 * (size_t)done, sizeof(Stage) * 4 * cnt);
    tch::TParams P;
    fill_tparams16(w, p, 1, N, &P);
    if (safe) {
      int rc = keep_input(const_cast<idff_weights_t*>(w), &src, d_out_final, done > 0, (size_t)N * 2 * sizeof(float), stream);
      if (rc) return rc;
    }
    P.x_in = src; P.out = d_out_final;
    P.x0_cfm = done > 0 ? d_x0 : nullptr;   // CFMModel keeps the x captured at t == 0 for the whole solve (order2.py:322-324)
    P.traj = d_out_traj ? d_out_traj + (size_t)done * N * p->dim : nullptr;
    int rc = launch_tc16(w, p, P, &tab, nullptr, stream);
    if (rc) return rc;
    if (safe) {
      rc = fixup_nonfinite(w, p, 1, 1, cnt, &tab, nullptr, nullptr, nullptr, src, P.x0_cfm, d_out_final,
                           P.traj ? P.traj + (size_t)N * 2 : nullptr, N, stream);
      if (rc) return rc;
    }
    src = d_out_final;
  }
  return IDFF_OK;
}

}  // namespace idff
