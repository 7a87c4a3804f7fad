// Layer-by-layer tensor-core IDFF sampler for WIDE networks (3 -> w -> w -> w -> 2, w a multiple of 128, 256..2048):
// BASELINE.json configs[2] (order-3 field, large-width MLP).  Activations of a w >= 512 layer do not fit on chip, so
// every layer pass is a persistent tcgen05 3xFP16 GEMM (kind::f16, fp32 accumulation in TMEM; the split of
// idff_sampler_tc16.cu: v = hi + 2^-11 lo', D = W_hi X_hi + 2^-11 (W_hi X_lo' + W_lo' X_hi)) over activations that live in
// global memory (L2/HBM) already in the tensor core's operand format; the SELU jet / adjoint algebra is the epilogue.
//
//   D[unit][col] = sum_k W[unit][k] * X[col][k]          col = jet*PPT + particle, 128 columns per column tile
//
//   * X operand in global memory: per (column tile, 64-wide k-block) a 32 KB image of a K-major SWIZZLE_128B tile,
//     rows 0..127 the hi copy, rows 128..255 the lo' copy -- one cp.async.bulk per stage lands it next to the 32 KB
//     (hi|lo') weight tile; [main | corr] += W_hi*[X_hi;X_lo'] (N = 256 MMA), corr += W_lo'*X_hi (N = 128 MMA): the two
//     shapes the tensor core runs at full rate [measured, profiles/microbench/tc_f16.cu].
//   * work items (column tile, m-tile) are spread over persistent CTAs; accumulators are double buffered in TMEM
//     (2 buffers x [main 128 | corr 128] = 512 columns), so the epilogue of item i overlaps the MMAs of item i+1;
//     3-stage TMA ring of 64 KB stages.
//   * epilogue thread <-> (hidden unit, quarter of the tile's particles): reads its TMEM lane, applies the layer's jet
//     or adjoint formulas and writes the next layer's operand straight into the global tile images (64 B per warp
//     store), the layer-2 derivative stash, or per-(m-tile, lane-quarter) partial sums of the two small layers
//     (x1hat = W3 a3, grad_x = W0^T P1), which the update kernel adds in a fixed order.
// One call = chunks of particles x (1 + 2|4 + 1) launches per field evaluation; the particle state, the stash and the
// operand buffers are scratch owned by the weights handle.  As in k_tc16, values beyond the fp16 range turn that
// particle's output into NaN (columns never mix); IDFF_IMPL_GENERIC serves such networks.
#include <cuda_fp16.h>

#include <algorithm>
#include <vector>

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

namespace wide {

constexpr int NCOL = 128;                // columns per column tile (jets x particles, padded)
constexpr int XT_BYTES = 2 * NCOL * 128; // one 64-wide k-block of the X operand: hi rows then lo' rows (32 KB)
constexpr int WT_BYTES = 32768;          // one (hi|lo') weight tile [128 units x 64 k] fp16
constexpr int STAGE_BYTES = WT_BYTES + XT_BYTES;
constexpr int NSTAGE = 3;
constexpr int NEW = 16;                  // epilogue warps: lane quarter = w % 4, particle quarter = w / 4
constexpr int NPART = NEW / 4;
constexpr int NE = NEW * 32;
constexpr int NTHREADS = NE + 64;        // + producer warp + MMA warp
constexpr int COL_CORR = 128, COL_BUF = 256;
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;

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
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- layouts ----------------------------------------------------------------------------------------------------------
// NJ jets -> PPT particles per column tile; column of (jet j, particle p) = j*PPT + p; columns >= NJ*PPT are padding.
__host__ __device__ constexpr int ppt_of(int nj) { return nj == 1 ? 128 : (nj == 2 ? 64 : 32); }
// Three jets x 32 particles fill 96 of a tile's 128 columns: a quarter of every MMA and of every X-tile load is padding.  Plain
// solves therefore use 40 particles per tile (120 columns) -- five groups of 8 particles on the four group-warps of a lane
// quarter, the first of which serves a second group (the accumulators are double buffered: there is a whole item of slack).
// Sweeps keep 32 (a tile must not straddle two parameter sets and the reference's 2048 rows are not a multiple of 40).
constexpr int kPpt3Dense = 40;
__host__ __device__ constexpr int hp_of(int ppt) { return ppt == kPpt3Dense ? 8 : ppt / 4; }   // particles per epilogue thread and group
__host__ __device__ constexpr int ng_of(int ppt) { return ppt / hp_of(ppt); }                    // particle groups per tile: 4, or 5

// byte offset of fp16 element (col, k) inside the X operand of one column tile (KB k-blocks of XT_BYTES)
__device__ __forceinline__ size_t xt_off(int col, int k, int lo) {
  return (size_t)(k >> 6) * XT_BYTES + (size_t)(lo * NCOL + col) * 128 + ((((k >> 3) & 7) ^ (col & 7)) << 4) + ((k & 7) << 1);
}
// hi / lo' halves of v for column col, unit k
__device__ __forceinline__ void put16(unsigned char* xt, int col, int k, float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(__fmul_rn(__fsub_rn(v, __half2float(h)), LO_SCALE));
  *reinterpret_cast<__half*>(xt + xt_off(col, k, 0)) = h;
  *reinterpret_cast<__half*>(xt + xt_off(col, k, 1)) = l;
}

// Boundary evaluations (t = 0, t = 1: one jet) of a field with >= 2 jets: the jet-1 column slot is free, so the K sum is
// split over the two slots by k-block parity -- unit k's value goes to slot (k / 64) % 2, an exact zero to the other.
// Adding exact zeros costs no rounding, so each accumulator sees half as many truncating accumulation steps where the
// t = 1 branch amplifies the network's error x100.
template <int PPT>
__device__ __forceinline__ void put_light(unsigned char* xt, int p, int k, float v) {
  const int odd = (k >> 6) & 1;
  put16(xt, p, k, odd ? 0.0f : v);
  put16(xt, PPT + p, k, odd ? v : 0.0f);
}

struct WParams {
  NetView net;
  int w, MT, KB, nj, ppt;      // m-tiles, k-blocks, jets, particles per column tile
  int n_ctiles;                // column tiles of this chunk
  long P;                      // particles of this chunk
  int gemm;                    // 0: W1, 1: W2, 2: W2^T, 3: W1^T
  int heavy;
  Stage st;
  const unsigned char* xin;    // X operand in  [ctile][KB][XT_BYTES]
  unsigned char* xout;         // X operand out
  float* stash;                // [ctile][MT][NPART][NJ*HP][128]
  float* part;                 // [ctile][MT][4 lane quarters][NPART][2*HP]   partial sums of a small layer
  const float* xs;             // [P][2] stage input (layer-1 recompute in the last reverse pass)
  // gamma sweeps: particle row r of the whole batch belongs to parameter set r / set_rows (set_rows is a multiple of every
  // column-tile size, so a column tile never straddles two sets); coef [n_sets][4] = {c1, c2, c3, a} of this evaluation
  const float* coef;
  long set_rows, row0;
};

template <int NV>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[NV], int lane) {
  int cnt = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cnt > 1) {
      cnt >>= 1;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i)
        if (i < cnt) {
          const float send = up ? v[i] : v[i + cnt];
          const float keep = up ? v[i + cnt] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
}

// after warp_reduce_scatter<NV>: NV = 64 -> lane L holds totals 2L, 2L+1; 32 -> L; 16 -> L/2 (pairs of lanes agree)
template <int NV>
__device__ __forceinline__ void store_part(float* part, const float (&v)[NV], int lane) {
  if (NV == 64) {
    part[2 * lane] = v[0];
    part[2 * lane + 1] = v[NV >= 2 ? 1 : 0];
  } else if (NV == 32) {
    part[lane] = v[0];
  } else {
    constexpr int R = 32 / (NV < 32 ? NV : 32);
    if ((lane & (R - 1)) == 0) part[lane / R] = v[0];
  }
}

// ---- layer 1 (K = 3): writes the X operand of gemm 0 -------------------------------------------------------------------
template <int NJ, int PPT>
__global__ void __launch_bounds__(256) k_wide_l1(const __grid_constant__ WParams P) {
  const int w = P.w;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long)P.n_ctiles * w; idx += (long)gridDim.x * blockDim.x) {
    const int ct = (int)(idx / w), n = (int)(idx - (long)ct * w);
    const float wx0 = __ldg(P.net.wt0 + n), wx1 = __ldg(P.net.wt0 + w + n), wtt = __ldg(P.net.wt0 + 2 * w + n);
    const float zb = fmaf(wtt, P.st.t, __ldg(P.net.b0 + n)), zd = __ldg(P.net.zd1 + n);
    unsigned char* xt = P.xout + (size_t)ct * P.KB * XT_BYTES;
    for (int p = 0; p < PPT; ++p) {
      const long row = (long)ct * PPT + p;
      const float x0 = row < P.P ? P.xs[2 * row] : 0.0f, x1 = row < P.P ? P.xs[2 * row + 1] : 0.0f;
      const float z = fmaf(wx1, x1, fmaf(wx0, x0, zb));
      float code, s1, E;
      const float a = selu_fwd(z, code);
      selu_decode(code, s1, E);
      float vals[3] = {a, s1 * zd, E * zd * zd};
      if (NJ >= 2 && !P.heavy) {
        put_light<PPT>(xt, p, n, vals[0]);
      } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) put16(xt, j * PPT + p, n, (j == 0 || P.heavy) ? vals[j] : 0.0f);
      }
    }
  }
}

// ---- one layer pass -------------------------------------------------------------------------------------------------------
template <int NJ, int PPT>
__global__ void __launch_bounds__(NTHREADS, 1) k_wide_layer(const __grid_constant__ WParams P) {
  constexpr int HP = hp_of(PPT), NG = ng_of(PPT);
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((s32(smem) & 1023u) != 0u) {   // SWIZZLE_128B operands need a 1024-byte aligned base: report and do nothing
    if (threadIdx.x == 0) g_tc_misaligned = 1;
    return;
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGE;
  uint64_t* acc_full = bars + 2 * NSTAGE;    // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(&acc_full[q], 1); mbar_init(&acc_empty[q], NE); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NEW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int MT = P.MT, KB = P.KB;
  const long n_items = (long)P.n_ctiles * MT;
  const unsigned char* wt = reinterpret_cast<const unsigned char*>(P.net.tcw16_tiles) + (size_t)P.gemm * MT * KB * WT_BYTES;

  if (warp == NEW) {
    // TMA producer: whole warp in uniform control flow, one elected lane issues (operands in uniform registers)
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ct = (int)(item / MT), mt = (int)(item - (long)ct * MT);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u);
        if (leader) {
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          unsigned char* dst = smem + stage * STAGE_BYTES;
          tma_load_1d(dst, wt + ((size_t)mt * KB + kb) * WT_BYTES, WT_BYTES, &full[stage]);
          tma_load_1d(dst + WT_BYTES, P.xin + ((size_t)ct * KB + kb) * XT_BYTES, XT_BYTES, &full[stage]);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == NEW + 1) {
    // MMA issuer: whole warp in uniform control flow, one elected lane issues -- descriptors and TMEM addresses stay in
    // uniform registers, a tcgen05.mma costs a few instructions (from an `if (lane == 0)` region each one is preceded by
    // an ELECT / R2UR.BROADCAST loop and the issuing thread, not the tensor pipe, sets the pace)
    // kind::f16: D fp32 (bit 4), A = B = fp16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc_w = (1u << 4) | ((uint32_t)(2 * NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_n = (1u << 4) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const bool leader = elect_one();
    const uint32_t ring0 = s32(smem);
    int stage = 0, q = 0;
    uint32_t phase = 0, ephase0 = 0, ephase1 = 0;
    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
      // the epilogue has read this accumulator buffer out
      if (q == 0) { mbar_wait(&acc_empty[0], ephase0 ^ 1u); ephase0 ^= 1u; }
      else { mbar_wait(&acc_empty[1], ephase1 ^ 1u); ephase1 ^= 1u; }
      tc_fence_after();
      const uint32_t dset = tmem + q * COL_BUF;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t a_hi0 = make_desc(ring0 + (uint32_t)(stage * STAGE_BYTES));
        const uint64_t a_lo0 = make_desc(ring0 + (uint32_t)(stage * STAGE_BYTES + 16384));
        const uint64_t b_0 = make_desc(ring0 + (uint32_t)(stage * STAGE_BYTES + WT_BYTES));
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {   // descriptors advance by 32 bytes (2 address units) per 16-wide k-step
            mma_f16(dset, a_hi0 + (uint64_t)(2 * ks), b_0 + (uint64_t)(2 * ks), idesc_w, (uint32_t)((kb | ks) != 0));   // [main | corr] += W_hi * [X_hi ; X_lo']
            mma_f16(dset + COL_CORR, a_lo0 + (uint64_t)(2 * ks), b_0 + (uint64_t)(2 * ks), idesc_n, 1u);                // corr += W_lo' * X_hi
          }
          mma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
      if (leader) mma_commit(&acc_full[q]);
      __syncwarp();
      q ^= 1;
    }
  } else {
    // ===== epilogue warps
    const int lq = warp & 3, h = warp >> 2;
    int q = 0;
    uint32_t fphase[2] = {0, 0};
    const int w = P.w;
    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ct = (int)(item / MT), mt = (int)(item - (long)ct * MT);
      const int n = mt * 128 + lq * 32 + lane;   // this thread's hidden unit (of the layer this pass produces)
      mbar_wait_sleep(&acc_full[q], fphase[q]);
      fphase[q] ^= 1u;
      tc_fence_after();
      const uint32_t lane_base = tmem + ((uint32_t)(lq * 32) << 16) + q * COL_BUF;
      unsigned char* xt = P.xout + (size_t)ct * KB * XT_BYTES;
      // particle groups of this warp: h, and for the 40-particle geometry group 4 on the h == 0 warps.  The accumulator buffer is
      // handed back to the MMA warp after the LAST group's read-out (double buffered: the next item's MMAs are not waiting for it)
#pragma unroll 1
      for (int grp = h; grp < NG; grp += NPART) {
      const int p0 = grp * HP;
      float acc[NJ][HP];
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j == 0 || P.heavy) {
#pragma unroll
          for (int c8 = 0; c8 < HP / 8; ++c8) {
            uint32_t m0[8], c0[8];
            const int col = j * PPT + p0 + c8 * 8;
            tmem_ld8(lane_base + col, m0);
            tmem_ld8(lane_base + COL_CORR + col, c0);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][c8 * 8 + i] = fmaf(__uint_as_float(c0[i]), LO_INV, __uint_as_float(m0[i]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < HP; ++i) acc[j][i] = 0.0f;
        }
      if (NJ >= 2 && !P.heavy) {
        // boundary evaluation: the K sum was split by k-block parity over the jet-0 / jet-1 column slots (see put_light)
#pragma unroll
        for (int c8 = 0; c8 < HP / 8; ++c8) {
          uint32_t m1[8], c1[8];
          const int col = PPT + p0 + c8 * 8;
          tmem_ld8(lane_base + col, m1);
          tmem_ld8(lane_base + COL_CORR + col, c1);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[0][c8 * 8 + i] += fmaf(__uint_as_float(c1[i]), LO_INV, __uint_as_float(m1[i]));
        }
      }
      if (grp + NPART >= NG) {
        tc_fence_before();
        mbar_arrive(&acc_empty[q]);
      }
      float* stash = P.stash + ((((size_t)ct * MT + mt) * NG + grp) * (NJ * HP)) * 128 + lq * 32 + lane;
      float* part = P.part + ((((size_t)ct * MT + mt) * 4 + lq) * NG + grp) * (2 * HP);
      if (P.gemm == 0) {
        // layer 2 forward: activations of all jets, derivative stash
        const float bb = __ldg(P.net.b1 + n);
#pragma unroll
        for (int i = 0; i < HP; ++i) {
          float code, s1, E;
          const float a = selu_fwd(acc[0][i] + bb, code);
          selu_decode(code, s1, E);
          float vals[3] = {a, 0.0f, 0.0f};
          if (NJ >= 2) vals[1] = s1 * acc[NJ >= 2 ? 1 : 0][i];
          if (NJ >= 3) vals[2] = fmaf(E * acc[NJ >= 2 ? 1 : 0][i], acc[NJ >= 2 ? 1 : 0][i], s1 * acc[NJ >= 3 ? 2 : 0][i]);
          if (P.heavy) {
            stash[(0 * HP + i) * 128] = code;
            if (NJ >= 2) stash[((NJ >= 2 ? 1 : 0) * HP + i) * 128] = acc[NJ >= 2 ? 1 : 0][i];
            if (NJ >= 3) stash[((NJ >= 3 ? 2 : 0) * HP + i) * 128] = acc[NJ >= 3 ? 2 : 0][i];
          }
          if (NJ >= 2 && !P.heavy) {
            put_light<PPT>(xt, p0 + i, n, vals[0]);
          } else {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
              if (j == 0 || P.heavy) put16(xt, j * PPT + p0 + i, n, vals[j]);
          }
        }
      } else if (P.gemm == 1) {
        // layer 3 forward: x1hat partial sums, adjoint seeds of c1*phi + c2*D_d phi + c3*D_d^2 phi
        const float bb = __ldg(P.net.b2 + n), u = __ldg(P.net.u3 + n);
        const float w30 = __ldg(P.net.w3 + n), w31 = __ldg(P.net.w3 + w + n);
        float c0 = P.st.c[0], c1 = P.st.c[1], c2 = P.st.c[2];
        if (P.coef) {
          const float4 cf = __ldg(reinterpret_cast<const float4*>(P.coef) + (P.row0 + (long)ct * PPT) / P.set_rows);
          c0 = cf.x; c1 = cf.y; c2 = cf.z;
        }
        const float A0 = c0 * u, A1 = c1 * u, A2 = c2 * u;
        float pv[2 * HP];
#pragma unroll
        for (int i = 0; i < HP; ++i) {
          float code, s1, E;
          const float a = selu_fwd(acc[0][i] + bb, code);
          selu_decode(code, s1, E);
          pv[2 * i] = a * w30;
          pv[2 * i + 1] = a * w31;
          if (P.heavy) {
            const float zd = NJ >= 2 ? acc[NJ >= 2 ? 1 : 0][i] : 0.0f, zdd = NJ >= 3 ? acc[NJ >= 3 ? 2 : 0][i] : 0.0f;
            float vals[3];
            vals[0] = A0 * s1;
            if (NJ >= 2) vals[0] = fmaf(A1 * E, zd, vals[0]);
            if (NJ >= 3) vals[0] = fmaf(A2 * E, fmaf(zd, zd, zdd), vals[0]);
            vals[1] = NJ >= 3 ? fmaf(2.0f * A2 * E, zd, A1 * s1) : A1 * s1;
            vals[2] = A2 * s1;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              put16(xt, j * PPT + p0 + i, n, vals[j]);
            }
          }
        }
        warp_reduce_scatter<2 * HP>(pv, lane);
        store_part<2 * HP>(part, pv, lane);
      } else if (P.gemm == 2) {
        // reverse, layer 3 -> 2: combine the incoming adjoints with the stash
#pragma unroll
        for (int i = 0; i < HP; ++i) {
          float s1, E;
          selu_decode(stash[(0 * HP + i) * 128], s1, E);
          const float zd = NJ >= 2 ? stash[((NJ >= 2 ? 1 : 0) * HP + i) * 128] : 0.0f;
          const float zdd = NJ >= 3 ? stash[((NJ >= 3 ? 2 : 0) * HP + i) * 128] : 0.0f;
          const float A0 = acc[0][i], A1 = NJ >= 2 ? acc[NJ >= 2 ? 1 : 0][i] : 0.0f, A2 = NJ >= 3 ? acc[NJ >= 3 ? 2 : 0][i] : 0.0f;
          float vals[3];
          vals[0] = A0 * s1;
          if (NJ >= 2) vals[0] = fmaf(A1 * E, zd, vals[0]);
          if (NJ >= 3) vals[0] = fmaf(A2 * E, fmaf(zd, zd, zdd), vals[0]);
          vals[1] = NJ >= 3 ? fmaf(2.0f * A2 * E, zd, A1 * s1) : A1 * s1;
          vals[2] = A2 * s1;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            put16(xt, j * PPT + p0 + i, n, vals[j]);
          }
        }
      } else {
        // reverse, layer 2 -> 1 -> x: layer-1 derivative data recomputed from x; grad_x partial sums
        const float wx0 = __ldg(P.net.wt0 + n), wx1 = __ldg(P.net.wt0 + w + n), wtt = __ldg(P.net.wt0 + 2 * w + n);
        const float zb = fmaf(wtt, P.st.t, __ldg(P.net.b0 + n)), zd = __ldg(P.net.zd1 + n);
        float pv[2 * HP];
#pragma unroll
        for (int i = 0; i < HP; ++i) {
          const long row = (long)ct * PPT + p0 + i;
          const float x0 = row < P.P ? P.xs[2 * row] : 0.0f, x1 = row < P.P ? P.xs[2 * row + 1] : 0.0f;
          const float z = fmaf(wx1, x1, fmaf(wx0, x0, zb));
          float code, s1, E;
          (void)selu_fwd(z, code);
          selu_decode(code, s1, E);
          const float A0 = acc[0][i], A1 = NJ >= 2 ? acc[NJ >= 2 ? 1 : 0][i] : 0.0f, A2 = NJ >= 3 ? acc[NJ >= 3 ? 2 : 0][i] : 0.0f;
          float P1 = A0 * s1;
          if (NJ >= 2) P1 = fmaf(A1 * E, zd, P1);
          if (NJ >= 3) P1 = fmaf(A2 * E, zd * zd, P1);
          pv[2 * i] = P1 * wx0;
          pv[2 * i + 1] = P1 * wx1;
        }
        warp_reduce_scatter<2 * HP>(pv, lane);
        store_part<2 * HP>(part, pv, lane);
      }
      }   // particle groups
      q ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NEW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- sum the partials of a small layer over (m-tile, lane quarter) in a fixed order ----------------------------------------
template <int NJ, int PPT>
__global__ void k_wide_sum(const __grid_constant__ WParams P, float* __restrict__ dst) {
  constexpr int HP = hp_of(PPT), NG = ng_of(PPT);
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < P.P * 2; o += (long)gridDim.x * blockDim.x) {
    const long row = o >> 1;
    const int d = (int)(o & 1), ct = (int)(row / PPT), p = (int)(row - (long)ct * PPT), h = p / HP, pl = p - h * HP;
    float s = 0.0f;
    for (int mt = 0; mt < P.MT; ++mt)
      for (int lq = 0; lq < 4; ++lq) s += P.part[((((size_t)ct * P.MT + mt) * 4 + lq) * NG + h) * (2 * HP) + 2 * pl + d];
    dst[o] = s;
  }
}

// ---- combine + integrator update (rk4 3/8 rule; mode 0 = plain evaluation) ------------------------------------------------------
struct UParams {
  long P, N_total, row0;
  int variant, mode, sg, heavy, last;
  float dt;
  Stage st;
  const float *pred, *grad, *b3;
  float *y, *k1, *k2, *k3, *xs, *x0c;
  float *out, *traj_row;   // global outputs (already offset to this chunk)
  const float* coef;       // gamma sweeps: [n_sets][4] = {c1, c2, c3, a} of this evaluation, set = (row0 + row) / set_rows
  long set_rows;
};
__global__ void k_wide_update(const __grid_constant__ UParams U) {
  const float third = (float)(1.0 / 3.0);
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < U.P * 2; o += (long)gridDim.x * blockDim.x) {
    const int d = (int)(o & 1);
    const float pr = U.pred[o] + __ldg(U.b3 + d), x = U.xs[o];
    float f;
    if (U.variant == IDFF_FIELD_CFM) {
      if (U.st.kind == 0) U.x0c[o] = x;
      f = __fsub_rn(pr, U.x0c[o]);
    } else {
      const float diff = __fsub_rn(pr, x);
      if (U.st.kind == 0) f = diff;
      else if (U.st.kind == 1) f = __fdiv_rn(diff, U.st.end_div);
      else {
        const float a = U.coef ? __ldg(U.coef + 4 * ((U.row0 + (o >> 1)) / U.set_rows) + 3) : U.st.a;
        f = __fdiv_rn(__fmul_rn(a, diff), U.st.den);
        if (U.heavy) f = __fadd_rn(f, U.grad[o]);
      }
    }
    if (U.mode == 0) {
      U.out[o] = f;
      continue;
    }
    const float dt = U.dt, y = U.y[o];
    float xn;
    if (U.sg == 0) {
      U.k1[o] = f;
      xn = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, f), third));
    } else if (U.sg == 1) {
      U.k2[o] = f;
      xn = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(f, __fmul_rn(U.k1[o], third))));
    } else if (U.sg == 2) {
      U.k3[o] = f;
      xn = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(U.k1[o], U.k2[o]), f)));
    } else {
      const float sum = __fadd_rn(__fadd_rn(U.k1[o], __fmul_rn(3.0f, __fadd_rn(U.k2[o], U.k3[o]))), f);
      xn = __fadd_rn(y, __fmul_rn(__fmul_rn(sum, dt), 0.125f));
      U.y[o] = xn;
      if (U.traj_row) U.traj_row[o] = xn;
      if (U.last) U.out[o] = xn;
    }
    U.xs[o] = xn;
  }
}

}  // namespace wide

// ---- host side ---------------------------------------------------------------------------------------------------------------
bool wide_supports(const idff_weights_t* w, const idff_field_params_t* p) {
  const NetView& v = w->v;
  const int min_w = p->impl == IDFF_IMPL_LAYERED ? 256 : 512;   // at 256 the fused kernel is the default
  if (v.tcw16_tiles == nullptr || v.w < min_w || p->dim != 2 || v.din != 3 || v.dout != 2) return false;
  return p->variant == IDFF_FIELD_MOMENTUM || p->variant == IDFF_FIELD_CFM;
}

static thread_local int g_wide_dense3 = 1;   // diagnostic: 0 = three-jet solves in the 32-particle tile geometry (idff_debug_wide_dense3)
void wide_set_dense3(int on) { g_wide_dense3 = on ? 1 : 0; }

struct WideScratch {
  unsigned char *xa, *xb;
  float *stash, *part, *pred, *grad, *y, *k1, *k2, *k3, *xs, *x0c;
};

static int wide_run(const idff_weights_t* w, const idff_field_params_t* p, int mode, const float* d_x, const float* d_x0,
                    const Stage* stages, const float* dts, int n_eval, float* d_out, float* d_traj, int64_t N, void* stream,
                    const float* d_coef = nullptr, int n_sets = 1, long set_rows = 0) {
  using namespace wide;
  if (!wide_supports(w, p)) {
    set_error("wide tensor-core sampler: needs a 3-w-w-w-2 network with w in {256..2048} (multiple of 128), dim 2");
    return IDFF_EUNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  const NetView& v = w->v;
  // particles per column tile: 128 / 64 / 32 for 1 / 2 / 3 jets; three-jet solves outside sweeps use the dense 40-particle geometry
  const bool dense3 = nj == 3 && d_coef == nullptr && g_wide_dense3;
  auto ppt_for = [&](int njk) { return (njk == 3 && dense3) ? kPpt3Dense : ppt_of(njk); };
  const int W = v.w, MT = W / 128, KB = W / 64, PPT = ppt_for(nj);
  // chunk of particles whose scratch stays around 3 GB
  // boundary evaluations of a one-jet field run in the two-jet tile geometry (64 particles per tile) so that they too
  // have a free column slot for the split-K accumulation (put_light): twice the column tiles for those evaluations
  // ... and boundary evaluations of a three-jet field outside sweeps need one jet + the split slot only: they too run in the
  // two-jet geometry (64 particles per tile, all 128 columns used) instead of 40 particles x 3 column slots of which one idles
  const int nj_light = (nj < 2 || dense3) ? 2 : nj;
  const int xmul = std::max(1, ppt_for(nj) / ppt_for(nj_light));
  // per tile: two X operand buffers, the stash [MT][groups][nj * HP][128] and the partial sums [MT][4][groups][2 * HP] (groups * HP = PPT)
  const size_t per_ctile = (size_t)2 * xmul * KB * XT_BYTES + (size_t)MT * nj * PPT * 128 * 4 + (size_t)MT * 4 * 2 * PPT * 4;
  long chunk_ctiles = (long)std::max<size_t>(1, ((size_t)3 << 30) / per_ctile);
  const long total_ctiles = (N + PPT - 1) / PPT;
  if (chunk_ctiles > total_ctiles) chunk_ctiles = total_ctiles;
  const long chunkP = chunk_ctiles * PPT;
  // (boundary evaluations may use a coarser tile geometry: their last tile pads to at most 127 rows beyond the chunk)
  const size_t part_slack = (size_t)MT * 4 * 2 * 128 * 4;
  const size_t bytes = (size_t)chunk_ctiles * per_ctile + part_slack + (size_t)chunkP * 2 * 4 * 8 + 4096;
  idff_weights_t* wm = const_cast<idff_weights_t*>(w);
  if (wm->stash_bytes < bytes) {
    if (wm->d_stash) cudaFree(wm->d_stash);
    wm->d_stash = nullptr;
    wm->stash_bytes = 0;
    IDFF_CUDA(cudaMalloc(&wm->d_stash, bytes));
    wm->stash_bytes = bytes;
  }
  WideScratch sc;
  {
    unsigned char* b = reinterpret_cast<unsigned char*>(wm->d_stash);
    sc.xa = b; b += (size_t)chunk_ctiles * xmul * KB * XT_BYTES;
    sc.xb = b; b += (size_t)chunk_ctiles * xmul * KB * XT_BYTES;
    sc.stash = reinterpret_cast<float*>(b); b += (size_t)chunk_ctiles * MT * nj * PPT * 128 * 4;
    sc.part = reinterpret_cast<float*>(b); b += (size_t)chunk_ctiles * MT * 4 * 2 * PPT * 4 + part_slack;
    float* f = reinterpret_cast<float*>(b);
    const size_t n2 = (size_t)chunkP * 2;
    sc.pred = f; sc.grad = f + n2; sc.y = f + 2 * n2; sc.k1 = f + 3 * n2; sc.k2 = f + 4 * n2; sc.k3 = f + 5 * n2;
    sc.xs = f + 6 * n2; sc.x0c = f + 7 * n2;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const int smem = NSTAGE * STAGE_BYTES + 256;
  auto set_attr = [&](auto kern) -> int {
    IDFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return IDFF_OK;
  };
  int rc = set_attr(k_wide_layer<1, 128>);
  if (!rc) rc = set_attr(k_wide_layer<2, 64>);
  if (!rc) rc = set_attr(k_wide_layer<3, 32>);
  if (!rc) rc = set_attr(k_wide_layer<3, kPpt3Dense>);
  if (rc) return rc;

  for (long c0 = 0; c0 < total_ctiles; c0 += chunk_ctiles) {
    const long nct = std::min(chunk_ctiles, total_ctiles - c0);
    const long row0 = c0 * PPT;
    const long Pn = std::min<long>(nct * PPT, N - row0);
    IDFF_CUDA(cudaMemcpyAsync(sc.y, d_x + row0 * 2, (size_t)Pn * 8, cudaMemcpyDeviceToDevice, s));
    IDFF_CUDA(cudaMemcpyAsync(sc.xs, d_x + row0 * 2, (size_t)Pn * 8, cudaMemcpyDeviceToDevice, s));
    if (mode == 0 && p->variant == IDFF_FIELD_CFM && d_x0)
      IDFF_CUDA(cudaMemcpyAsync(sc.x0c, d_x0 + row0 * 2, (size_t)Pn * 8, cudaMemcpyDeviceToDevice, s));
    for (int e = 0; e < n_eval; ++e) {
      const Stage& st = stages[e];
      const int heavy = (st.kind == 2 && rev) ? 1 : 0;
      const int njk = heavy ? nj : nj_light, PPTk = ppt_for(njk);
      const bool d3 = njk == 3 && PPTk == kPpt3Dense;
      const long nctk = (Pn + PPTk - 1) / PPTk;
      WParams P;
      memset(&P, 0, sizeof(P));
      P.net = v; P.w = W; P.MT = MT; P.KB = KB; P.nj = njk; P.ppt = PPTk; P.n_ctiles = (int)nctk; P.P = Pn; P.heavy = heavy;
      P.st = st; P.stash = sc.stash; P.part = sc.part; P.xs = sc.xs;
      P.coef = d_coef ? d_coef + (size_t)e * n_sets * 4 : nullptr; P.set_rows = set_rows; P.row0 = row0;
      const int ew_grid = (int)std::min<long>((nctk * W + 255) / 256, (long)sms * 8);
      P.xout = sc.xa;
      if (njk == 1) k_wide_l1<1, 128><<<ew_grid, 256, 0, s>>>(P);
      else if (njk == 2) k_wide_l1<2, 64><<<ew_grid, 256, 0, s>>>(P);
      else if (d3) k_wide_l1<3, kPpt3Dense><<<ew_grid, 256, 0, s>>>(P);
      else k_wide_l1<3, 32><<<ew_grid, 256, 0, s>>>(P);
      count_launch();
      const int ng = heavy ? 4 : 2;
      const int lgrid = (int)std::min<long>(nctk * MT, (long)sms);
      for (int g = 0; g < ng; ++g) {
        P.gemm = g;
        P.xin = (g & 1) ? sc.xb : sc.xa;
        P.xout = (g & 1) ? sc.xa : sc.xb;
        if (njk == 1) k_wide_layer<1, 128><<<lgrid, NTHREADS, smem, s>>>(P);
        else if (njk == 2) k_wide_layer<2, 64><<<lgrid, NTHREADS, smem, s>>>(P);
        else if (d3) k_wide_layer<3, kPpt3Dense><<<lgrid, NTHREADS, smem, s>>>(P);
        else k_wide_layer<3, 32><<<lgrid, NTHREADS, smem, s>>>(P);
        count_launch();
        if (g == 1 || g == 3) {
          float* dst = g == 1 ? sc.pred : sc.grad;
          const int sg_grid = (int)std::min<long>((Pn * 2 + 255) / 256, (long)sms * 8);
          if (njk == 1) k_wide_sum<1, 128><<<sg_grid, 256, 0, s>>>(P, dst);
          else if (njk == 2) k_wide_sum<2, 64><<<sg_grid, 256, 0, s>>>(P, dst);
          else if (d3) k_wide_sum<3, kPpt3Dense><<<sg_grid, 256, 0, s>>>(P, dst);
          else k_wide_sum<3, 32><<<sg_grid, 256, 0, s>>>(P, dst);
          count_launch();
        }
      }
      UParams U;
      memset(&U, 0, sizeof(U));
      U.P = Pn; U.N_total = N; U.row0 = row0; U.variant = p->variant; U.mode = mode; U.sg = e & 3; U.heavy = heavy;
      U.last = (e == n_eval - 1); U.dt = mode == 1 ? dts[e >> 2] : 0.0f; U.st = st;
      U.pred = sc.pred; U.grad = sc.grad; U.b3 = v.b3;
      U.coef = P.coef; U.set_rows = set_rows;
      U.y = sc.y; U.k1 = sc.k1; U.k2 = sc.k2; U.k3 = sc.k3; U.xs = sc.xs; U.x0c = sc.x0c;
      U.out = d_out + row0 * 2;
      U.traj_row = (mode == 1 && d_traj && (e & 3) == 3) ? d_traj + ((size_t)((e >> 2) + 1) * N + row0) * 2 : nullptr;
      k_wide_update<<<(int)std::min<long>((Pn * 2 + 255) / 256, (long)sms * 8), 256, 0, s>>>(U);
      count_launch();
    }
  }
  IDFF_CUDA(cudaGetLastError());
  return verify_tc_launch(wm, stream, 2);
}

int wide_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, const float* d_x0,
                    float* d_out, int64_t N, void* stream) {
  Stage st;
  make_stage(*p, t, &st);
  return wide_run(w, p, 0, d_x, d_x0, &st, nullptr, 1, d_out, nullptr, N, stream);
}

int wide_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                    const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream) {
  return wide_run(w, p, 1, d_x0, nullptr, stages, dts, 4 * n_iv, d_out_final, d_out_traj, N, stream);
}

// idff_sample_rk4_sweep on the layered engine: n_sets parameter sets (same variant, same jets_needed, same t grid, sigma
// and end_div -- only the folded gamma coefficients and a = sqrt(1 - Gamma sigma_t^2) differ) over the same x0, solved as ONE
// batch of n_sets * N particles: d_out [n_sets][N][2] first receives n_sets copies of x0 and is then integrated in place, the
// kernels look the per-set coefficients up by particle row.  N must be a multiple of 128 (the largest column tile), so that
// a column tile never straddles two sets.  Per particle the arithmetic is that of wide_sample_rk4 with the set's parameters.
__global__ void k_wide_replicate(const float2* __restrict__ x0, float2* __restrict__ out, long N, long total) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) out[i] = x0[i % N];
}

bool wide_sweep_supports(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, int64_t N) {
  if (N % 128 != 0 || n_sets < 2) return false;
  int nj0, rev0;
  jets_needed(params[0], &nj0, &rev0);
  for (int g = 0; g < n_sets; ++g) {
    const idff_field_params_t& q = params[g];
    int nj, rev;
    jets_needed(q, &nj, &rev);
    if (!wide_supports(w, &q) || q.variant != IDFF_FIELD_MOMENTUM || q.variant != params[0].variant || nj != nj0 || rev != rev0 ||
        q.sigma != params[0].sigma || q.end_div != params[0].end_div || q.order != params[0].order)
      return false;
  }
  return true;
}

int wide_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, const float* d_x0,
                          const Stage* stages /*[n_sets][4 n_iv]*/, const float* dts, int n_iv, float* d_out, int64_t N,
                          void* stream) {
  const int n_eval = 4 * n_iv;
  std::vector<float> coef((size_t)n_eval * n_sets * 4);
  for (int e = 0; e < n_eval; ++e)
    for (int g = 0; g < n_sets; ++g) {
      const Stage& st = stages[(size_t)g * n_eval + e];
      float* c = &coef[((size_t)e * n_sets + g) * 4];
      c[0] = st.c[0]; c[1] = st.c[1]; c[2] = st.c[2]; c[3] = st.a;
    }
  idff_weights_t* wm = const_cast<idff_weights_t*>(w);
  const size_t bytes = coef.size() * sizeof(float);
  if (wm->sweep_bytes < bytes) {
    if (wm->d_sweep) cudaFree(wm->d_sweep);
    wm->d_sweep = nullptr;
    wm->sweep_bytes = 0;
    IDFF_CUDA(cudaMalloc(&wm->d_sweep, bytes));
    wm->sweep_bytes = bytes;
  }
  cudaStream_t s = (cudaStream_t)stream;
  IDFF_CUDA(cudaMemcpyAsync(wm->d_sweep, coef.data(), bytes, cudaMemcpyHostToDevice, s));
  IDFF_CUDA(cudaStreamSynchronize(s));   // `coef` is pageable host memory that dies with this call
  const long total = (long)n_sets * N;
  k_wide_replicate<<<(int)std::min<long>((total + 255) / 256, 148 * 8), 256, 0, s>>>(
      reinterpret_cast<const float2*>(d_x0), reinterpret_cast<float2*>(d_out), (long)N, total);
  count_launch();
  return wide_run(w, &params[0], 1, d_out, nullptr, stages, dts, n_eval, d_out, nullptr, total, stream,
                  reinterpret_cast<const float*>(wm->d_sweep), n_sets, (long)N);
}

}  // namespace idff
