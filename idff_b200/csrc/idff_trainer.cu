// The 2-D IDFF training step as ONE persistent cooperative kernel.
//
// Reference step (2D_examples/Autograd_example_order2.py:89-112, identical in Autograd_example_order3.py:90-109):
//   t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1)        conditional_flow_matching.py:153-193
//   vt   = base_model(cat[xt, t])                                           MLP 3-256-256-256-2, models.py:4-21
//   loss = mean(((1/(1e-2+1-t)) * (vt - FM.x1))**2)                         :105
//   loss.backward(); clip_grad_norm_(params, 1); AdamW(lr=2e-4).step()      :109-112
//
// At the reference's batch (256) this is 0.2 GFLOP: launch- and latency-bound on any GPU (60 torch kernels, 290 us even as
// a CUDA graph).  Here `n_steps` consecutive steps run inside one launch; a step is three phases separated by a grid
// barrier (monotonic counter, one red.release per CTA; the launch is cooperative so all CTAs are co-resident):
//   A  rows:    B/4 CTAs own 4 batch rows each: path sample -> forward -> loss -> backward-data for those rows, no cross-CTA
//               traffic.  The four 256x256 weight matrices of the step (W1^T, W2^T, W2, W1: 1 MB, L2 resident) stream
//               through a 4 x 32 KB shared-memory ring filled by a producer warp with TMA bulk copies (cp.async.bulk +
//               mbarrier complete_tx), running ahead across layer boundaries.  A consumer thread owns 4 units x 4 rows over
//               one K-slice (128-bit conflict-free weight reads, broadcast activation reads); slices are summed in a fixed
//               order.  Activations and the pre-activation gradients dZ go to global memory (L2-resident, 1.5 MB).  The
//               producer warp also prefetches the next step's input rows.  (An opt-in variant runs a row group on a CTA
//               PAIR -- cluster of 2, each CTA streaming / computing half of the units, activation halves exchanged
//               through distributed shared memory with a cluster-scope mbarrier per layer: measured slower, kept as a
//               diagnostic)
//   B  weights: 145 jobs, one per CTA: 2 x 64 tiles (32 x 32) of dW1 / dW2 = dZ^T A (K = batch, operands staged in shared
//               memory, 4 x 4 outputs per thread over one eighth of K), 8 + 8 jobs for the thin layers, one for db3 and the
//               loss value.  A gradient never leaves the CTA that owns its parameter; only the CTA's sum of squares goes
//               to global memory.  The owner's P / exp_avg / exp_avg_sq loads are issued before the barrier.
//   C  update:  every CTA adds the partial sums of squares in the same fixed order (-> identical clip coefficient), then
//               each thread applies clip + AdamW (torch's op order) to the parameters it owns; the streaming copies of
//               W1 / W2 (forward = transposed, backward = as stored, each split into output-unit halves) are refreshed,
//               the transposed ones through a shared-memory transpose (coalesced stores).
// Every reduction has a fixed order: results are bit-reproducible run to run.
#include <cmath>

#include "idff_common.cuh"

namespace idff {
namespace trn {

constexpr int W = 256, RMIN = 4, RMAX = 4, NC = 512, NT = NC + 32, MAXB = IDFF_TRAIN_MAX_BATCH;   // RMIN / RMAX: batch rows per row group (template R of the kernels; both 4 -- 8 rows per CTA pair measured slower in both rounds: 33.4 vs 30.1 us at B = 512 with the st.async hand-over, profiles/r2_trainer_pair8_rate.txt)
constexpr int MDOUT = 4, NSLOT = 6;    // widest output layer served; parameters a thread can own
constexpr int JOB_W0 = 128, JOB_W3 = 136, JOB_B3 = 144, N_JOBS = 145;
constexpr int MAX_GRID = 256;
constexpr int NSMAX = 8, CH_ROWS = 32, CH_FLOATS = CH_ROWS * W, CH_PER_PASS = W / CH_ROWS, CH_PER_STEP = 4 * CH_PER_PASS;
constexpr int HALF = W * 128;          // floats of one output-unit half of a streaming copy: [256 reduce rows][128 units]
constexpr int HCH = CH_ROWS * 128;     // floats of one half-chunk (32 rows x 128 units = 16 KB)
static_assert(CH_PER_STEP % NSMAX == 0, "ring phase bookkeeping");

// clock64 phase stamps of the last step of the last launch, per CTA: diagnostic only (idff_debug_train_stamps)
__device__ long long g_train_stamps[256][8];

// The two networks the 2-D scripts train.  NET 0: MLP(dim=2): cat[xt, t] (3) -> 256 -> 256 -> 256 -> 2, flow loss
// (Autograd_example_order2.py:79,98-105).  NET 1: the score-head MLP(dim=4) of the order-1 script: cat[xt, xt, t] (5) -> ... -> 4,
// outputs (vt | st), flow loss + score regression (example_order1_with_prediction_head.py:101,119-132).
// Parameter block (nn.Linear layout) and offsets (floats) into the caller-owned state buffer.
// streaming copies [2 halves][256][128]: F_l[h][k][jj] = W_l[128 h + jj][k] (forward), B_l[h][j][kk] = W_l[j][128 h + kk] (backward)
template <int DIN_, int DOUT_>
struct LayT {
  static constexpr int DIN = DIN_, DOUT = DOUT_;
  static constexpr int oW0 = 0, ob0 = oW0 + W * DIN, oW1 = ob0 + W, ob1 = oW1 + W * W, oW2 = ob1 + W, ob2 = oW2 + W * W,
                       oW3 = ob2 + W, ob3 = oW3 + DOUT * W, NP = ob3 + DOUT;
  static constexpr int NPP = (NP + 63) / 64 * 64;
  static constexpr size_t sP = 0, sM = sP + NPP, sV = sM + NPP, sF1 = sV + NPP, sF2 = sF1 + W * W, sB1 = sF2 + W * W, sB2 = sB1 + W * W,
                          sIN = sB2 + W * W,
                          sA1 = sIN + (size_t)MAXB * 4, sA2 = sA1 + (size_t)MAXB * W, sA3 = sA2 + (size_t)MAXB * W,
                          sZ1 = sA3 + (size_t)MAXB * W, sZ2 = sZ1 + (size_t)MAXB * W, sZ3 = sZ2 + (size_t)MAXB * W,
                          sDO = sZ3 + (size_t)MAXB * W, sLP = sDO + (size_t)MAXB * DOUT, sNP = sLP + MAX_GRID, sBAR = sNP + MAX_GRID,
                          sTOTAL = sBAR + 64;
};
template <int NET> struct Lay;
template <> struct Lay<0> : LayT<3, 2> {};
template <> struct Lay<1> : LayT<5, 4> {};

// dynamic shared memory (floats).  Phase A: ring | act | part | raw | in4 | dout | lterm.  Phases B / C alias it.
constexpr int mRING = 0, mACT = mRING + 4 * CH_FLOATS, mPART = mACT + 2 * RMAX * W, mRAW = mPART + 2048 * RMAX,
              mIN4 = mRAW + 2 * RMAX * 8, mDOUT = mIN4 + RMAX * 4, mLT = mDOUT + RMAX * MDOUT, mTOTAL = mLT + RMAX * MDOUT;
constexpr int mDZS = 0, mAS = mDZS + 256 * 32, mRED = mAS + 256 * 32, mREDB = mRED + 8 * 16 * 64, mPT = mREDB + 16 * 32,
              mBTOTAL = mPT + 32 * 33;
static_assert(mBTOTAL <= mRAW, "phase B scratch must not overlap the input prefetch buffer");
constexpr size_t kSmemBytes = (size_t)mTOTAL * sizeof(float);

struct Args {
  float* S;   // state
  const float *x0, *x1, *t, *eps;
  float *loss_out, *gnorm_out, *grad_dump;
  int B, n_steps;
  long long steps_done;
  float sigma;
  int sigma_mode;
  float score_sigma;   // NET 1: the sigma0_sq argument of log_prob_xt_given_x1_dimwise (the script passes sigma, :126)
  double lr, beta1, beta2, b1pow, b2pow;            // bias corrections in double like torch's Python scalars; b?pow = beta?^steps_done
  float decay, omb1, omb2, beta2f, eps_adam, clip;  // 1 - lr*wd, 1 - beta1, 1 - beta2 (rounded to fp32 from double)
};

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
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(ok)
        : "r"(s32(b)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 consumer warps
__device__ __forceinline__ void fullsync() { asm volatile("bar.sync 0, 544;" ::: "memory"); }   // + the producer warp

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(s32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// remote store that signals `bytes` on the destination CTA's mbarrier when it has landed (SASS STAS): data and hand-over in one
__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(addr), "f"(v), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(ok)
        : "r"(s32(b)), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ float sigma_t_of(float t, float sigma, int mode) {
  if (mode == 0) return sigma;
  return __fmul_rn(sigma, __fsqrt_rn(__fmul_rn(t, __fsub_rn(1.0f, t))));   // conditional_flow_matching.py:231-233
}

// Grid barrier for the consumer warps: one release-add per CTA on a monotonic counter, then poll.  `with_producer`: the
// closing CTA barrier also admits the producer warp (top of the next step).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch, bool with_producer) {
  csync();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < epoch);
  }
  if (with_producer) fullsync(); else csync();
}

__device__ __forceinline__ float selu_d1(float code) { return code < 0.0f ? kSeluLambda : code; }

// Per-CTA geometry of the rows phase.  CL = 1: the CTA computes all 256 units of its R rows; CL = 2: 128 of them.
template <int CL, int R>
struct Geo {
  static constexpr int UPC = W / CL;          // output units per CTA
  static constexpr int NU4 = UPC / 4;         // threads across units (4 units each)
  static constexpr int KSL = NC / NU4;        // K-slices: 8 / 16
  static constexpr int RPS = CH_ROWS / KSL;   // k rows of a chunk per slice: 4 / 2
  static constexpr int NQ = R * UPC / NC;     // outputs (row, unit) finalised per thread
  static constexpr int CHF = CH_FLOATS / CL;  // floats of one ring chunk
  static constexpr int NS = 4 * CL;           // ring stages (128 KB either way)
};

// One 256 x 256 layer pass over the ring: thread (ks, u4) accumulates units 4 u4 .. 4 u4 + 3 (of this CTA's units) of the 4 rows
// over the k rows {32 c + RPS ks + i} of every chunk c, then the K-slices meet in `part` [KSL][R][UPC].
template <int CL, int R>
__device__ __forceinline__ void gemm_ring(const float* ring, const float* act, float* part, uint64_t* full, uint64_t* empty,
                                          uint32_t& cc, int ks, int u4, int lane) {
  using G = Geo<CL, R>;
  float acc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
  const int woff = (u4 >> 5) * HCH + (G::RPS * ks) * 128 + 4 * (u4 & 31);   // [half][32 rows][128 units] within a chunk
#pragma unroll 1
  for (int c = 0; c < CH_PER_PASS; ++c) {
    const uint32_t stage = cc % G::NS;
#ifndef IDFF_TRN_NOTMA   // (diagnostic builds: -DIDFF_TRN_NOTMA = compute on whatever is in the ring, -DIDFF_TRN_NOFMA = stream only)
    mbar_wait(&full[stage], (cc / G::NS) & 1);
#endif
    const float* wb = ring + stage * G::CHF + woff;
#ifndef IDFF_TRN_NOFMA
    float av[R][G::RPS];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (G::RPS == 4) {
        const float4 v = *reinterpret_cast<const float4*>(act + r * W + c * CH_ROWS + 4 * ks);
        av[r][0] = v.x; av[r][1] = v.y; av[r][G::RPS - 2] = v.z; av[r][G::RPS - 1] = v.w;
      } else {
        const float2 v = *reinterpret_cast<const float2*>(act + r * W + c * CH_ROWS + 2 * ks);
        av[r][0] = v.x; av[r][1] = v.y;
      }
    }
#pragma unroll
    for (int i = 0; i < G::RPS; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(wb + i * 128);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][0] = fmaf(av[r][i], w.x, acc[r][0]);
        acc[r][1] = fmaf(av[r][i], w.y, acc[r][1]);
        acc[r][2] = fmaf(av[r][i], w.z, acc[r][2]);
        acc[r][3] = fmaf(av[r][i], w.w, acc[r][3]);
      }
    }
#endif
    __syncwarp();
#ifndef IDFF_TRN_NOTMA
    if (lane == 0) mbar_arrive(&empty[stage]);
#endif
    ++cc;
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
    *reinterpret_cast<float4*>(part + (ks * R + r) * G::UPC + 4 * u4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
}

template <int CL, int R>
__device__ __forceinline__ float part_sum(const float* part, int r, int jl) {
  using G = Geo<CL, R>;
  float z = part[r * G::UPC + jl];
#pragma unroll
  for (int k = 1; k < G::KSL; ++k) z += part[(k * R + r) * G::UPC + jl];
  return z;
}

// raw input rows of one step -> shared memory: {x0[0], x0[1], x1[0], x1[1], eps[0], eps[1], t, 0}
__device__ __forceinline__ void load_raw(const Args& a, int s, int row_in_cta, int row0, int R, float* raw) {
  const size_t row = (size_t)s * a.B + row0 + row_in_cta;
  const float2 x0 = *reinterpret_cast<const float2*>(a.x0 + row * 2), x1 = *reinterpret_cast<const float2*>(a.x1 + row * 2);
  const float2 e = a.eps ? *reinterpret_cast<const float2*>(a.eps + row * 2) : make_float2(0.f, 0.f);
  float* o = raw + ((s & 1) * R + row_in_cta) * 8;
  *reinterpret_cast<float4*>(o) = make_float4(x0.x, x0.y, x1.x, x1.y);
  *reinterpret_cast<float4*>(o + 4) = make_float4(e.x, e.y, a.t[row], 0.f);
}

// ---- phase A (consumer warps): path sample, forward, loss, backward-data for the 4 rows of this CTA (pair) ------------------
template <int NET, int CL, int R>
__device__ void phase_rows(const Args& a, int s, float* smem, int grp, uint32_t rank, uint64_t* full, uint64_t* empty,
                           uint64_t* actbar, uint32_t& aph, uint32_t& cc) {
  using G = Geo<CL, R>;
  using L = Lay<NET>;
  const float* ring = smem + mRING;
  float* actb = smem + mACT;      // [2][R][W]: layer n reads buffer n & 1 and writes the other one
  float* part = smem + mPART;     // [KSL][R][UPC]
  const float* raw = smem + mRAW + (s & 1) * R * 8;
  float* in4 = smem + mIN4;       // [R][4]   xt0, xt1, t, 1
  float* dout = smem + mDOUT;     // [R][DOUT]
  float* lterm = smem + mLT;      // [R*DOUT]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = tid / G::NU4, u4 = tid % G::NU4;
  const int row0 = grp * R;
  const float* P = a.S + L::sP;
  const uint32_t peer_act = CL == 2 ? mapa(actb, rank ^ 1) : 0, peer_bar = CL == 2 ? mapa(actbar, rank ^ 1) : 0;

  // the (row, unit) outputs this thread finalises in every layer, and its share of the thin layers / biases (issued now)
  int orow[G::NQ], oj[G::NQ], ojl[G::NQ];
  float w0c[G::NQ][L::DIN], bias0[G::NQ], bias1[G::NQ], bias2[G::NQ], w3c[G::NQ][L::DOUT];
#pragma unroll
  for (int q = 0; q < G::NQ; ++q) {
    const int o = tid + q * NC;
    orow[q] = o / G::UPC; ojl[q] = o % G::UPC; oj[q] = (int)rank * G::UPC + ojl[q];
    const int j = oj[q];
#pragma unroll
    for (int c = 0; c < L::DIN; ++c) w0c[q][c] = __ldcg(P + L::oW0 + j * L::DIN + c);
    bias0[q] = __ldcg(P + L::ob0 + j); bias1[q] = __ldcg(P + L::ob1 + j); bias2[q] = __ldcg(P + L::ob2 + j);
#pragma unroll
    for (int c = 0; c < L::DOUT; ++c) w3c[q][c] = __ldcg(P + L::oW3 + c * W + j);
  }
  // publish one value of the next layer's input: own copy + (pair) the peer's copy through distributed shared memory
  // (pair: the remote copy is a st.async that signals the peer's hand-over barrier of this exchange: actbar[aph & 1])
  auto put = [&](int buf, int r, int j, float v) {
    actb[(buf * R + r) * W + j] = v;
    if (CL == 2) st_async_f32(peer_act + ((buf * R + r) * W + j) * 4, v, peer_bar + 8u * (aph & 1u));
  };
  // all 256 units of the next layer's input are in place in this CTA (and, pair, this CTA's half in the peer)
  auto layer_sync = [&]() {
    if (CL == 1) {
      csync();
    } else {
      // exchange number aph: barrier aph & 1, phase (aph >> 1) & 1.  The peer's half (R x 128 values) arrives as st.async stores
      // that complete_tx on that barrier; thread 0 armed it with the byte count two exchanges ago and re-arms it right after
      // the wait (the peer cannot push exchange aph + 2 before it has received this CTA's exchange aph + 1).
      mbar_wait(&actbar[aph & 1u], (aph >> 1) & 1u);
      if (tid == 0) mbar_arrive_expect_tx(&actbar[aph & 1u], (uint32_t)(R * (W / 2) * 4));
      ++aph;
      csync();   // own half: the local stores of every thread
    }
  };

  if (tid < R) {   // K2 inlined: xt = t*x1 + (1-t)*x0 + sigma_t*eps   (conditional_flow_matching.py:98-123, no contraction)
    const float4 q0 = *reinterpret_cast<const float4*>(raw + tid * 8), q1 = *reinterpret_cast<const float4*>(raw + tid * 8 + 4);
    const float t = q1.z;
    const float st = sigma_t_of(t, a.sigma, a.sigma_mode);
    const float omt = __fsub_rn(1.0f, t);
    const float xt0 = __fadd_rn(__fadd_rn(__fmul_rn(t, q0.z), __fmul_rn(omt, q0.x)), __fmul_rn(st, q1.x));
    const float xt1 = __fadd_rn(__fadd_rn(__fmul_rn(t, q0.w), __fmul_rn(omt, q0.y)), __fmul_rn(st, q1.y));
    const float4 v = make_float4(xt0, xt1, t, 1.0f);
    *reinterpret_cast<float4*>(in4 + tid * 4) = v;
    if (rank == 0) *reinterpret_cast<float4*>(a.S + L::sIN + (size_t)(row0 + tid) * 4) = v;
  }
  csync();

  float code1[G::NQ], code2[G::NQ], code3[G::NQ];
#pragma unroll
  for (int q = 0; q < G::NQ; ++q) {   // layer 1 (K = 3; NET 1: K = 5, the input is cat[xt, xt, t]) -> buffer 0
    const int r = orow[q];
    float z = fmaf(in4[r * 4 + 1], w0c[q][1], fmaf(in4[r * 4], w0c[q][0], bias0[q]));
    if (NET == 1) z = fmaf(in4[r * 4 + 1], w0c[q][3], fmaf(in4[r * 4], w0c[q][2], z));
    z = fmaf(in4[r * 4 + 2], w0c[q][L::DIN - 1], z);
    const float av = selu_fwd(z, code1[q]);
    put(0, r, oj[q], av);
    a.S[L::sA1 + (size_t)(row0 + r) * W + oj[q]] = av;
  }
  layer_sync();

#pragma unroll 1
  for (int l = 0; l < 2; ++l) {   // layers 2, 3: ring passes 0 (W1^T), 1 (W2^T); reads buffer l, writes buffer l ^ 1
    gemm_ring<CL, R>(ring, actb + l * R * W, part, full, empty, cc, ks, u4, lane);
    csync();
    float* Aout = a.S + (l ? L::sA3 : L::sA2);
#pragma unroll
    for (int q = 0; q < G::NQ; ++q) {
      const int r = orow[q];
      const float z = part_sum<CL, R>(part, r, ojl[q]) + (l ? bias2[q] : bias1[q]);
      float c;
      const float av = selu_fwd(z, c);
      if (l) code3[q] = c; else code2[q] = c;
      put(l ^ 1, r, oj[q], av);
      Aout[(size_t)(row0 + r) * W + oj[q]] = av;
    }
    layer_sync();
  }

  // output layer + loss (Autograd_example_order2.py:105) + d loss / d vt; a3 is in buffer 0 (both CTAs of a pair compute it)
  if (warp < R * L::DOUT) {
    const int r = warp / L::DOUT, o = warp % L::DOUT;
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < W / 32; ++i) sum = fmaf(actb[r * W + lane + 32 * i], __ldcg(P + L::oW3 + o * W + lane + 32 * i), sum);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) {
      const float vt = sum + __ldcg(P + L::ob3 + o);
      const float t = in4[r * 4 + 2];
      const float gk = 1.0f / (float)(a.B * 2);   // both loss terms are means over B x 2 elements
      float d;
      if (o < 2) {   // flow loss, Autograd_example_order2.py:105 (= example_order1_with_prediction_head.py:129)
        const float x1v = raw[r * 8 + 2 + o];
        const float wgt = __fdiv_rn(1.0f, __fsub_rn(1.01f, t));
        const float rr = __fmul_rn(wgt, __fsub_rn(vt, x1v));
        lterm[warp] = __fmul_rn(rr, rr);
        d = 2.0f * rr * wgt * gk;
      } else {       // score regression mean((st - logp)^2), logp = log_prob_xt_given_x1_dimwise(xt, x1, x0, t, sigma) (:26-51, :126-131;
                     // the operation order of k_loss<1>, idff_train.cu)
        const int dd = o - 2;
        const float x0v = raw[r * 8 + dd], x1v = raw[r * 8 + 2 + dd], xtv = in4[r * 4 + dd];
        const float omt = __fsub_rn(1.0f, t);
        const float mu = __fadd_rn(__fmul_rn(t, x1v), __fmul_rn(omt, x0v));
        const float s2v = __fadd_rn(__fmul_rn(__fmul_rn(a.score_sigma, t), omt), 1e-4f);
        const float diff = __fsub_rn(xtv, mu);
        const float lp = __fmul_rn(-0.5f, __fadd_rn(logf(__fmul_rn(6.283185307179586f, s2v)), __fdiv_rn(__fmul_rn(diff, diff), s2v)));
        const float rr = __fsub_rn(vt, lp);
        lterm[warp] = __fmul_rn(rr, rr);
        d = 2.0f * rr * gk;
      }
      dout[r * L::DOUT + o] = d;
      if (rank == 0) a.S[L::sDO + (size_t)(row0 + r) * L::DOUT + o] = d;
    }
  }
  csync();
  if (tid == 0 && rank == 0) {
    float ls = 0.0f;
#pragma unroll
    for (int i = 0; i < R * L::DOUT; ++i) ls += lterm[i];
    a.S[L::sLP + grp] = ls;
  }
#pragma unroll
  for (int q = 0; q < G::NQ; ++q) {   // dZ3 = (dOut W3) * selu'(z3) -> buffer 1
    const int r = orow[q];
    float dsum = fmaf(dout[r * L::DOUT + 1], w3c[q][1], dout[r * L::DOUT] * w3c[q][0]);
    if (NET == 1) dsum = fmaf(dout[r * L::DOUT + 3], w3c[q][3], fmaf(dout[r * L::DOUT + 2], w3c[q][2], dsum));
    const float dz = dsum * selu_d1(code3[q]);
    put(1, r, oj[q], dz);
    a.S[L::sZ3 + (size_t)(row0 + r) * W + oj[q]] = dz;
  }
  layer_sync();
#pragma unroll 1
  for (int l = 0; l < 2; ++l) {   // dZ2 = (dZ3 W2) * selu'(z2), dZ1 = (dZ2 W1) * selu'(z1): ring passes 2 (W2), 3 (W1)
    gemm_ring<CL, R>(ring, actb + (l ^ 1) * R * W, part, full, empty, cc, ks, u4, lane);
    csync();
    float* Zout = a.S + (l ? L::sZ1 : L::sZ2);
#pragma unroll
    for (int q = 0; q < G::NQ; ++q) {
      const int r = orow[q];
      const float dz = part_sum<CL, R>(part, r, ojl[q]) * selu_d1(l ? code1[q] : code2[q]);
      if (l == 0) put(0, r, oj[q], dz);
      Zout[(size_t)(row0 + r) * W + oj[q]] = dz;
    }
    if (l == 0) layer_sync();
  }
}

// One thread's share of the parameters: up to 4 gradients and their flat indices (-1: none).
struct Owned {
  float g[NSLOT];
  int idx[NSLOT];
  int tile;     // tile jobs: the job index (layer, jb, kb) whose streaming copies this CTA refreshes, else -1
};

// ---- phase B ---------------------------------------------------------------------------------------------------------------
// tile (jb, kb) of dW = dZ^T A: thread (ksl, jj, kk) accumulates the 4 x 4 block (4 jj.., 4 kk..) over rows r = ksl (mod 8);
// the 8 partial blocks meet in shared memory; thread (j = tid / 16, kp = tid % 16) ends up owning (j, 2 kp), (j, 2 kp + 1).
template <int NET>
__device__ void job_tile(const Args& a, int job, float* smem, Owned& own) {
  using L = Lay<NET>;
  float* dZs = smem + mDZS;     // [256][32]
  float* As = smem + mAS;       // [256][32]
  float* red = smem + mRED;     // [8][16][64]
  float* redb = smem + mREDB;   // [16][32]
  const int tid = threadIdx.x;
  const int m = job >> 6, tile = job & 63, jb = tile >> 3, kb = tile & 7;
  const float* dZ = a.S + (m ? L::sZ3 : L::sZ2);
  const float* Ain = a.S + (m ? L::sA2 : L::sA1);
  const int ksl = tid >> 6, t64 = tid & 63, jj = t64 >> 3, kk = t64 & 7;
  float acc[4][4], gbs = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
  // The two operand panels (dZ and A, 32 columns each) move through shared memory in chunks of 128 batch rows, double buffered: the
  // loads of chunk c + 1 are in flight while chunk c is multiplied (all 128 tile jobs ask the L2 for their 64 KB at the same moment
  // right after the grid barrier).  Per thread the rows are still visited in ascending order: the sums keep their order.
#ifndef IDFF_TRN_CR
#define IDFF_TRN_CR 128
#endif
  constexpr int CR = IDFF_TRN_CR, NLD = CR * 8 / NC;   // rows per chunk (buffer b of a panel = its rows [b * CR, b * CR + CR)); 128-bit loads per thread and panel
  const int nch = (a.B + CR - 1) / CR;
  float4 vz[NLD], va[NLD];
  auto fetch = [&](int c) {
    const int rows = min(CR, a.B - c * CR);
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * NC, r = i >> 3, c4 = i & 7;
      if (r < rows) {
        vz[u] = __ldcg(reinterpret_cast<const float4*>(dZ + (size_t)(c * CR + r) * W + jb * 32) + c4);
        va[u] = __ldcg(reinterpret_cast<const float4*>(Ain + (size_t)(c * CR + r) * W + kb * 32) + c4);
      }
    }
  };
  fetch(0);
  float bsum = 0.f;
  for (int c = 0; c < nch; ++c) {
    const int rows = min(CR, a.B - c * CR);
    float* dzb = dZs + (c & 1) * CR * 32;
    float* ab = As + (c & 1) * CR * 32;
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * NC, r = i >> 3, c4 = i & 7;
      if (r < rows) {
        *reinterpret_cast<float4*>(dzb + r * 32 + c4 * 4) = vz[u];
        *reinterpret_cast<float4*>(ab + r * 32 + c4 * 4) = va[u];
      }
    }
    if (c + 1 < nch) fetch(c + 1);
    csync();   // chunk c is staged; everybody has finished chunk c - 1, whose buffer the NEXT iteration overwrites
#pragma unroll 4
    for (int r = ksl; r < rows; r += 8) {
      const float4 dz = *reinterpret_cast<const float4*>(dzb + r * 32 + 4 * jj);
      const float4 av = *reinterpret_cast<const float4*>(ab + r * 32 + 4 * kk);
      const float d[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(d[i], av.x, acc[i][0]);
        acc[i][1] = fmaf(d[i], av.y, acc[i][1]);
        acc[i][2] = fmaf(d[i], av.z, acc[i][2]);
        acc[i][3] = fmaf(d[i], av.w, acc[i][3]);
      }
    }
    if (kb == 0) {   // column sums of dZ: the bias gradient (tiles of the first k-block only); thread = (16 row groups, column);
      const int cc = tid & 31, rg = tid >> 5;   // one running sum per 256 batch rows, as before the panels were chunked
      for (int r = rg; r < rows; r += 16) bsum += dzb[r * 32 + cc];
      if ((c + 1) % (256 / CR) == 0 || c + 1 == nch) { gbs += bsum; bsum = 0.f; }
    }
  }
  csync();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) red[(ksl * 16 + i * 4 + c) * 64 + t64] = acc[i][c];
  if (kb == 0) redb[tid] = gbs;   // [16 row groups][32 columns]
  csync();
  {
    const int j = tid >> 4, kp = tid & 15;
    const int src = (j >> 2) * 8 + (kp >> 1), e = (j & 3) * 4 + (kp & 1) * 2;
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g0 += red[(k * 16 + e) * 64 + src];
      g1 += red[(k * 16 + e + 1) * 64 + src];
    }
    const int oW = m ? L::oW2 : L::oW1;
    own.g[0] = g0; own.g[1] = g1;
    own.idx[0] = oW + (jb * 32 + j) * W + kb * 32 + 2 * kp;
    own.idx[1] = own.idx[0] + 1;
    own.tile = job;
    if (kb == 0 && kp == 0) {
      float b = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) b += redb[k * 32 + j];
      own.g[2] = b;
      own.idx[2] = (m ? L::ob2 : L::ob1) + jb * 32 + j;
    }
  }
}

// thin layers: G[j][c] = sum_r X[r][j] * Y[r][c]; X = dZ1, Y = IN (c < 4, IN[r][3] = 1 -> db0) or X = A3, Y = dOut (c < 2)
template <int C>
__device__ void job_thin(const Args& a, const float* X, const float* Y, int q, float* smem, float (&out)[4]) {
  const int tid = threadIdx.x, jl = tid & 31, rg = tid >> 5, j = 32 * q + jl;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int r = rg; r < a.B; r += 16) {
    const float x = __ldcg(X + (size_t)r * W + j);
    if (C == 4) {
      const float4 y = __ldcg(reinterpret_cast<const float4*>(Y) + r);
      acc[0] = fmaf(x, y.x, acc[0]); acc[1] = fmaf(x, y.y, acc[1]); acc[2] = fmaf(x, y.z, acc[2]); acc[3] = fmaf(x, y.w, acc[3]);
    } else {
      const float2 y = __ldcg(reinterpret_cast<const float2*>(Y) + r);
      acc[0] = fmaf(x, y.x, acc[0]); acc[1] = fmaf(x, y.y, acc[1]);
    }
  }
  *reinterpret_cast<float4*>(smem + (rg * 32 + jl) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  csync();
  if (tid < 32) {
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = 0.0f;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      const float4 v = *reinterpret_cast<const float4*>(smem + (g * 32 + tid) * 4);
      out[0] += v.x; out[1] += v.y; out[2] += v.z; out[3] += v.w;
    }
  }
}

template <int NET, int CL, int R>
__global__ void __launch_bounds__(NT, 1) k_train_flow(const Args a) {
  using G = Geo<CL, R>;
  using L = Lay<NET>;
  constexpr int NS = G::NS;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t full[NSMAX], empty[NSMAX], actbar[2];
  __shared__ double sred[NC / 32];
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
  const uint32_t rank = CL == 2 ? cluster_ctarank() : 0;
  const int grp = cta / CL;              // row group: the CTA (pair) that owns batch rows [4 grp, 4 grp + 4)
  const bool row_cta = grp < a.B / R;
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NC / 32); }
    mbar_init(&actbar[0], 1);   // pair: hand-over barriers of the even / odd activation exchanges (armed by thread 0, completed
    mbar_init(&actbar[1], 1);   // by the bytes of the peer's st.async stores)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (row_cta && tid < R) load_raw(a, 0, tid, grp * R, R, smem + mRAW);
  __syncthreads();
  if (CL == 2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  if (CL == 2 && tid == 0 && row_cta) {
    mbar_arrive_expect_tx(&actbar[0], (uint32_t)(R * (W / 2) * 4));
    mbar_arrive_expect_tx(&actbar[1], (uint32_t)(R * (W / 2) * 4));
  }
  uint32_t cc = 0;    // ring chunks issued (producer) / consumed (consumers): the same sequence on both sides
  uint32_t aph = 0;   // activation exchanges done (pair mode)

  if (warp == NC / 32) {
    // ---- producer warp: streams this CTA's share of the step's four weight matrices through the ring (pair: the half that
    // produces its 128 output units; single: both halves), and prefetches the next step's input rows ----
    for (int s = 0; s < a.n_steps; ++s) {
      if (s > 0) fullsync();   // weights of step s are final (the consumers passed the closing grid barrier of step s - 1)
      if (!row_cta) continue;
#ifndef IDFF_TRN_NOTMA
      if (lane == 0) {
        fence_async_global();
        const float* src[4] = {a.S + L::sF1, a.S + L::sF2, a.S + L::sB2, a.S + L::sB1};
#pragma unroll 1
        for (int c = 0; c < CH_PER_STEP; ++c, ++cc) {
          const uint32_t stage = cc % NS;
          if (cc >= NS) mbar_wait(&empty[stage], ((cc / NS) - 1) & 1);
          mbar_arrive_expect_tx(&full[stage], G::CHF * 4);
          const float* m = src[c / CH_PER_PASS] + (size_t)(c % CH_PER_PASS) * HCH;
          float* dst = smem + mRING + stage * G::CHF;
          if (CL == 2) {
            tma_load_1d(dst, m + (size_t)rank * HALF, HCH * 4, &full[stage]);
          } else {
            tma_load_1d(dst, m, HCH * 4, &full[stage]);
            tma_load_1d(dst + HCH, m + HALF, HCH * 4, &full[stage]);
          }
        }
      }
#endif
      if (s + 1 < a.n_steps && lane < R) load_raw(a, s + 1, lane, grp * R, R, smem + mRAW);
    }
    return;
  }

  // ---- consumer warps ------------------------------------------------------------------------------------------------------
  unsigned epoch = 0;
  unsigned* bar = reinterpret_cast<unsigned*>(a.S + L::sBAR);
  float* P = a.S + L::sP;
  float* M = a.S + L::sM;
  float* V = a.S + L::sV;
  double b1pow = a.b1pow, b2pow = a.b2pow;

  for (int s = 0; s < a.n_steps; ++s) {
    const bool stamp = tid == 0 && s == a.n_steps - 1;
    if (stamp) g_train_stamps[cta][0] = clock64();
    if (row_cta) phase_rows<NET, CL, R>(a, s, smem, grp, rank, full, empty, actbar, aph, cc);
    if (stamp) g_train_stamps[cta][1] = clock64();
    grid_barrier(bar, epoch, false);
    if (stamp) g_train_stamps[cta][2] = clock64();

    Owned own;
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) { own.g[i] = 0.0f; own.idx[i] = -1; }
    own.tile = -1;
    if (cta < JOB_W0) {
      job_tile<NET>(a, cta, smem, own);
    } else if (cta < JOB_W3) {   // dW0, db0 for units [32 q, 32 q + 32): sums of dZ1 * (xt0, xt1, t, 1)
      float o[4];
      const int q = cta - JOB_W0;
      job_thin<4>(a, a.S + L::sZ1, a.S + L::sIN, q, smem, o);
      if (tid < 32) {
        const int j = 32 * q + tid;
        if (NET == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { own.g[c] = o[c]; own.idx[c] = L::oW0 + j * 3 + c; }
          own.g[3] = o[3]; own.idx[3] = L::ob0 + j;
        } else {   // input cat[xt, xt, t]: columns 0, 2 see xt0, columns 1, 3 see xt1
          own.g[0] = o[0]; own.g[1] = o[1]; own.g[2] = o[0]; own.g[3] = o[1]; own.g[4] = o[2]; own.g[5] = o[3];
#pragma unroll
          for (int c = 0; c < 5; ++c) own.idx[c] = L::oW0 + j * L::DIN + c;
          own.idx[5] = L::ob0 + j;
        }
      }
    } else if (cta < JOB_B3) {   // dW3 for units [32 q, 32 q + 32)
      float o[4];
      const int q = cta - JOB_W3;
      job_thin<L::DOUT>(a, a.S + L::sA3, a.S + L::sDO, q, smem, o);
      if (tid < 32) {
        const int j = 32 * q + tid;
#pragma unroll
        for (int c = 0; c < L::DOUT; ++c) { own.g[c] = o[c]; own.idx[c] = L::oW3 + c * W + j; }
      }
    } else if (cta == JOB_B3) {   // db3 and the loss value
      if (warp < L::DOUT) {
        float sum = 0.0f;
        for (int r = lane; r < a.B; r += 32) sum += __ldcg(a.S + L::sDO + (size_t)r * L::DOUT + warp);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) { own.g[0] = sum; own.idx[0] = L::ob3 + warp; }
      } else if (warp == 8) {
        double sum = 0.0;
        for (int c = lane; c < a.B / R; c += 32) sum += (double)__ldcg(a.S + L::sLP + c);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0 && a.loss_out) a.loss_out[s] = (float)(sum / (double)(a.B * 2));   // each loss term: a mean over B x 2 elements
      }
    }
    // the owner's parameter and moment loads do not depend on the clip coefficient: in flight across the barrier
    float pv[NSLOT], mv[NSLOT], vv[NSLOT];
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) {
      const int ix = own.idx[i];
      pv[i] = mv[i] = vv[i] = 0.0f;
      if (ix >= 0) { pv[i] = __ldcg(P + ix); mv[i] = __ldcg(M + ix); vv[i] = __ldcg(V + ix); }
    }
    {   // this CTA's sum of squared gradients
      double ss = 0.0;
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) ss += (double)own.g[i] * (double)own.g[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      if (lane == 0) sred[warp] = ss;
      csync();
      if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < NC / 32; ++i) tot += sred[i];
        a.S[L::sNP + cta] = (float)tot;
      }
    }
    if (stamp) g_train_stamps[cta][3] = clock64();
    grid_barrier(bar, epoch, false);
    if (stamp) g_train_stamps[cta][4] = clock64();

    b1pow *= a.beta1;
    b2pow *= a.beta2;
    if (warp == 0) {   // clip_grad_norm_(params, clip): coef = min(1, clip / (norm + 1e-6))
      double tot = 0.0;
#pragma unroll
      for (int i = 0; i < MAX_GRID / 32; ++i) {
        const int c = lane + 32 * i;
        tot += c < (int)gridDim.x ? (double)__ldcg(a.S + L::sNP + c) : 0.0;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
      if (lane == 0) {
        const float norm = (float)sqrt(tot);
        s_coef = a.clip > 0.0f ? fminf(1.0f, a.clip / (norm + 1e-6f)) : 1.0f;
        if (cta == 0 && a.gnorm_out) a.gnorm_out[s] = norm;
        s_step_size = (float)(a.lr / (1.0 - b1pow));
        s_bc2_sqrt = (float)sqrt(1.0 - b2pow);
      }
    }
    csync();
    {   // AdamW, torch/optim/adamw.py op order
      const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
      float pn[NSLOT];
#pragma unroll
      for (int i = 0; i < NSLOT; ++i) {
        pn[i] = 0.0f;
        const int ix = own.idx[i];
        if (ix < 0) continue;
        if (a.grad_dump && s == a.n_steps - 1) a.grad_dump[ix] = own.g[i];
        const float g = own.g[i] * coef;
        const float p = pv[i] * a.decay;
        const float mm = fmaf(g - mv[i], a.omb1, mv[i]);
        const float v2 = fmaf(a.omb2 * g, g, vv[i] * a.beta2f);
        const float denom = sqrtf(v2) / bc2_sqrt + a.eps_adam;
        pn[i] = p - step_size * (mm / denom);
        P[ix] = pn[i]; M[ix] = mm; V[ix] = v2;
      }
      if (own.tile >= 0) {   // streaming copies of this 32 x 32 tile: backward (as stored) and forward (transposed via smem)
        const int m = own.tile >> 6, jb = (own.tile >> 3) & 7, kb = own.tile & 7;
        float* Bc = a.S + (m ? L::sB2 : L::sB1) + (size_t)(kb >> 2) * HALF + (kb & 3) * 32;   // B[h][j][kk], kk = k - 128 h
        float* Fc = a.S + (m ? L::sF2 : L::sF1) + (size_t)(jb >> 2) * HALF + (jb & 3) * 32;   // F[h][k][jj], jj = j - 128 h
        float* pt = smem + mPT;
        const int j = tid >> 4, kp = tid & 15;
        *reinterpret_cast<float2*>(Bc + (jb * 32 + j) * 128 + 2 * kp) = make_float2(pn[0], pn[1]);
        pt[(2 * kp) * 33 + j] = pn[0];
        pt[(2 * kp + 1) * 33 + j] = pn[1];
        csync();
        const int k = tid >> 4, j2 = (tid & 15) * 2;
        *reinterpret_cast<float2*>(Fc + (kb * 32 + k) * 128 + j2) = make_float2(pt[k * 33 + j2], pt[k * 33 + j2 + 1]);
      }
      fence_async_global();   // the next step's TMA reads (async proxy) of these generic-proxy writes
    }
    if (stamp) g_train_stamps[cta][5] = clock64();
    grid_barrier(bar, epoch, s + 1 < a.n_steps);
    if (stamp) g_train_stamps[cta][6] = clock64();
  }
}

template <int NET>
__global__ void k_train_init(float* S) {
  using L = Lay<NET>;   // streaming copies of layers 2 and 3 from the parameter block
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W * W) {
    const int j = i / W, k = i % W;   // W_l[j][k]
    const float w1 = S[L::sP + L::oW1 + i], w2 = S[L::sP + L::oW2 + i];
    const size_t f = (size_t)(j >> 7) * HALF + k * 128 + (j & 127), b = (size_t)(k >> 7) * HALF + j * 128 + (k & 127);
    S[L::sF1 + f] = w1; S[L::sF2 + f] = w2;
    S[L::sB1 + b] = w1; S[L::sB2 + b] = w2;
  }
}

}  // namespace trn
}  // namespace idff

using namespace idff;
using namespace idff::trn;

template <int NET>
static int offsets_impl(int64_t* h_out12) {
  using L = Lay<NET>;
  const int64_t v[12] = {L::oW0, L::ob0, L::oW1, L::ob1, L::oW2, L::ob2, L::oW3, L::ob3, L::NP, (int64_t)L::sP, (int64_t)L::sM, (int64_t)L::sV};
  for (int i = 0; i < 12; ++i) h_out12[i] = v[i];
  return IDFF_OK;
}

extern "C" int64_t idff_train_state_floats(int32_t net) {
  return net == IDFF_TRAIN_NET_FLOW ? (int64_t)Lay<0>::sTOTAL : net == IDFF_TRAIN_NET_SCOREHEAD ? (int64_t)Lay<1>::sTOTAL : (int64_t)-1;
}
extern "C" int64_t idff_train_flow_state_floats(void) { return idff_train_state_floats(IDFF_TRAIN_NET_FLOW); }

extern "C" int idff_train_offsets(int32_t net, int64_t* h_out12) {
  IDFF_CHECK_ARG(h_out12, "idff_train_offsets: null pointer");
  IDFF_CHECK_ARG(net == IDFF_TRAIN_NET_FLOW || net == IDFF_TRAIN_NET_SCOREHEAD, "idff_train_offsets: unknown network %d", net);
  return net == IDFF_TRAIN_NET_FLOW ? offsets_impl<0>(h_out12) : offsets_impl<1>(h_out12);
}
extern "C" int idff_train_flow_offsets(int64_t* h_out12) { return idff_train_offsets(IDFF_TRAIN_NET_FLOW, h_out12); }

static thread_local int g_pair_mode = 1;   // 1 (default) = run a row group on a CTA pair where possible; 0 = always one CTA per row group (idff_debug_train_pair_mode)
extern "C" int idff_debug_train_pair_mode(int32_t on) {
  g_pair_mode = on ? 1 : 0;
  return IDFF_OK;
}

extern "C" int idff_debug_train_stamps(int64_t* h_out, int32_t n_ctas) {
  IDFF_CHECK_ARG(h_out && n_ctas >= 1 && n_ctas <= 256, "idff_debug_train_stamps: bad argument");
  IDFF_CUDA(cudaMemcpyFromSymbol(h_out, g_train_stamps, sizeof(long long) * 8 * n_ctas));
  return IDFF_OK;
}

extern "C" int idff_train_init(int32_t net, float* d_state, void* stream) {
  IDFF_CHECK_ARG(net == IDFF_TRAIN_NET_FLOW || net == IDFF_TRAIN_NET_SCOREHEAD, "idff_train_init: unknown network %d", net);
  IDFF_CHECK_ARG(d_state && (reinterpret_cast<uintptr_t>(d_state) & 15) == 0, "idff_train_init: state must be a 16-byte aligned device pointer");
  PtrDeviceGuard guard(d_state);
  if (net == IDFF_TRAIN_NET_FLOW) k_train_init<0><<<(W * W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_state);
  else k_train_init<1><<<(W * W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_state);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}
extern "C" int idff_train_flow_init(float* d_state, void* stream) { return idff_train_init(IDFF_TRAIN_NET_FLOW, d_state, stream); }

template <int NET>
static int steps_impl(float* d_state, const float* d_x0, const float* d_x1, const float* d_t, const float* d_eps, int32_t n_steps,
                      int32_t B, float sigma, int32_t sigma_mode, float score_sigma, const idff_adamw_t* opt, int64_t steps_done,
                      float* d_loss, float* d_gnorm, float* d_grad_dump, void* stream) {
  using L = Lay<NET>;
  PtrDeviceGuard guard(d_state);
  int dev = guard.dev, sms = 0, coop = 0;
  IDFF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDFF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop || sms < N_JOBS) {
    set_error("idff_train_steps: needs cooperative launch and >= %d SMs (device has %d)", N_JOBS, sms);
    return IDFF_EUNSUPPORTED;
  }
  int grid = sms < MAX_GRID ? sms : MAX_GRID;
  static unsigned long long attr_set = 0;   // bit per device: the attribute is per device, not per process (one static per NET)
  if (dev >= 64 || !((attr_set >> dev) & 1ull)) {
    IDFF_CUDA(cudaFuncSetAttribute(k_train_flow<NET, 1, RMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    IDFF_CUDA(cudaFuncSetAttribute(k_train_flow<NET, 2, RMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    if (dev < 64) attr_set |= 1ull << dev;
  }
  Args a;
  a.S = d_state;
  a.x0 = d_x0; a.x1 = d_x1; a.t = d_t; a.eps = d_eps;
  a.loss_out = d_loss; a.gnorm_out = d_gnorm; a.grad_dump = d_grad_dump;
  a.B = B; a.n_steps = n_steps; a.steps_done = steps_done;
  a.sigma = sigma; a.sigma_mode = sigma_mode; a.score_sigma = score_sigma;
  a.lr = opt->lr; a.beta1 = opt->beta1; a.beta2 = opt->beta2;
  a.decay = (float)(1.0 - opt->lr * opt->weight_decay);
  a.omb1 = (float)(1.0 - opt->beta1); a.omb2 = (float)(1.0 - opt->beta2);
  a.beta2f = (float)opt->beta2; a.eps_adam = (float)opt->eps; a.clip = (float)opt->max_grad_norm;
  a.b1pow = std::pow(opt->beta1, (double)steps_done); a.b2pow = std::pow(opt->beta2, (double)steps_done);
  cudaStream_t st = (cudaStream_t)stream;
  IDFF_CUDA(cudaMemsetAsync(d_state + L::sBAR, 0, 64 * sizeof(float), st));
  // Default where the batch allows it (2 * B / 4 CTAs fit the grid: B <= 296 on 148 SMs): a CTA PAIR (cluster of 2) per 4 rows, each
  // CTA streaming and computing half of the units, activation halves exchanged through distributed shared memory.  With the
  // hand-over as st.async stores that signal the peer's mbarrier this measures 22.0 against 25.8 us/step at B = 256 (round 1,
  // with st.shared::cluster + a cluster-scope arrive / wait per layer, the pair was SLOWER: 27 vs 25 us).  Larger batches, and
  // idff_debug_train_pair_mode(0): one CTA per 4 rows.
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cfg.attrs = attr;
  int pairs = 0;
  const int grid2 = grid & ~1;
  if (g_pair_mode && 2 * (B / RMIN) <= grid2 && grid2 >= N_JOBS) {
    cfg.gridDim = dim3(grid2);
    cfg.numAttrs = 2;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, k_train_flow<NET, 2, RMIN>, &cfg) == cudaSuccess && 2 * ncl >= grid2) pairs = 1;
    else (void)cudaGetLastError();
  }
  if (pairs) {
    IDFF_CUDA(cudaLaunchKernelEx(&cfg, k_train_flow<NET, 2, RMIN>, a));
  } else {
    cfg.gridDim = dim3(grid);
    cfg.numAttrs = 1;
    IDFF_CUDA(cudaLaunchKernelEx(&cfg, k_train_flow<NET, 1, RMIN>, a));
  }
  count_launch();
  return IDFF_OK;
}

extern "C" int idff_train_steps(int32_t net, float* d_state, const float* d_x0, const float* d_x1, const float* d_t,
                                const float* d_eps, int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode, float score_sigma,
                                const idff_adamw_t* opt, int64_t steps_done, float* d_loss, float* d_gnorm, float* d_grad_dump,
                                void* stream) {
  IDFF_CHECK_ARG(net == IDFF_TRAIN_NET_FLOW || net == IDFF_TRAIN_NET_SCOREHEAD, "idff_train_steps: unknown network %d", net);
  IDFF_CHECK_ARG(d_state && (reinterpret_cast<uintptr_t>(d_state) & 15) == 0, "idff_train_steps: state must be a 16-byte aligned device pointer");
  IDFF_CHECK_ARG(n_steps >= 0 && n_steps <= (1 << 20), "idff_train_steps: n_steps %d outside [0, 2^20]", n_steps);
  if (n_steps == 0) return IDFF_OK;
  IDFF_CHECK_ARG(d_x0 && d_x1 && d_t && opt, "idff_train_steps: null pointer");
  IDFF_CHECK_ARG(((reinterpret_cast<uintptr_t>(d_x0) | reinterpret_cast<uintptr_t>(d_x1) | reinterpret_cast<uintptr_t>(d_eps)) & 7) == 0,
                 "idff_train_steps: x0, x1, eps must be 8-byte aligned");
  IDFF_CHECK_ARG(B >= RMIN && B <= MAXB && B % RMIN == 0, "idff_train_steps: batch %d must be a multiple of %d in [%d, %d]", B, RMIN, RMIN, MAXB);
  IDFF_CHECK_ARG(sigma_mode == 0 || sigma_mode == 1, "idff_train_steps: sigma_mode %d", sigma_mode);
  IDFF_CHECK_ARG(steps_done >= 0, "idff_train_steps: steps_done < 0");
  if (net == IDFF_TRAIN_NET_FLOW)
    return steps_impl<0>(d_state, d_x0, d_x1, d_t, d_eps, n_steps, B, sigma, sigma_mode, 0.0f, opt, steps_done, d_loss, d_gnorm, d_grad_dump, stream);
  return steps_impl<1>(d_state, d_x0, d_x1, d_t, d_eps, n_steps, B, sigma, sigma_mode, score_sigma, opt, steps_done, d_loss, d_gnorm, d_grad_dump, stream);
}

extern "C" int idff_train_flow_steps(float* d_state, const float* d_x0, const float* d_x1, const float* d_t,
                                     const float* d_eps, int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode,
                                     const idff_adamw_t* opt, int64_t steps_done, float* d_loss, float* d_gnorm,
                                     float* d_grad_dump, void* stream) {
  return idff_train_steps(IDFF_TRAIN_NET_FLOW, d_state, d_x0, d_x1, d_t, d_eps, n_steps, B, sigma, sigma_mode, 0.0f, opt, steps_done,
                          d_loss, d_gnorm, d_grad_dump, stream);
}
