// The PolyALA training step (TrajectoryGenerator.train_model, timeseris_example/MD_simulation.py:117-144) as ONE persistent kernel
// on a thread-block CLUSTER of 8 CTAs whose shared memories hold the whole training state.
//
// Reference step (batch_size = 10, MD_simulation.py:544):
//   index_sel = randint(1, T, (B,)); x0 = dataset[index_sel - 1]; x1 = dataset[index_sel]                      :128-130
//   t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1)      ExactOT matcher, sigma 0.5            :121,132
//   vt = model(cat[xt, t], index_sel)         MLP_Embedding: cat[xt, t, Embedding(T,128)[idx]] (175) -> 128 -> 128 -> 128 -> 46   :135
//   loss = mean((vt - x1)^2) + mean((vt - (x1 + 360))^2) + mean((vt - (x1 - 360))^2)                            :137-141
//   loss.backward(); Adam().step()            torch.optim.Adam defaults, no clipping                            :119,143-144
//
// At batch 10 the step is 87 086 parameters and 3.6 MFLOP: pure latency (~70 torch kernels; 300 us even as a CUDA graph).  Here the
// parameters AND both Adam moments (3 x 348 KB) live in the shared memory of 8 CTAs for the whole launch, `n_steps` steps run back to
// back, and nothing but the step's inputs (a few KB) is read from global memory:
//   * CTA c owns hidden units [16c, 16c+16) of every hidden layer (rows of W0, W1, W2 with their biases), outputs [6c, 6c+6) (rows of
//     W3) and columns [16c, 16c+16) of the embedding table -- parameter, exp_avg and exp_avg_sq side by side.
//   * forward:  a layer's input activations [B][128] are complete in every CTA; a CTA computes its 16 units and PUSHES them into the
//     8 activation buffers of the cluster (16-byte st.async stores that signal the destination's mbarrier), one hand-over per layer.
//   * backward: dA_{l-1}[r][k] = sum_j dZ_l[r][j] W_l[j][k] is a sum over all units j: a CTA forms the partial sums of its 16 units for
//     all k and scatters them to the owners of k (reduce-scatter through distributed shared memory, fixed summation order ->
//     bit-reproducible); the owner applies SELU' (derivative codes kept from the forward pass).
//   * weight gradients never leave their owner: dW_l[j][k] = sum_r dZ_l[r][j] a_{l-1}[r][k] needs only the owner's dZ rows and the
//     (complete) activations; Adam (torch's operation order, bias corrections in double) is applied in the same thread, right after the
//     layer's backward-data partials have been formed from the old weights.  No gradient norm, so no other cross-CTA reduction.
//   * the embedding rows of the batch are gathered by the column owners and pushed like activations; their gradient comes back through
//     the same reduce-scatter (duplicated indices accumulate in row order); the dense Adam update touches all T rows, as torch does.
// 8 exchanges per step, each a set of st.async stores that signal the destination's mbarrier (no cluster barrier in the loop).  Parameters are read from / written back to the caller's state buffer (nn.Module layout, so the torch
// parameters can alias it) at the start / end of a launch.
#include <cmath>

#include "idff_common.cuh"

namespace idff {
namespace tmd {

constexpr int W = 128, CL = 8, UPC = 16, NT = 512, BMAX = IDFF_TRAIN_MD_MAX_BATCH, DMAX = 47, DP = 48, TMAX = IDFF_TRAIN_MD_MAX_FRAMES;
constexpr int OPC = 6;                // outputs per CTA (ceil(47 / 8))
constexpr int K0P = DP + W;           // layer-0 input as laid out on chip: [x (D) | t | zero pad to 48 | embedding (128)]
constexpr int W0S = K0P + 4, WS = W + 4;   // row strides of the weight slices (bank spread for 16 rows read at once)
constexpr int C0S = K0P + 4;          // row stride of the layer-0 input rows
constexpr int AS = W + 4;             // row stride of the activation buffers
constexpr int NX = 8;                 // exchange points per step, one mbarrier each

// shared-memory plan (floats).  Parameter slices: three copies (param, exp_avg, exp_avg_sq) `PT` apart.
constexpr int pW0 = 0, pb0 = pW0 + UPC * W0S, pW1 = pb0 + UPC, pb1 = pW1 + UPC * WS, pW2 = pb1 + UPC, pb2 = pW2 + UPC * WS,
              pW3 = pb2 + UPC, pb3 = pW3 + OPC * WS, pE = pb3 + 8, PT = pE + TMAX * UPC;
constexpr int mCAT = 3 * PT, mA1 = mCAT + 2 * BMAX * C0S, mA2 = mA1 + BMAX * AS, mA3 = mA2 + BMAX * AS, mX1 = mA3 + BMAX * AS,
              mDVT = mX1 + BMAX * DP, mCODE = mDVT + BMAX * 8, mDZ = mCODE + 3 * BMAX * UPC, mDE = mDZ + BMAX * UPC,
              mSTG = mDE + BMAX * UPC, mRECV = mSTG + BMAX * UPC, mCOMB = mRECV + 2 * CL * BMAX * UPC, mLT = mCOMB + 8 * BMAX * UPC,
              mLRECV = mLT + BMAX * 8 * 3, mMISC = mLRECV + CL * 3 * 2, mROW = mMISC + 64, mBAR = mROW + 3 * TMAX, mTOTAL = mBAR + 2 * NX;
constexpr size_t kSmemBytes = (size_t)mTOTAL * sizeof(float);
static_assert(kSmemBytes <= 232448, "shared memory budget");
static_assert(mLRECV % 2 == 0 && mMISC % 2 == 0 && mBAR % 2 == 0, "8-byte aligned slots");

// clock64 stamps of the last step of the last launch (CTA 0, thread 0): diagnostic (idff_debug_train_md_stamps)
__device__ long long g_md_stamps[32];

struct Args {
  float* S;                 // state: [param block | exp_avg | exp_avg_sq], each NPP floats
  const float* dataset;     // [T][D]
  const long long* idx;     // [n_steps][B]
  const long long *row0, *row1;   // [n_steps][B] dataset rows of x0 / x1 when the batch was re-paired (NULL: idx - 1 / idx)
  const float *t, *eps;     // [n_steps][B], [n_steps][B][D] (eps may be null)
  float *loss_out, *grad_dump;
  int B, D, T, n_steps;
  long long npp;            // floats per block of the state
  float sigma;
  int sigma_mode;
  double lr, beta1, beta2, b1pow, b2pow;
  float omb1, omb2, beta2f, eps_adam;
};

// parameter-block offsets (module.parameters() order of MLP_Embedding: time_index_embedding.weight, net.0.weight, net.0.bias, ...)
struct Off {
  int K0, oE, oW0, ob0, oW1, ob1, oW2, ob2, oW3, ob3, NP;
  __host__ __device__ explicit Off(int D, int T) {
    K0 = D + 1 + W;
    oE = 0; oW0 = oE + T * W; ob0 = oW0 + W * K0; oW1 = ob0 + W; ob1 = oW1 + W * W; oW2 = ob1 + W; ob2 = oW2 + W * W;
    oW3 = ob2 + W; ob3 = oW3 + D * W; NP = ob3 + D;
  }
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// remote store that carries its own completion signal: the destination CTA's mbarrier `bar` receives complete_tx(bytes) when the
// data has landed (SASS STAS) -- the consumer waits for its expected byte count, no fence / arrive / cluster barrier on either side
__device__ __forceinline__ void st_async4(uint32_t addr, float4 v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_async_f64(uint32_t addr, double v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr),
               "l"(__double_as_longlong(v)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ float selu_d1(float code) { return code < 0.0f ? kSeluLambda : code; }

// torch.optim.Adam (single tensor, weight_decay 0): exp_avg.lerp_(g, 1 - b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2);
// denom = sqrt(v) / sqrt(bc2) + eps; p -= (lr / bc1) * m / denom.  The step is instruction-bound at 16 warps per SM, and IEEE
// sqrt / division cost ~35 instructions per element: sqrt and the quotient use the approximate units (sqrt.approx, rcp-based
// division: <= 2 ulp each), 1 / sqrt(bc2) is formed once per step in double.  The update differs from torch's by ~1e-7 relative.
struct AdamK {
  float omb1, omb2, beta2f, eps, step_size, inv_bc2_sqrt;
};
__device__ __forceinline__ void adam1(const AdamK& k, float g, float& p, float& m, float& v) {
  m = fmaf(g - m, k.omb1, m);
  v = fmaf(k.omb2 * g, g, v * k.beta2f);
  float sq;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
  const float denom = fmaf(sq, k.inv_bc2_sqrt, k.eps);
  p = fmaf(-k.step_size, __fdividef(m, denom), p);
}

// ---- one hidden layer forward for this CTA's 16 units.  Thread (kh, rp, up): rows {rp, rp + 8} x units {up, up + 8} over K-slice kh
// of 8 -- per float4 step two weight loads (8 distinct rows, conflict free at stride 132 / 180) and two activation loads serve 16
// FMAs.  The slices meet in `comb` [8][16][16]; a = selu(z) goes to `stg` [r][u], the derivative code to `code` [r][u].
// in: [BMAX][is] (complete), Wp: [16][ws] (K4 float4 per row used), bias [16].
template <int K4, int IS, int WSTR>
__device__ __forceinline__ void fwd_units(const float* in, const float* Wp, const float* bias, float* comb, float* stg, float* code,
                                          int tid) {
  constexpr int is = IS, ws = WSTR;
  const int kh = tid >> 6, rp = (tid >> 3) & 7, up = tid & 7;
  constexpr int per = (K4 + 7) >> 3;
  const int k0 = kh * per, k1 = min(K4, k0 + per);
  const float4* w0 = reinterpret_cast<const float4*>(Wp + up * ws);
  const float4* w1 = reinterpret_cast<const float4*>(Wp + (up + 8) * ws);
  const float4* x0 = reinterpret_cast<const float4*>(in + rp * is);
  const float4* x1 = reinterpret_cast<const float4*>(in + (rp + 8) * is);
  float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
#pragma unroll
  for (int kk = 0; kk < per; ++kk) {
    const int k = k0 + kk;
    if (k >= k1) break;
    const float4 wa = w0[k], wb = w1[k], xa = x0[k], xb = x1[k];
    c00 = fmaf(xa.x, wa.x, c00); c00 = fmaf(xa.y, wa.y, c00); c00 = fmaf(xa.z, wa.z, c00); c00 = fmaf(xa.w, wa.w, c00);
    c01 = fmaf(xa.x, wb.x, c01); c01 = fmaf(xa.y, wb.y, c01); c01 = fmaf(xa.z, wb.z, c01); c01 = fmaf(xa.w, wb.w, c01);
    c10 = fmaf(xb.x, wa.x, c10); c10 = fmaf(xb.y, wa.y, c10); c10 = fmaf(xb.z, wa.z, c10); c10 = fmaf(xb.w, wa.w, c10);
    c11 = fmaf(xb.x, wb.x, c11); c11 = fmaf(xb.y, wb.y, c11); c11 = fmaf(xb.z, wb.z, c11); c11 = fmaf(xb.w, wb.w, c11);
  }
  float* cb = comb + kh * (BMAX * UPC);
  cb[rp * UPC + up] = c00; cb[rp * UPC + up + 8] = c01; cb[(rp + 8) * UPC + up] = c10; cb[(rp + 8) * UPC + up + 8] = c11;
  __syncthreads();
  if (tid < 256) {
    float z = comb[tid];
#pragma unroll
    for (int k = 1; k < 8; ++k) z += comb[k * (BMAX * UPC) + tid];
    z += bias[tid & 15];
    float c;
    stg[tid] = selu_fwd(z, c);
    code[tid] = c;
  }
  __syncthreads();
}

// push stg [B][16] into column block `rank` of the [BMAX][stride] buffer at shared offset `dst_off` (+ col0) of every CTA
__device__ __forceinline__ void push_units(const float* stg, int B, uint32_t dst_saddr, int stride, int col0, uint32_t rank, uint32_t bar,
                                           int tid) {
  if (tid < 64 * CL) {
    const int peer = tid >> 6, r = (tid >> 2) & 15, u4 = tid & 3;
    if (r < B) {
      const float4 v = *reinterpret_cast<const float4*>(stg + r * UPC + 4 * u4);
      st_async4(mapa(dst_saddr + (uint32_t)((r * stride + col0 + (int)rank * UPC + 4 * u4) * 4), (uint32_t)peer), v,
                mapa(bar, (uint32_t)peer));
    }
  }
}

// partial sums of the backward-data product over this CTA's J rows: out[r][k] = sum_{j<J} dz[r][j] * Wp[j][wc0 + k], k < 128,
// scattered to the owners of k: recv[rank][r][k % 16] of CTA k / 16.  ALIGNED: wc0 is a multiple of 4 floats.
template <bool ALIGNED>
__device__ __forceinline__ void partial_scatter(const float* dz, int dzs, int J, const float* Wp, int ws, int wc0, int B,
                                                uint32_t recv_saddr, uint32_t rank, uint32_t bar, int tid) {
  const int r = tid >> 5, kq = tid & 31;
  if (r < B) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < J; ++j) {
      const float d = dz[r * dzs + j];
      float4 w;
      if (ALIGNED) {
        w = *reinterpret_cast<const float4*>(Wp + j * ws + wc0 + 4 * kq);
      } else {
        const float* wp = Wp + j * ws + wc0 + 4 * kq;
        w = make_float4(wp[0], wp[1], wp[2], wp[3]);
      }
      acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
    }
    st_async4(mapa(recv_saddr + (uint32_t)((((int)rank * BMAX + r) * UPC + 4 * (kq & 3)) * 4), (uint32_t)(kq >> 2)), acc,
              mapa(bar, (uint32_t)(kq >> 2)));
  }
}

// dst[r][u] = (sum over the 8 sources of recv[src][r][u]) [* selu'(code[r][u])], r < B; rows >= B are left alone (zero)
__device__ __forceinline__ void gather_recv(const float* recv, const float* code, float* dst, int B, int tid) {
  if (tid < 256) {
    const int r = tid >> 4, u = tid & 15;
    if (r < B) {
      float s = recv[r * UPC + u];
#pragma unroll
      for (int c = 1; c < CL; ++c) s += recv[(c * BMAX + r) * UPC + u];
      dst[r * UPC + u] = code ? s * selu_d1(code[r * UPC + u]) : s;
    }
  }
}

// weight gradient of J rows x K4 float4 columns, dW[j][k] = sum_{r<B} dz[r][j] * act[r][k], and Adam on the owner's copies.
// gidx(j, k) -> index in the parameter block (for the gradient dump) or -1 for padding columns.
template <int K4, int AST, int WSTR, int DZS, class GI>
__device__ __forceinline__ void wgrad_adam(const float* dz, int J, const float* act, float* P, int B, const AdamK& ak, float* dump,
                                           GI gidx, int tid) {
  constexpr int as = AST, ws = WSTR, dzs = DZS;
  for (int it = tid; it < J * K4; it += NT) {
    const int j = it / K4, k4 = it - j * K4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < B; ++r) {
      const float d = dz[r * dzs + j];
      const float4 a = *reinterpret_cast<const float4*>(act + r * as + 4 * k4);
      g.x = fmaf(d, a.x, g.x); g.y = fmaf(d, a.y, g.y); g.z = fmaf(d, a.z, g.z); g.w = fmaf(d, a.w, g.w);
    }
    float4* p4 = reinterpret_cast<float4*>(P + j * ws + 4 * k4);
    float4* m4 = reinterpret_cast<float4*>(P + PT + j * ws + 4 * k4);
    float4* v4 = reinterpret_cast<float4*>(P + 2 * PT + j * ws + 4 * k4);
    float4 p = *p4, m = *m4, v = *v4;
    adam1(ak, g.x, p.x, m.x, v.x); adam1(ak, g.y, p.y, m.y, v.y); adam1(ak, g.z, p.z, m.z, v.z); adam1(ak, g.w, p.w, m.w, v.w);
    *p4 = p; *m4 = m; *v4 = v;
    if (dump) {
      const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ix = gidx(j, 4 * k4 + c);
        if (ix >= 0) dump[ix] = gg[c];
      }
    }
  }
}
__device__ __forceinline__ void bgrad_adam(const float* dz, int dzs, int J, float* P, int B, const AdamK& ak, float* dump, int gbase,
                                           int tid) {
  if (tid < J) {
    float g = 0.f;
    for (int r = 0; r < B; ++r) g += dz[r * dzs + tid];
    adam1(ak, g, P[tid], P[PT + tid], P[2 * PT + tid]);
    if (dump) dump[gbase + tid] = g;
  }
}

// chip column c of the layer-0 input <-> column of net.0.weight: [x (D) | t] first, the embedding after the pad
__device__ __forceinline__ int w0_col(int c, int D) { return c <= D ? c : (c >= DP ? c - (DP - (D + 1)) : -1); }

// copy this CTA's slices between the state buffer and shared memory (dir 0: load P, M, V; dir 1: store)
__device__ void move_params(const Args& a, const Off& o, float* sm, uint32_t rank, int n_own, int dir, int tid) {
  const int D = a.D, T = a.T, K0 = o.K0;
  for (int b = 0; b < 3; ++b) {
    float* G = a.S + (size_t)b * a.npp;
    float* L = sm + b * PT;
    auto mv = [&](int li, int gi) {
      if (dir == 0) L[li] = gi >= 0 ? G[gi] : 0.0f;
      else if (gi >= 0) G[gi] = L[li];
    };
    for (int i = tid; i < UPC * K0P; i += NT) {
      const int j = i / K0P, c = i - j * K0P, gc = w0_col(c, D);
      mv(pW0 + j * W0S + c, gc >= 0 ? o.oW0 + ((int)rank * UPC + j) * K0 + gc : -1);
    }
    for (int i = tid; i < UPC * W; i += NT) {
      const int j = i >> 7, c = i & 127;
      mv(pW1 + j * WS + c, o.oW1 + ((int)rank * UPC + j) * W + c);
      mv(pW2 + j * WS + c, o.oW2 + ((int)rank * UPC + j) * W + c);
    }
    for (int i = tid; i < OPC * W; i += NT) {
      const int j = i >> 7, c = i & 127;
      mv(pW3 + j * WS + c, j < n_own ? o.oW3 + ((int)rank * OPC + j) * W + c : -1);
    }
    if (tid < UPC) {
      mv(pb0 + tid, o.ob0 + (int)rank * UPC + tid);
      mv(pb1 + tid, o.ob1 + (int)rank * UPC + tid);
      mv(pb2 + tid, o.ob2 + (int)rank * UPC + tid);
    }
    if (tid < 8) mv(pb3 + tid, tid < n_own ? o.ob3 + (int)rank * OPC + tid : -1);
    for (int i = tid; i < T * UPC; i += NT) {
      const int row = i >> 4, e = i & 15;
      mv(pE + row * UPC + e, o.oE + row * W + (int)rank * UPC + e);
    }
  }
}

// phase stamp of the last step (CTA 0, thread 0): one predicated store per call site
#define MD_STAMP() do { if (stamp) g_md_stamps[sti] = clock64(); ++sti; } while (0)

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) k_train_md(const Args a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, B = a.B, D = a.D, T = a.T;
  const uint32_t rank = cluster_ctarank();
  const Off o(D, T);
  const int n_own = max(0, min(OPC, D - (int)rank * OPC));   // outputs [6 rank, 6 rank + n_own)
  // layer-0 input rows [BMAX][C0S]: x (D), t, pad, embedding row (128 at column DP) -- one buffer per step parity: a peer that is
  // already in the next step pushes its embedding columns while this CTA may still read the current rows for dW0
  float* a1 = sm + mA1; float* a2 = sm + mA2; float* a3 = sm + mA3;
  float* x1s = sm + mX1;       // [BMAX][DP]
  float* dvt = sm + mDVT;      // [BMAX][8]
  float* code = sm + mCODE;    // [3][BMAX][16]
  float* dz = sm + mDZ;        // [BMAX][16]
  float* dE = sm + mDE;        // [BMAX][16]: gradient of this CTA's embedding columns, per batch row
  float* stg = sm + mSTG;      // [BMAX][16]
  float* recv = sm + mRECV;    // [2][CL][BMAX][16]
  float* comb = sm + mCOMB;
  float* lt = sm + mLT;        // [BMAX][8][3]
  double* lrecv = reinterpret_cast<double*>(sm + mLRECV);   // [CL][3] (CTA 0)
  int* idxb = reinterpret_cast<int*>(sm + mMISC);           // [2][BMAX]: frame indices of the current / the other step (by parity)
  float* sc = sm + mMISC + 32;                              // [3][2]: step_size, 1 / bc2_sqrt of step s at slot s % 3
  int* rowfirst = reinterpret_cast<int*>(sm + mROW);        // [2][TMAX]: first batch row that uses frame `row` (-1: none), by parity
  int* need = rowfirst + 2 * TMAX;                          // [TMAX]: the next step gathers this embedding row

  // one mbarrier per exchange point of a step: every remote store (st.async) signals its byte count on the destination's barrier,
  // the consumer waits for the bytes of all 8 sources.  A step uses one phase of each barrier (parity = step parity); thread 0
  // re-arms a barrier for the next step right after its own wait.
  const uint32_t xbar = s32(sm + mBAR);
  const uint32_t xbytes = (uint32_t)(CL * B * UPC * 4);
  auto xb = [&](int k) { return xbar + 8u * (uint32_t)k; };
  auto xexpect = [&](int k) { return xbytes + ((k == 4 && rank == 0) ? (uint32_t)(CL * 3 * 8) : 0u); };   // + the loss sums (CTA 0)

  for (int i = tid; i < mTOTAL - 3 * PT; i += NT) sm[3 * PT + i] = 0.0f;
  __syncthreads();
  for (int i = tid; i < 2 * TMAX; i += NT) rowfirst[i] = -1;
  if (tid == 0) {
    for (int k = 0; k < NX; ++k) mbar_init(xb(k), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  move_params(a, o, sm, rank, n_own, 0, tid);
  __syncthreads();
  cluster_sync_all();   // every CTA's buffers and barriers are initialised before the first push arrives
  if (tid == 0 && a.n_steps > 0)
    for (int k = 0; k < NX; ++k) mbar_arrive_expect_tx(xb(k), xexpect(k));

  const int n_el = B * D;   // (row, dim) elements of a batch: up to 2 per thread
  float px0[2], px1[2], pep[2], ptt[2];
  int pidx[2], er[2], ed[2];   // this thread's elements: row, dimension (fixed for the launch)
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int e = tid + q * NT;
    er[q] = e / D;
    ed[q] = e - er[q] * D;
  }
  auto prefetch = [&](int s) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + q * NT;
      if (e < n_el) {
        const int r = er[q], d = ed[q];
        long long ix = a.idx[(size_t)s * B + r];
        ix = ix < 1 ? 1 : (ix > T - 1 ? T - 1 : ix);   // frames [1, T): randint(1, T) in the reference (torch would raise outside)
        pidx[q] = (int)ix;
        long long r0 = a.row0 ? a.row0[(size_t)s * B + r] : ix - 1, r1 = a.row1 ? a.row1[(size_t)s * B + r] : ix;
        r0 = r0 < 0 ? 0 : (r0 > T - 1 ? T - 1 : r0);
        r1 = r1 < 0 ? 0 : (r1 > T - 1 ? T - 1 : r1);
        px0[q] = a.dataset[(size_t)r0 * D + d];
        px1[q] = a.dataset[(size_t)r1 * D + d];
        pep[q] = a.eps ? a.eps[((size_t)s * B + r) * D + d] : 0.0f;
        ptt[q] = a.t[(size_t)s * B + r];
      }
    }
  };
  // frame indices of the prefetched step -> idxb[par]; rowfirst[par] (first batch row per frame)
  auto publish_idx = [&](int par) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + q * NT;
      if (e < n_el && ed[q] == 0) idxb[par * BMAX + er[q]] = pidx[q];
    }
  };
  double b1pow = a.b1pow, b2pow = a.b2pow;
  auto adam_of = [&](int slot) {
    AdamK k;
    k.omb1 = a.omb1; k.omb2 = a.omb2; k.beta2f = a.beta2f; k.eps = a.eps_adam; k.step_size = sc[slot * 2]; k.inv_bc2_sqrt = sc[slot * 2 + 1];
    return k;
  };
  // bias corrections in double like torch's Python scalars; the last thread forms them one step ahead, off the critical path
  auto step_scalars = [&](int slot) {
    b1pow *= a.beta1;
    b2pow *= a.beta2;
    sc[slot * 2] = (float)(a.lr / (1.0 - b1pow));
    sc[slot * 2 + 1] = (float)(1.0 / sqrt(1.0 - b2pow));
  };
  // Adam on one element of this CTA's embedding columns with the gradient of step parity `par` (rows of the batch that use the frame)
  auto emb_item = [&](int row, int e, int par, const AdamK& k, float* dump) {
    const int* ix = idxb + par * BMAX;
    float g = 0.f;
    const int rs = rowfirst[par * TMAX + row];
    if (rs >= 0)
      for (int r = rs; r < B; ++r)
        if (ix[r] == row) g += dE[r * UPC + e];
    float* p = sm + pE + row * UPC + e;
    adam1(k, g, p[0], p[PT], p[2 * PT]);
    if (dump) dump[o.oE + row * W + (int)rank * UPC + e] = g;
  };
  const uint32_t a1_s = s32(a1), a2_s = s32(a2), a3_s = s32(a3), recv_s = s32(recv);

  if (a.n_steps > 0) {
    prefetch(0);
    publish_idx(0);
    if (tid == NT - 1) step_scalars(0);
  }
  __syncthreads();

  for (int s = 0; s < a.n_steps; ++s) {
    const int par = s & 1;
    const bool more = s + 1 < a.n_steps;
    const int* idxs = idxb + par * BMAX;
    float* cat0 = sm + mCAT + par * (BMAX * C0S);
    const uint32_t cat_s = s32(cat0);
    // wait for exchange k of this step; thread 0 arms the barrier for the next step
    auto xwait = [&](int k) {
      mbar_wait(xb(k), (uint32_t)par);
      if (tid == 0 && more) mbar_arrive_expect_tx(xb(k), xexpect(k));
    };
    float* dump = (a.grad_dump && s == a.n_steps - 1) ? a.grad_dump : nullptr;
    const bool stamp = rank == 0 && tid == 0 && s == a.n_steps - 1;
    int sti = 0;
    MD_STAMP();
    // ---- phase 0: path sample (conditional_flow_matching.py:98-123, no contraction), embedding gather + push
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + q * NT;
      if (e < n_el) {
        const int r = er[q], d = ed[q];
        const float t = ptt[q];
        const float st = a.sigma_mode == 0 ? a.sigma : __fmul_rn(a.sigma, __fsqrt_rn(__fmul_rn(t, __fsub_rn(1.0f, t))));
        const float xt = __fadd_rn(__fadd_rn(__fmul_rn(t, px1[q]), __fmul_rn(__fsub_rn(1.0f, t), px0[q])), __fmul_rn(st, pep[q]));
        cat0[r * C0S + d] = xt;
        x1s[r * DP + d] = px1[q];
        if (d == 0) cat0[r * C0S + D] = t;
      }
    }
    if (tid < B) {   // first batch row of every frame in use (duplicates add up in row order)
      bool first = true;
      for (int r = 0; r < tid; ++r) first = first && idxs[r] != idxs[tid];
      if (first) rowfirst[par * TMAX + idxs[tid]] = tid;
    }
    if (tid >= 64 && tid < 128) {   // this CTA's 16 embedding columns of the batch rows -> stg
      const int r = (tid - 64) >> 2, u4 = tid & 3;
      if (r < B) *reinterpret_cast<float4*>(stg + r * UPC + 4 * u4) = *reinterpret_cast<const float4*>(sm + pE + idxs[r] * UPC + 4 * u4);
    }
    __syncthreads();
    if (s + 1 < a.n_steps) prefetch(s + 1);   // in flight for the whole step
    const AdamK ak = adam_of(s % 3);
    push_units(stg, B, cat_s, C0S, DP, rank, xb(0), tid);                                 // exchange 0
    MD_STAMP();
    if (tid == NT - 1 && more) step_scalars((s + 1) % 3);
    // The previous step's dense Adam update of the embedding rows this step does NOT gather (the gathered ones were updated at the
    // end of that step) runs in four slices behind the four forward hand-overs, while the pushes of the cluster are in flight.
    const AdamK akp = adam_of((s + 2) % 3);
    auto emb_deferred = [&](int q) {
      if (s > 0)
        for (int it = tid + q * NT; it < T * UPC; it += 4 * NT)
          if (!need[it >> 4]) emb_item(it >> 4, it & 15, par ^ 1, akp, nullptr);
    };
    emb_deferred(0);
    xwait(0);
    MD_STAMP();
    // ---- forward
    fwd_units<K0P / 4, C0S, W0S>(cat0, sm + pW0, sm + pb0, comb, stg, code, tid);
    push_units(stg, B, a1_s, AS, 0, rank, xb(1), tid);                                    // exchange 1
    MD_STAMP();
    emb_deferred(1);
    xwait(1);
    MD_STAMP();
    fwd_units<W / 4, AS, WS>(a1, sm + pW1, sm + pb1, comb, stg, code + BMAX * UPC, tid);
    push_units(stg, B, a2_s, AS, 0, rank, xb(2), tid);                                    // exchange 2
    MD_STAMP();
    emb_deferred(2);
    xwait(2);
    MD_STAMP();
    fwd_units<W / 4, AS, WS>(a2, sm + pW2, sm + pb2, comb, stg, code + 2 * BMAX * UPC, tid);
    push_units(stg, B, a3_s, AS, 0, rank, xb(3), tid);                                    // exchange 3
    MD_STAMP();
    emb_deferred(3);
    __syncthreads();
    if (s > 0 && tid < B) {   // the deferred update is complete: retire its bookkeeping
      need[idxs[tid]] = 0;
      rowfirst[(par ^ 1) * TMAX + idxb[(par ^ 1) * BMAX + tid]] = -1;
    }
    xwait(3);
    MD_STAMP();
    // ---- output layer (this CTA's outputs), periodic loss (MD_simulation.py:137-141) and d loss / d vt
    {
      const int i8 = tid >> 2, ks = tid & 3, r = i8 >> 3, oo = i8 & 7;
      const bool valid = r < B && oo < n_own;
      float acc = 0.f;
      if (valid) {
        const float4* w4 = reinterpret_cast<const float4*>(sm + pW3 + oo * WS);
        const float4* a4 = reinterpret_cast<const float4*>(a3 + r * AS);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = w4[ks + 4 * i], av = a4[ks + 4 * i];
          c0 = fmaf(av.x, w.x, c0); c1 = fmaf(av.y, w.y, c1); c2 = fmaf(av.z, w.z, c2); c3 = fmaf(av.w, w.w, c3);
        }
        acc = (c0 + c1) + (c2 + c3);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (ks == 0) {
        float d = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f;
        if (valid) {
          const float vt = acc + sm[pb3 + oo];
          const float b1 = x1s[r * DP + (int)rank * OPC + oo];
          const float r0 = __fsub_rn(vt, b1), r1 = __fsub_rn(vt, __fadd_rn(b1, 360.0f)), r2 = __fsub_rn(vt, __fsub_rn(b1, 360.0f));
          l0 = __fmul_rn(r0, r0); l1 = __fmul_rn(r1, r1); l2 = __fmul_rn(r2, r2);
          d = 2.0f * (r0 + r1 + r2) * (1.0f / (float)(B * D));
        }
        dvt[r * 8 + oo] = d;
        lt[(r * 8 + oo) * 3 + 0] = l0; lt[(r * 8 + oo) * 3 + 1] = l1; lt[(r * 8 + oo) * 3 + 2] = l2;
      }
    }
    __syncthreads();
    if (warp < 3) {   // this CTA's share of the three loss sums (fp64, fixed order: lane-strided, then a butterfly) -> CTA 0
      double sum = 0.0;
      for (int i = lane; i < B * 8; i += 32) sum += (double)lt[i * 3 + warp];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      if (lane == 0) st_async_f64(mapa(s32(lrecv + (int)rank * 3 + warp), 0u), sum, mapa(xb(4), 0u));
    }
    partial_scatter<true>(dvt, 8, n_own, sm + pW3, WS, 0, B, recv_s, rank, xb(4), tid);               // exchange 4 -> recv[0]
    MD_STAMP();
    __syncthreads();   // W3 has been read by the partial products
    wgrad_adam<W / 4, AS, WS, 8>(dvt, n_own, a3, sm + pW3, B, ak, dump,
                                 [&](int j, int c) { return o.oW3 + ((int)rank * OPC + j) * W + c; }, tid);
    bgrad_adam(dvt, 8, n_own, sm + pb3, B, ak, dump, o.ob3 + (int)rank * OPC, tid);
    xwait(4);
    MD_STAMP();
    if (rank == 0 && tid == 0 && a.loss_out) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int c = 0; c < CL; ++c) { s0 += lrecv[c * 3]; s1 += lrecv[c * 3 + 1]; s2 += lrecv[c * 3 + 2]; }
      const double n = (double)(B * D);
      a.loss_out[s] = ((float)(s0 / n) + (float)(s1 / n)) + (float)(s2 / n);
    }
    // ---- backward through layers 3, 2 (dZ of the owner's units, partials of the next dA; the weight gradients + Adam run
    // after the CTA has pushed the layer's partial sums, i.e. while they travel and the peers catch up)
    gather_recv(recv, code + 2 * BMAX * UPC, dz, B, tid);
    __syncthreads();
    partial_scatter<true>(dz, UPC, UPC, sm + pW2, WS, 0, B, recv_s + CL * BMAX * UPC * 4, rank, xb(5), tid);   // exchange 5 -> recv[1]
    MD_STAMP();
    __syncthreads();
    wgrad_adam<W / 4, AS, WS, UPC>(dz, UPC, a2, sm + pW2, B, ak, dump,
                                   [&](int j, int c) { return o.oW2 + ((int)rank * UPC + j) * W + c; }, tid);
    bgrad_adam(dz, UPC, UPC, sm + pb2, B, ak, dump, o.ob2 + (int)rank * UPC, tid);
    __syncthreads();   // dz is rewritten below
    xwait(5);
    MD_STAMP();
    gather_recv(recv + CL * BMAX * UPC, code + BMAX * UPC, dz, B, tid);
    __syncthreads();
    partial_scatter<true>(dz, UPC, UPC, sm + pW1, WS, 0, B, recv_s, rank, xb(6), tid);                // exchange 6 -> recv[0]
    MD_STAMP();
    __syncthreads();
    wgrad_adam<W / 4, AS, WS, UPC>(dz, UPC, a1, sm + pW1, B, ak, dump,
                                   [&](int j, int c) { return o.oW1 + ((int)rank * UPC + j) * W + c; }, tid);
    bgrad_adam(dz, UPC, UPC, sm + pb1, B, ak, dump, o.ob1 + (int)rank * UPC, tid);
    __syncthreads();
    xwait(6);
    MD_STAMP();
    // ---- layer 1: dZ1; the gradient of the embedding inputs goes back to the column owners; dW0 over the whole input row
    gather_recv(recv, code, dz, B, tid);
    __syncthreads();
    partial_scatter<true>(dz, UPC, UPC, sm + pW0, W0S, DP, B, recv_s + CL * BMAX * UPC * 4, rank, xb(7), tid);   // exchange 7 -> recv[1]
    MD_STAMP();
    __syncthreads();
    wgrad_adam<K0P / 4, C0S, W0S, UPC>(dz, UPC, cat0, sm + pW0, B, ak, dump,
                                       [&](int j, int c) { const int gc = w0_col(c, D); return gc >= 0 ? o.oW0 + ((int)rank * UPC + j) * o.K0 + gc : -1; }, tid);
    bgrad_adam(dz, UPC, UPC, sm + pb0, B, ak, dump, o.ob0 + (int)rank * UPC, tid);
    if (more) publish_idx(par ^ 1);   // the next step's frames (prefetched long ago)
    xwait(7);
    MD_STAMP();
    // ---- embedding: gradient rows of the batch; dense Adam over all T rows, as torch's.  Now: the rows the next step gathers
    // (the rest follows behind that step's forward hand-overs); last step of the launch: all rows.
    gather_recv(recv + CL * BMAX * UPC, nullptr, dE, B, tid);
    __syncthreads();
    if (more) {
      const int* ixn = idxb + (par ^ 1) * BMAX;
      if (tid < 256) {
        const int r = tid >> 4, e = tid & 15;
        if (r < B) {
          bool first = true;
          for (int q = 0; q < r; ++q) first = first && ixn[q] != ixn[r];
          if (first) {
            emb_item(ixn[r], e, par, ak, nullptr);
            if (e == 0) need[ixn[r]] = 1;
          }
        }
      }
    } else {
      for (int it = tid; it < T * UPC; it += NT) emb_item(it >> 4, it & 15, par, ak, dump);
    }
    __syncthreads();
    MD_STAMP();
  }
  move_params(a, o, sm, rank, n_own, 1, tid);
  cluster_sync_all();   // no CTA exits while a peer may still push into it
}
#undef MD_STAMP

}  // namespace tmd
}  // namespace idff

using namespace idff;
using namespace idff::tmd;

extern "C" int idff_debug_train_md_stamps(int64_t* h_out32) {
  IDFF_CHECK_ARG(h_out32, "idff_debug_train_md_stamps: null pointer");
  IDFF_CUDA(cudaMemcpyFromSymbol(h_out32, g_md_stamps, sizeof(long long) * 32));
  return IDFF_OK;
}

static int check_md_shape(int32_t D, int32_t T, const char* who) {
  IDFF_CHECK_ARG(D >= 1 && D <= DMAX, "%s: state dimension %d outside [1, %d]", who, D, DMAX);
  IDFF_CHECK_ARG(T >= 2 && T <= TMAX, "%s: %d frames outside [2, %d]", who, T, TMAX);
  return IDFF_OK;
}

extern "C" int64_t idff_train_md_state_floats(int32_t D, int32_t T) {
  if (check_md_shape(D, T, "idff_train_md_state_floats")) return -1;
  const Off o(D, T);
  return 3 * (int64_t)((o.NP + 63) / 64 * 64);
}

extern "C" int idff_train_md_offsets(int32_t D, int32_t T, int64_t* h_out13) {
  IDFF_CHECK_ARG(h_out13, "idff_train_md_offsets: null pointer");
  int rc = check_md_shape(D, T, "idff_train_md_offsets");
  if (rc) return rc;
  const Off o(D, T);
  const int64_t npp = (o.NP + 63) / 64 * 64;
  const int64_t v[13] = {o.oE, o.oW0, o.ob0, o.oW1, o.ob1, o.oW2, o.ob2, o.oW3, o.ob3, o.NP, 0, npp, 2 * npp};
  for (int i = 0; i < 13; ++i) h_out13[i] = v[i];
  return IDFF_OK;
}

extern "C" int idff_train_md_steps(float* d_state, const float* d_dataset, int32_t T, int32_t D, const int64_t* d_idx,
                                   const float* d_t, const float* d_eps, int32_t n_steps, int32_t B, float sigma,
                                   int32_t sigma_mode, const idff_adamw_t* opt, int64_t steps_done, float* d_loss,
                                   float* d_grad_dump, void* stream) {
  return idff_train_md_steps_paired(d_state, d_dataset, T, D, d_idx, nullptr, nullptr, d_t, d_eps, n_steps, B, sigma, sigma_mode, opt,
                                    steps_done, d_loss, d_grad_dump, stream);
}

// The same loop with the batch re-paired before the path sample, as the `_dynamic` matcher's sample_plan does
// (conditional_flow_matching_dynamic.py:268): x0 = dataset[d_row0], x1 = dataset[d_row1] ([n_steps][B] each, NULL = idx - 1 / idx),
// while the embedding still sees index_sel = d_idx in its original order (MD_simulation.py:135).
extern "C" int idff_train_md_steps_paired(float* d_state, const float* d_dataset, int32_t T, int32_t D, const int64_t* d_idx,
                                          const int64_t* d_row0, const int64_t* d_row1, const float* d_t, const float* d_eps,
                                          int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode, const idff_adamw_t* opt,
                                          int64_t steps_done, float* d_loss, float* d_grad_dump, void* stream) {
  int rc = check_md_shape(D, T, "idff_train_md_steps");
  if (rc) return rc;
  IDFF_CHECK_ARG(d_state && (reinterpret_cast<uintptr_t>(d_state) & 15) == 0, "idff_train_md_steps: state must be a 16-byte aligned device pointer");
  IDFF_CHECK_ARG(n_steps >= 0 && n_steps <= (1 << 20), "idff_train_md_steps: n_steps %d outside [0, 2^20]", n_steps);
  if (n_steps == 0) return IDFF_OK;
  IDFF_CHECK_ARG(d_dataset && d_idx && d_t && opt, "idff_train_md_steps: null pointer");
  IDFF_CHECK_ARG(B >= 1 && B <= BMAX, "idff_train_md_steps: batch %d outside [1, %d]", B, BMAX);
  IDFF_CHECK_ARG(sigma_mode == 0 || sigma_mode == 1, "idff_train_md_steps: sigma_mode %d", sigma_mode);
  IDFF_CHECK_ARG(steps_done >= 0, "idff_train_md_steps: steps_done < 0");
  IDFF_CHECK_ARG(opt->weight_decay == 0.0 && opt->max_grad_norm <= 0.0,
                 "idff_train_md_steps: the PolyALA loop uses torch.optim.Adam without weight decay or clipping (MD_simulation.py:119,143-144)");
  PtrDeviceGuard guard(d_state);
  int cl = 0;
  IDFF_CUDA(cudaDeviceGetAttribute(&cl, cudaDevAttrClusterLaunch, guard.dev));
  if (!cl) {
    set_error("idff_train_md_steps: the device does not support thread-block clusters");
    return IDFF_EUNSUPPORTED;
  }
  IDFF_CUDA(cudaFuncSetAttribute(k_train_md, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const Off o(D, T);
  Args a;
  memset(&a, 0, sizeof(a));
  a.S = d_state; a.dataset = d_dataset; a.idx = reinterpret_cast<const long long*>(d_idx); a.t = d_t; a.eps = d_eps;
  a.row0 = reinterpret_cast<const long long*>(d_row0); a.row1 = reinterpret_cast<const long long*>(d_row1);
  a.loss_out = d_loss; a.grad_dump = d_grad_dump;
  a.B = B; a.D = D; a.T = T; a.n_steps = n_steps;
  a.npp = (o.NP + 63) / 64 * 64;
  a.sigma = sigma; a.sigma_mode = sigma_mode;
  a.lr = opt->lr; a.beta1 = opt->beta1; a.beta2 = opt->beta2;
  a.b1pow = std::pow(opt->beta1, (double)steps_done); a.b2pow = std::pow(opt->beta2, (double)steps_done);
  a.omb1 = (float)(1.0 - opt->beta1); a.omb2 = (float)(1.0 - opt->beta2); a.beta2f = (float)opt->beta2; a.eps_adam = (float)opt->eps;
  k_train_md<<<CL, NT, kSmemBytes, (cudaStream_t)stream>>>(a);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}
