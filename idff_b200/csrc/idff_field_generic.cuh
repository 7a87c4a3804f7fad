// Generic IDFF field evaluation (device code shared by the generic eval / rk4 / sde kernels): any hidden
// width (multiple of 32, <= 2048), any state dimension.
//
// One CTA owns a group of PB particles for the whole solve.  Every hidden layer is a skinny GEMM
//   out[n][r] = sum_k Wt[k][n] * act[k][r]        r = jet*PB + particle
// with one thread per hidden unit n (weights read coalesced from L2, activations broadcast from shared
// memory as float4).  The higher-order terms of the reference (nested autograd.grad with ones as
// grad_outputs, Autograd_example_order2.py:63-70) are computed as forward Taylor jets (z, z', z'') along
// d = (1,..,1) plus ONE reverse sweep of  c1*phi + c2*D_d phi + c3*D_d^2 phi,  phi = sum_k out_k:
// the per-unit SELU derivative data stays with the thread that owns the unit (registers when the width
// is <= 256, local memory otherwise).
//
// This is the shape-agnostic path (IDFF_IMPL_GENERIC); the persistent warp-tiled kernel in
// idff_sampler_fused.cu serves the benchmark shapes.
#pragma once
#include "idff_common.cuh"

namespace idff {

struct FieldCfg {
  int variant;   // IDFF_FIELD_* or kRaw
  int D;         // state dimension
  int tidx;      // folded-bias row
  int reverse;   // interior evaluations need the reverse sweep (some c != 0)
};
constexpr int kRaw = 100;  // plain MLP forward on caller-provided input rows

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[r] += sum_k wcol[k*ld] * act[k*RT + r]
template <int RT>
__device__ __forceinline__ void gemm_col(const float* __restrict__ wcol, int ld, const float* __restrict__ act,
                                         int K, float (&acc)[RT]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float wv = __ldg(wcol + (size_t)k * ld);
    const float4* a4 = reinterpret_cast<const float4*>(act + (size_t)k * RT);
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) {
      const float4 a = a4[q];
      acc[4 * q + 0] = fmaf(a.x, wv, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(a.y, wv, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(a.z, wv, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(a.w, wv, acc[4 * q + 3]);
    }
  }
}

// out[p][d] = bias[d] + sum_n M[d*w + n] * act[n*RT + p]   (one warp per output, lanes over n)
template <int PB>
__device__ __forceinline__ void small_out(const float* __restrict__ M, int w, int nd, const float* __restrict__ act,
                                          int RT, float* __restrict__ out, const float* __restrict__ bias) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o = warp; o < PB * nd; o += nwarps) {
    const int p = o / nd, d = o - p * nd;
    float s = 0.0f;
    for (int n = lane; n < w; n += 32) s = fmaf(__ldg(M + (size_t)d * w + n), act[(size_t)n * RT + p], s);
    s = warp_sum(s);
    if (lane == 0) out[p * nd + d] = s + (bias ? __ldg(bias + d) : 0.0f);
  }
}

// (P,Q,R) = adjoints of (z, z', z'') given adjoints (A0,A1,A2) of (a, a', a'') with
//   a = s(z), a' = s'(z) z', a'' = E z'^2 + s'(z) z''
template <int NJA>
__device__ __forceinline__ void adjoint_unit(float code, float zd, float zdd, float A0, float A1, float A2,
                                             float& P, float& Q, float& R) {
  float s1, E;
  selu_decode(code, s1, E);
  P = A0 * s1;
  Q = 0.0f;
  R = 0.0f;
  if (NJA >= 2) {
    P = fmaf(A1 * E, zd, P);
    Q = A1 * s1;
  }
  if (NJA >= 3) {
    P = fmaf(A2 * E, fmaf(zd, zd, zdd), P);
    Q = fmaf(2.0f * A2 * E, zd, Q);
    R = A2 * s1;
  }
}

// One evaluation of the field for the CTA's PB particles.  xs/x0s/fs live in shared memory ([PB][D]).
// NJA = number of active jets, REV = run the reverse sweep (interior evaluation with a non-zero coefficient),
// MULTI = a thread owns several hidden units (width > blockDim).
template <int NJA, bool REV, int PB, bool MULTI>
__device__ __noinline__ void eval_field(const NetView& net, const FieldCfg& cfg, const Stage& st, const float* xs,
                                        const float* rawin, float* x0s, float* fs, float* actA, float* actB,
                                        float* pred, float* gout) {
  constexpr int RT = NJA * PB;
  constexpr int NPT = MULTI ? 8 : 1;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int D = cfg.D, w = net.w;
  const int nblk = MULTI ? (w + nth - 1) / nth : 1;

  // ---- layer-0 input rows: jet 0 = [x, t] (score-head network, din = 2 D + 1: [x, x, t]); jet 1 = d = (1..1, 0); jet 2 = 0
  for (int i = tid; i < net.din_pad * RT; i += nth) {
    const int k = i / RT, r = i - k * RT;
    const int j = r / PB, p = r - j * PB;
    float v = 0.0f;
    if (k < net.din) {
      if (j == 0) {
        if (cfg.variant == kRaw) v = rawin[p * net.din + k];
        else if (net.din == 2 * D + 1) v = (k < 2 * D) ? xs[p * D + (k % D)] : st.t;   // score-head network: cat[x, x, t]
        else v = (k < D) ? xs[p * D + k] : st.t;
      } else if (j == 1) {
        v = (k < D) ? 1.0f : 0.0f;
      }
    }
    actA[i] = v;
  }
  __syncthreads();

  float saved[NPT][3][REV ? RT : 1];   // per owned unit and layer: derivative code, z', z'' of each particle
  float* ain = actA;
  float* aout = actB;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const float* Wt = l == 0 ? net.wt0 : (l == 1 ? net.wt1 : net.wt2);
    const float* bias = l == 0 ? net.b0 + (size_t)cfg.tidx * w : (l == 1 ? net.b1 : net.b2);
    const int K = l == 0 ? net.din_pad : w;
#pragma unroll 1
    for (int nb = 0; nb < nblk; ++nb) {
      const int n = nb * nth + tid;
      if (n < w) {
        float acc[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) acc[r] = 0.0f;
        gemm_col<RT>(Wt + n, w, ain, K, acc);
        const float b = __ldg(bias + n);
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          float code, s1, E;
          const float a = selu_fwd(acc[p] + b, code);
          selu_decode(code, s1, E);
          if (REV) saved[nb][l][p] = code;
          aout[(size_t)n * RT + p] = a;
          if (NJA >= 2) {
            const float zd = acc[PB + p];
            if (REV) saved[nb][l][PB + p] = zd;
            aout[(size_t)n * RT + PB + p] = s1 * zd;
            if (NJA >= 3) {
              const float zdd = acc[2 * PB + p];
              if (REV) saved[nb][l][2 * PB + p] = zdd;
              aout[(size_t)n * RT + 2 * PB + p] = fmaf(E * zd, zd, s1 * zdd);
            }
          }
        }
      }
    }
    __syncthreads();
    float* tmp = ain; ain = aout; aout = tmp;
  }

  // ---- output layer on jet 0
  small_out<PB>(net.w3, w, net.dout, ain, RT, pred, net.b3);
  __syncthreads();

  if (REV) {
    // seeds: adjoints of (a3, a3', a3'') are c_i * u3[n]
#pragma unroll 1
    for (int nb = 0; nb < nblk; ++nb) {
      const int n = nb * nth + tid;
      if (n < w) {
        const float u = __ldg(net.u3 + n);
        const float A0 = st.c[0] * u, A1 = st.c[1] * u, A2 = st.c[2] * u;
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          float P, Q, R;
          adjoint_unit<NJA>(saved[nb][2][p], NJA >= 2 ? saved[nb][2][(NJA >= 2 ? PB : 0) + p] : 0.0f,
                            NJA >= 3 ? saved[nb][2][(NJA >= 3 ? 2 * PB : 0) + p] : 0.0f, A0, A1, A2, P, Q, R);
          aout[(size_t)n * RT + p] = P;
          if (NJA >= 2) aout[(size_t)n * RT + PB + p] = Q;
          if (NJA >= 3) aout[(size_t)n * RT + 2 * PB + p] = R;
        }
      }
    }
    __syncthreads();
    { float* tmp = ain; ain = aout; aout = tmp; }
#pragma unroll
    for (int l = 1; l >= 0; --l) {
      const float* Wn = l == 1 ? net.w2 : net.w1;   // [out][in]: column k of the transposed product
#pragma unroll 1
      for (int nb = 0; nb < nblk; ++nb) {
        const int k = nb * nth + tid;
        if (k < w) {
          float acc[RT];
#pragma unroll
          for (int r = 0; r < RT; ++r) acc[r] = 0.0f;
          gemm_col<RT>(Wn + k, w, ain, w, acc);
#pragma unroll
          for (int p = 0; p < PB; ++p) {
            float P, Q, R;
            adjoint_unit<NJA>(saved[nb][l][p], NJA >= 2 ? saved[nb][l][(NJA >= 2 ? PB : 0) + p] : 0.0f,
                              NJA >= 3 ? saved[nb][l][(NJA >= 3 ? 2 * PB : 0) + p] : 0.0f, acc[p],
                              NJA >= 2 ? acc[(NJA >= 2 ? PB : 0) + p] : 0.0f,
                              NJA >= 3 ? acc[(NJA >= 3 ? 2 * PB : 0) + p] : 0.0f, P, Q, R);
            aout[(size_t)k * RT + p] = P;
            if (l == 1) {
              if (NJA >= 2) aout[(size_t)k * RT + PB + p] = Q;
              if (NJA >= 3) aout[(size_t)k * RT + 2 * PB + p] = R;
            }
          }
        }
      }
      __syncthreads();
      float* tmp = ain; ain = aout; aout = tmp;
    }
    // gradient wrt x: g[p][d] = sum_n W0[n][d] * P1[n][p]
    small_out<PB>(net.wt0, w, D, ain, RT, gout, nullptr);
    __syncthreads();
  }

  // ---- combine (reference operation order: (a*(pred-x))/(1-t) + momentum)
  if (cfg.variant != kRaw) {
    for (int o = tid; o < PB * D; o += nth) {
      const int p = o / D, d = o - p * D;
      const float x = xs[o];
      const float pr = pred[p * net.dout + d];
      float f;
      if (cfg.variant == IDFF_FIELD_CFM) {
        if (st.kind == 0) x0s[o] = x;
        f = __fsub_rn(pr, x0s[o]);
      } else {
        const float diff = __fsub_rn(pr, x);
        if (st.kind == 0) f = diff;
        else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
        else {
          f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
          if (cfg.variant == IDFF_FIELD_SCOREHEAD)
            f = __fadd_rn(__fadd_rn(f, __fmul_rn(st.c[0], pred[p * net.dout + D + d])), st.c[1]);
          else if (REV)
            f = __fadd_rn(f, gout[o]);
        }
      }
      fs[o] = f;
    }
  }
  __syncthreads();
}

template <int NJ, int PB, bool MULTI>
__device__ __forceinline__ void eval_dispatch(const NetView& net, const FieldCfg& cfg, const Stage& st,
                                              const float* xs, float* x0s, float* fs, float* actA, float* actB,
                                              float* pred, float* gout) {
  if (st.kind == 2 && cfg.reverse)
    eval_field<NJ, true, PB, MULTI>(net, cfg, st, xs, nullptr, x0s, fs, actA, actB, pred, gout);
  else
    eval_field<1, false, PB, MULTI>(net, cfg, st, xs, nullptr, x0s, fs, actA, actB, pred, gout);
}

struct SmemPlan {
  float *actA, *actB, *xs, *x0s, *fs, *y, *k1, *k2, *k3, *pred, *gout;
};
template <int NJ, int PB>
__device__ __forceinline__ SmemPlan carve(float* smem, const NetView& net, int D) {
  const int kmax = net.w > net.din_pad ? net.w : net.din_pad;
  const int act = kmax * NJ * PB;
  const int sd = (PB * (D > net.din ? D : net.din) + 3) & ~3;
  SmemPlan s;
  s.actA = smem; s.actB = s.actA + act;
  s.xs = s.actB + act; s.x0s = s.xs + sd; s.fs = s.x0s + sd; s.y = s.fs + sd;
  s.k1 = s.y + sd; s.k2 = s.k1 + sd; s.k3 = s.k2 + sd; s.gout = s.k3 + sd;
  s.pred = s.gout + sd;   // PB * dout
  return s;
}
inline size_t generic_smem_bytes(const NetView& net, int D, int NJ, int PB) {
  const int kmax = net.w > net.din_pad ? net.w : net.din_pad;
  const size_t act = (size_t)kmax * NJ * PB;
  const size_t sd = ((size_t)PB * (D > net.din ? D : net.din) + 3) & ~size_t(3);
  return sizeof(float) * (2 * act + 8 * sd + (((size_t)PB * net.dout + 3) & ~size_t(3)));
}

// ---- host-side planning shared by the three generic translation units ----------------------------------------
struct GenPlan {
  int nj, pb, multi, threads, grid;
  size_t smem;
};
int plan_generic(const idff_weights_t* w, int D, int nj, int64_t N, GenPlan* pl);

// KERNEL<NJ, PB, MULTI>: (PB, MULTI) is (8, false) for widths <= 256 and (4, true) above.
#define IDFF_GEN_DISPATCH(KERNEL, pl, stream, ...)                                                        \
  do {                                                                                                     \
    auto launch = [&](auto kern) -> int {                                                                  \
      IDFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pl).smem));  \
      kern<<<(pl).grid, (pl).threads, (pl).smem, (cudaStream_t)(stream)>>>(__VA_ARGS__);                   \
      IDFF_CUDA(cudaGetLastError());                                                                       \
      count_launch();                                                                                      \
      return IDFF_OK;                                                                                      \
    };                                                                                                     \
    switch ((pl).nj * 10 + (pl).multi) {                                                                   \
      case 10: return launch(KERNEL<1, 8, false>);                                                         \
      case 11: return launch(KERNEL<1, 4, true>);                                                          \
      case 20: return launch(KERNEL<2, 8, false>);                                                         \
      case 21: return launch(KERNEL<2, 4, true>);                                                          \
      case 30: return launch(KERNEL<3, 8, false>);                                                         \
      case 31: return launch(KERNEL<3, 4, true>);                                                          \
      default:                                                                                             \
        set_error("generic kernel: no instance for jets=%d multi=%d", (pl).nj, (pl).multi);                \
        return IDFF_EUNSUPPORTED;                                                                          \
    }                                                                                                      \
  } while (0)

}  // namespace idff
