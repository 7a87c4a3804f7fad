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
// barrier (monotonic counter, one arrival per CTA; the launch is cooperative so all CTAs are co-resident):
//   A  rows:    B/4 CTAs own 4 batch rows each and run path sample -> forward -> loss -> backward-data for their rows
//               with no cross-CTA traffic (thread = hidden unit; the K sum of every 256x256 layer is split over the two
//               halves of the CTA; weights stream from L2 with 32 loads in flight per thread).  Activations and the
//               pre-activation gradients dZ go to global memory (L2-resident, 1.5 MB).
//   B  weights: 137 jobs, one per CTA: 2 x 64 tiles (32 x 32) of dW1 / dW2 = dZ^T A (K = batch, operands staged in shared
//               memory, 2 x 2 outputs per thread, K split over the CTA halves), 4 + 4 jobs for the thin layers, one for
//               db3 and the loss value.  A gradient never leaves the registers of the thread that owns its parameter;
//               only the CTA's sum of squares goes to global memory.
//   C  update:  every CTA adds the 148 partial sums of squares in the same fixed order (-> identical clip coefficient),
//               then each thread applies clip + AdamW (torch's op order) to the parameters it owns and refreshes the
//               transposed forward copies of W1 / W2.
// Every reduction has a fixed order: results are bit-reproducible run to run.
#include <cmath>

#include "idff_common.cuh"

namespace idff {
namespace trn {

constexpr int W = 256, DIN = 3, DOUT = 2, R = 4, NT = 512, MAXB = IDFF_TRAIN_MAX_BATCH;
constexpr int oW0 = 0, ob0 = oW0 + W * DIN, oW1 = ob0 + W, ob1 = oW1 + W * W, oW2 = ob1 + W, ob2 = oW2 + W * W,
              oW3 = ob2 + W, ob3 = oW3 + DOUT * W, NP = ob3 + DOUT;
constexpr int NPP = (NP + 63) / 64 * 64;
constexpr int JOB_W0 = 128, JOB_W3 = 132, JOB_B3 = 136, N_JOBS = 137;
constexpr int MAX_GRID = 256;

// offsets (floats) into the caller-owned state buffer
constexpr size_t sP = 0, sM = sP + NPP, sV = sM + NPP, sWT1 = sV + NPP, sWT2 = sWT1 + W * W, sIN = sWT2 + W * W,
                 sA1 = sIN + (size_t)MAXB * 4, sA2 = sA1 + (size_t)MAXB * W, sA3 = sA2 + (size_t)MAXB * W,
                 sZ1 = sA3 + (size_t)MAXB * W, sZ2 = sZ1 + (size_t)MAXB * W, sZ3 = sZ2 + (size_t)MAXB * W,
                 sDO = sZ3 + (size_t)MAXB * W, sLP = sDO + (size_t)MAXB * DOUT, sNP = sLP + MAX_GRID, sBAR = sNP + MAX_GRID,
                 sTOTAL = sBAR + 64;

struct Args {
  float* S;   // state
  const float *x0, *x1, *t, *eps;
  float *loss_out, *gnorm_out, *grad_dump;
  int B, n_steps;
  long long steps_done;
  float sigma;
  int sigma_mode;
  double lr, beta1, beta2;                          // bias corrections are computed in double like torch's Python scalars
  float decay, omb1, omb2, beta2f, eps_adam, clip;  // 1 - lr*wd, 1 - beta1, 1 - beta2 (rounded to fp32 from double)
};

__device__ __forceinline__ float sigma_t_of(float t, float sigma, int mode) {
  if (mode == 0) return sigma;
  return __fmul_rn(sigma, __fsqrt_rn(__fmul_rn(t, __fsub_rn(1.0f, t))));   // conditional_flow_matching.py:231-233
}

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < epoch);
  }
  __syncthreads();
}

// acc[r] = sum over this half's 128 k of act[r][k] * Wm[k][j]   (Wm row-major [256][256], read through L2)
__device__ __forceinline__ void gemm_half(const float* __restrict__ Wm, const float* act, int h, int j, float (&acc)[R]) {
  const float* wp = Wm + (size_t)(h * 128) * W + j;
  const float* ap = act + h * 128;
  float wa[32], wb[32];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) wa[i] = __ldcg(wp + i * W);
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) {
    float(&cur)[32] = (blk & 1) ? wb : wa;
    float(&nxt)[32] = (blk & 1) ? wa : wb;
    if (blk < 3) {
#pragma unroll
      for (int i = 0; i < 32; ++i) nxt[i] = __ldcg(wp + ((blk + 1) * 32 + i) * W);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(ap + r * W + blk * 32 + i);
        acc[r] = fmaf(a.x, cur[i], acc[r]);
        acc[r] = fmaf(a.y, cur[i + 1], acc[r]);
        acc[r] = fmaf(a.z, cur[i + 2], acc[r]);
        acc[r] = fmaf(a.w, cur[i + 3], acc[r]);
      }
    }
  }
}

__device__ __forceinline__ float selu_d1(float code) { return code < 0.0f ? kSeluLambda : code; }

// ---- phase A: path sample, forward, loss, backward-data for rows [4 cta, 4 cta + 4) -------------------------------------
__device__ void phase_rows(const Args& a, int s, float* smem, int cta) {
  float* act = smem;                 // [R][W]
  float* part = smem + R * W;        // [2][R][W]
  float* in4 = part + 2 * R * W;     // [R][4]   xt0, xt1, t, 1
  float* dout = in4 + R * 4;         // [R][2]
  float* lterm = dout + R * DOUT;    // [R*DOUT]
  const int tid = threadIdx.x, h = tid >> 8, j = tid & 255, lane = tid & 31, warp = tid >> 5;
  const int row0 = cta * R;
  const size_t sb = (size_t)s * a.B;
  const float* P = a.S + sP;

  if (tid < R) {   // K2 inlined: xt = t*x1 + (1-t)*x0 + sigma_t*eps   (conditional_flow_matching.py:98-123, no contraction)
    const size_t row = sb + row0 + tid;
    const float t = a.t[row];
    const float st = sigma_t_of(t, a.sigma, a.sigma_mode);
    float xt[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const float x0 = a.x0[row * 2 + d], x1 = a.x1[row * 2 + d], e = a.eps ? a.eps[row * 2 + d] : 0.0f;
      const float mu = __fadd_rn(__fmul_rn(t, x1), __fmul_rn(__fsub_rn(1.0f, t), x0));
      xt[d] = __fadd_rn(mu, __fmul_rn(st, e));
    }
    const float4 v = make_float4(xt[0], xt[1], t, 1.0f);
    *reinterpret_cast<float4*>(in4 + tid * 4) = v;
    *reinterpret_cast<float4*>(a.S + sIN + (size_t)(row0 + tid) * 4) = v;
  }
  __syncthreads();

  float code1[2], code2[2], code3[2];
  {   // layer 1 (K = 3)
    const float w0 = __ldcg(P + oW0 + j * 3), w1 = __ldcg(P + oW0 + j * 3 + 1), w2 = __ldcg(P + oW0 + j * 3 + 2);
    const float b = __ldcg(P + ob0 + j);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = 2 * h + q;
      const float z = fmaf(in4[r * 4 + 2], w2, fmaf(in4[r * 4 + 1], w1, fmaf(in4[r * 4], w0, b)));
      const float av = selu_fwd(z, code1[q]);
      act[r * W + j] = av;
      a.S[sA1 + (size_t)(row0 + r) * W + j] = av;
    }
  }
  __syncthreads();

  float acc[R];
#pragma unroll 1
  for (int l = 0; l < 2; ++l) {   // layers 2, 3
    gemm_half(a.S + (l ? sWT2 : sWT1), act, h, j, acc);
#pragma unroll
    for (int r = 0; r < R; ++r) part[(h * R + r) * W + j] = acc[r];
    __syncthreads();
    const float b = __ldcg(P + (l ? ob2 : ob1) + j);
    float* Aout = a.S + (l ? sA3 : sA2);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = 2 * h + q;
      const float z = (part[r * W + j] + part[(R + r) * W + j]) + b;
      float c;
      const float av = selu_fwd(z, c);
      if (l) code3[q] = c; else code2[q] = c;
      act[r * W + j] = av;
      Aout[(size_t)(row0 + r) * W + j] = av;
    }
    __syncthreads();
  }

  // output layer + loss (Autograd_example_order2.py:105) + d loss / d vt
  if (warp < R * DOUT) {
    const int r = warp >> 1, o = warp & 1;
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < W / 32; ++i) sum = fmaf(act[r * W + lane + 32 * i], __ldcg(P + oW3 + o * W + lane + 32 * i), sum);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) {
      const float vt = sum + __ldcg(P + ob3 + o);
      const float t = in4[r * 4 + 2];
      const float x1v = a.x1[(sb + row0 + r) * 2 + o];
      const float wgt = __fdiv_rn(1.0f, __fsub_rn(1.01f, t));
      const float rr = __fmul_rn(wgt, __fsub_rn(vt, x1v));
      lterm[warp] = __fmul_rn(rr, rr);
      const float gk = 1.0f / (float)(a.B * DOUT);
      const float d = 2.0f * rr * wgt * gk;
      dout[r * DOUT + o] = d;
      a.S[sDO + (size_t)(row0 + r) * DOUT + o] = d;
    }
  }
  __syncthreads();
  if (tid == 0) {
    float ls = 0.0f;
#pragma unroll
    for (int i = 0; i < R * DOUT; ++i) ls += lterm[i];
    a.S[sLP + cta] = ls;
  }
  {   // dZ3 = (dOut W3) * selu'(z3)
    const float w30 = __ldcg(P + oW3 + j), w31 = __ldcg(P + oW3 + W + j);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = 2 * h + q;
      const float dz = fmaf(dout[r * 2 + 1], w31, dout[r * 2] * w30) * selu_d1(code3[q]);
      act[r * W + j] = dz;
      a.S[sZ3 + (size_t)(row0 + r) * W + j] = dz;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int l = 0; l < 2; ++l) {   // dZ2 = (dZ3 W2) * selu'(z2), dZ1 = (dZ2 W1) * selu'(z1)
    gemm_half(P + (l ? oW1 : oW2), act, h, j, acc);
#pragma unroll
    for (int r = 0; r < R; ++r) part[(h * R + r) * W + j] = acc[r];
    __syncthreads();
    float* Zout = a.S + (l ? sZ1 : sZ2);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = 2 * h + q;
      const float dz = (part[r * W + j] + part[(R + r) * W + j]) * selu_d1(l ? code1[q] : code2[q]);
      act[r * W + j] = dz;
      Zout[(size_t)(row0 + r) * W + j] = dz;
    }
    __syncthreads();
  }
}

// One thread's share of the parameters: up to 5 gradients and their flat indices (-1: none); tile jobs also own the
// transposed forward copy of their 2 x 2 block.
struct Owned {
  float g[5];
  int idx[5];
  int wt_off;   // >= 0: state offset of WT[k0][j0] for the 2 x 2 block g[0..3] = (j0,k0) (j0,k0+1) (j0+1,k0) (j0+1,k0+1)
};

// ---- phase B ---------------------------------------------------------------------------------------------------------------
__device__ void job_tile(const Args& a, int job, float* smem, Owned& own) {
  float* dZs = smem;               // [256][32]
  float* As = smem + 256 * 32;     // [256][32]
  float* red = As + 256 * 32;      // [256][4]
  const int tid = threadIdx.x;
  const int m = job >> 6, tile = job & 63, jb = tile >> 3, kb = tile & 7;
  const float* dZ = a.S + (m ? sZ3 : sZ2);
  const float* Ain = a.S + (m ? sA2 : sA1);
  const int gK = tid >> 8, jj = (tid & 255) >> 4, kk = tid & 15;
  float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f, gb = 0.f;
  for (int r0 = 0; r0 < a.B; r0 += 256) {
    const int rows = min(256, a.B - r0);
    for (int i = tid; i < rows * 8; i += NT) {
      const int r = i >> 3, c4 = i & 7;
      *reinterpret_cast<float4*>(dZs + r * 32 + c4 * 4) =
          __ldcg(reinterpret_cast<const float4*>(dZ + (size_t)(r0 + r) * W + jb * 32) + c4);
      *reinterpret_cast<float4*>(As + r * 32 + c4 * 4) =
          __ldcg(reinterpret_cast<const float4*>(Ain + (size_t)(r0 + r) * W + kb * 32) + c4);
    }
    __syncthreads();
    const int half = rows >> 1, rb = gK * half;
#pragma unroll 4
    for (int r = rb; r < rb + half; ++r) {
      const float2 dz = *reinterpret_cast<const float2*>(dZs + r * 32 + 2 * jj);
      const float2 av = *reinterpret_cast<const float2*>(As + r * 32 + 2 * kk);
      g00 = fmaf(dz.x, av.x, g00);
      g01 = fmaf(dz.x, av.y, g01);
      g10 = fmaf(dz.y, av.x, g10);
      g11 = fmaf(dz.y, av.y, g11);
    }
    if (kb == 0 && tid < 32) {
      for (int r = 0; r < rows; ++r) gb += dZs[r * 32 + tid];
    }
    __syncthreads();
  }
  if (gK == 1) *reinterpret_cast<float4*>(red + (tid - 256) * 4) = make_float4(g00, g01, g10, g11);
  __syncthreads();
  if (gK == 0) {
    const float4 o = *reinterpret_cast<const float4*>(red + tid * 4);
    const int j0 = jb * 32 + 2 * jj, k0 = kb * 32 + 2 * kk, oW = m ? oW2 : oW1;
    own.g[0] = g00 + o.x; own.g[1] = g01 + o.y; own.g[2] = g10 + o.z; own.g[3] = g11 + o.w;
    own.idx[0] = oW + j0 * W + k0; own.idx[1] = own.idx[0] + 1; own.idx[2] = own.idx[0] + W; own.idx[3] = own.idx[2] + 1;
    own.wt_off = (int)(m ? sWT2 : sWT1) + k0 * W + j0;
    if (kb == 0 && tid < 32) {
      own.g[4] = gb;
      own.idx[4] = (m ? ob2 : ob1) + jb * 32 + tid;
    }
  }
}

// thin layers: G[j][c] = sum_r X[r][j] * Y[r][c]; X = dZ1, Y = IN (c < 4, IN[r][3] = 1 -> db0) or X = A3, Y = dOut (c < 2)
template <int C>
__device__ void job_thin(const Args& a, const float* X, const float* Y, int q, float* smem, float (&out)[4]) {
  const int tid = threadIdx.x, jl = tid & 63, rg = tid >> 6, j = 64 * q + jl;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int r = rg; r < a.B; r += 8) {
    const float x = __ldcg(X + (size_t)r * W + j);
    if (C == 4) {
      const float4 y = __ldcg(reinterpret_cast<const float4*>(Y) + r);
      acc[0] = fmaf(x, y.x, acc[0]); acc[1] = fmaf(x, y.y, acc[1]); acc[2] = fmaf(x, y.z, acc[2]); acc[3] = fmaf(x, y.w, acc[3]);
    } else {
      const float2 y = __ldcg(reinterpret_cast<const float2*>(Y) + r);
      acc[0] = fmaf(x, y.x, acc[0]); acc[1] = fmaf(x, y.y, acc[1]);
    }
  }
  *reinterpret_cast<float4*>(smem + (rg * 64 + jl) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  __syncthreads();
  if (tid < 64) {
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = 0.0f;
    for (int g = 0; g < 8; ++g) {
      const float4 v = *reinterpret_cast<const float4*>(smem + (g * 64 + tid) * 4);
      out[0] += v.x; out[1] += v.y; out[2] += v.z; out[3] += v.w;
    }
  }
}

__global__ void __launch_bounds__(NT, 1) k_train_flow(const Args a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double sred[NT / 32];
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  unsigned epoch = 0;
  unsigned* bar = reinterpret_cast<unsigned*>(a.S + sBAR);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
  const int nA = a.B / R;
  float* P = a.S + sP;
  float* M = a.S + sM;
  float* V = a.S + sV;

  for (int s = 0; s < a.n_steps; ++s) {
    if (cta < nA) phase_rows(a, s, smem, cta);
    grid_barrier(bar, epoch);

    Owned own;
#pragma unroll
    for (int i = 0; i < 5; ++i) { own.g[i] = 0.0f; own.idx[i] = -1; }
    own.wt_off = -1;
    if (cta < JOB_W0) {
      job_tile(a, cta, smem, own);
    } else if (cta < JOB_W3) {   // dW0, db0 for units [64 q, 64 q + 64)
      float o[4];
      const int q = cta - JOB_W0;
      job_thin<4>(a, a.S + sZ1, a.S + sIN, q, smem, o);
      if (tid < 64) {
        const int j = 64 * q + tid;
#pragma unroll
        for (int c = 0; c < 3; ++c) { own.g[c] = o[c]; own.idx[c] = oW0 + j * 3 + c; }
        own.g[3] = o[3]; own.idx[3] = ob0 + j;
      }
    } else if (cta < JOB_B3) {   // dW3 for units [64 q, 64 q + 64)
      float o[4];
      const int q = cta - JOB_W3;
      job_thin<2>(a, a.S + sA3, a.S + sDO, q, smem, o);
      if (tid < 64) {
        const int j = 64 * q + tid;
        own.g[0] = o[0]; own.idx[0] = oW3 + j;
        own.g[1] = o[1]; own.idx[1] = oW3 + W + j;
      }
    } else if (cta == JOB_B3) {   // db3 and the loss value
      if (warp < 2) {
        float sum = 0.0f;
        for (int r = lane; r < a.B; r += 32) sum += __ldcg(a.S + sDO + (size_t)r * DOUT + warp);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) { own.g[0] = sum; own.idx[0] = ob3 + warp; }
      } else if (warp == 2) {
        double sum = 0.0;
        for (int c = lane; c < nA; c += 32) sum += (double)__ldcg(a.S + sLP + c);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0 && a.loss_out) a.loss_out[s] = (float)(sum / (double)(a.B * DOUT));
      }
    }
    {   // this CTA's sum of squared gradients
      double ss = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) ss += (double)own.g[i] * (double)own.g[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      if (lane == 0) sred[warp] = ss;
      __syncthreads();
      if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < NT / 32; ++i) tot += sred[i];
        a.S[sNP + cta] = (float)tot;
      }
    }
    grid_barrier(bar, epoch);

    if (warp == 0) {   // clip_grad_norm_(params, clip): coef = min(1, clip / (norm + 1e-6))
      double tot = 0.0;
      for (int c = lane; c < (int)gridDim.x; c += 32) tot += (double)__ldcg(a.S + sNP + c);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
      if (lane == 0) {
        const float norm = (float)sqrt(tot);
        s_coef = a.clip > 0.0f ? fminf(1.0f, a.clip / (norm + 1e-6f)) : 1.0f;
        if (cta == 0 && a.gnorm_out) a.gnorm_out[s] = norm;
        const double stepn = (double)(a.steps_done + s + 1);
        s_step_size = (float)(a.lr / (1.0 - pow(a.beta1, stepn)));
        s_bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2, stepn));
      }
    }
    __syncthreads();
    {   // AdamW, torch/optim/adamw.py op order
      const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
      const float decay = a.decay, omb1 = a.omb1, omb2 = a.omb2;
      float pn[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        pn[i] = 0.0f;
        const int ix = own.idx[i];
        if (ix < 0) continue;
        if (a.grad_dump && s == a.n_steps - 1) a.grad_dump[ix] = own.g[i];
        const float g = own.g[i] * coef;
        float p = __ldcg(P + ix), mm = __ldcg(M + ix), vv = __ldcg(V + ix);
        p = p * decay;
        mm = fmaf(g - mm, omb1, mm);
        vv = fmaf(omb2 * g, g, vv * a.beta2f);
        const float denom = sqrtf(vv) / bc2_sqrt + a.eps_adam;
        p = p - step_size * (mm / denom);
        P[ix] = p; M[ix] = mm; V[ix] = vv;
        pn[i] = p;
      }
      if (own.wt_off >= 0) {
        *reinterpret_cast<float2*>(a.S + own.wt_off) = make_float2(pn[0], pn[2]);
        *reinterpret_cast<float2*>(a.S + own.wt_off + W) = make_float2(pn[1], pn[3]);
      }
    }
    grid_barrier(bar, epoch);
  }
}

__global__ void k_train_init(float* S) {   // WT[k][j] = W[j][k] for layers 2 and 3
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W * W) {
    const int k = i / W, j = i % W;
    S[sWT1 + i] = S[sP + oW1 + j * W + k];
    S[sWT2 + i] = S[sP + oW2 + j * W + k];
  }
}

constexpr size_t kSmemBytes = (2 * 256 * 32 + 256 * 4) * sizeof(float);   // phase B tiles (phase A needs 12.2 KB)

}  // namespace trn
}  // namespace idff

using namespace idff;
using namespace idff::trn;

extern "C" int64_t idff_train_flow_state_floats(void) { return (int64_t)sTOTAL; }

extern "C" int idff_train_flow_offsets(int64_t* h_out12) {
  IDFF_CHECK_ARG(h_out12, "idff_train_flow_offsets: null pointer");
  const int64_t v[12] = {oW0, ob0, oW1, ob1, oW2, ob2, oW3, ob3, NP, (int64_t)sP, (int64_t)sM, (int64_t)sV};
  for (int i = 0; i < 12; ++i) h_out12[i] = v[i];
  return IDFF_OK;
}

extern "C" int idff_train_flow_init(float* d_state, void* stream) {
  IDFF_CHECK_ARG(d_state && (reinterpret_cast<uintptr_t>(d_state) & 15) == 0, "idff_train_flow_init: state must be a 16-byte aligned device pointer");
  k_train_init<<<(W * W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_state);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}

extern "C" int idff_train_flow_steps(float* d_state, const float* d_x0, const float* d_x1, const float* d_t,
                                     const float* d_eps, int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode,
                                     const idff_adamw_t* opt, int64_t steps_done, float* d_loss, float* d_gnorm,
                                     float* d_grad_dump, void* stream) {
  IDFF_CHECK_ARG(d_state && (reinterpret_cast<uintptr_t>(d_state) & 15) == 0, "idff_train_flow_steps: state must be a 16-byte aligned device pointer");
  IDFF_CHECK_ARG(d_x0 && d_x1 && d_t && opt, "idff_train_flow_steps: null pointer");
  IDFF_CHECK_ARG(n_steps >= 0 && n_steps <= (1 << 20), "idff_train_flow_steps: n_steps %d outside [0, 2^20]", n_steps);
  IDFF_CHECK_ARG(B >= R && B <= MAXB && B % R == 0, "idff_train_flow_steps: batch %d must be a multiple of %d in [%d, %d]", B, R, R, MAXB);
  IDFF_CHECK_ARG(sigma_mode == 0 || sigma_mode == 1, "idff_train_flow_steps: sigma_mode %d", sigma_mode);
  IDFF_CHECK_ARG(steps_done >= 0, "idff_train_flow_steps: steps_done < 0");
  if (n_steps == 0) return IDFF_OK;
  int dev = 0, sms = 0, coop = 0;
  IDFF_CUDA(cudaGetDevice(&dev));
  IDFF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDFF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop || sms < N_JOBS) {
    set_error("idff_train_flow_steps: needs cooperative launch and >= %d SMs (device has %d)", N_JOBS, sms);
    return IDFF_EUNSUPPORTED;
  }
  const int grid = sms < MAX_GRID ? sms : MAX_GRID;
  static bool attr_set = false;
  if (!attr_set) {
    IDFF_CUDA(cudaFuncSetAttribute(k_train_flow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  Args a;
  a.S = d_state;
  a.x0 = d_x0; a.x1 = d_x1; a.t = d_t; a.eps = d_eps;
  a.loss_out = d_loss; a.gnorm_out = d_gnorm; a.grad_dump = d_grad_dump;
  a.B = B; a.n_steps = n_steps; a.steps_done = steps_done;
  a.sigma = sigma; a.sigma_mode = sigma_mode;
  a.lr = opt->lr; a.beta1 = opt->beta1; a.beta2 = opt->beta2;
  a.decay = (float)(1.0 - opt->lr * opt->weight_decay);
  a.omb1 = (float)(1.0 - opt->beta1); a.omb2 = (float)(1.0 - opt->beta2);
  a.beta2f = (float)opt->beta2; a.eps_adam = (float)opt->eps; a.clip = (float)opt->max_grad_norm;
  cudaStream_t st = (cudaStream_t)stream;
  IDFF_CUDA(cudaMemsetAsync(d_state + sBAR, 0, 64 * sizeof(float), st));
  void* kargs[] = {&a};
  IDFF_CUDA(cudaLaunchCooperativeKernel((const void*)k_train_flow, dim3(grid), dim3(NT), kargs, kSmemBytes, st));
  count_launch();
  return IDFF_OK;
}
