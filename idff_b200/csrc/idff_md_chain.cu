// The autoregressive PolyALA generator in ONE launch.
//
// Reference: TrajectoryGenerator.generate_trajectories, timeseris_example/MD_simulation.py:197-213 -- for t_idx in range(T):
// SDE_cal(model, gamma2 = gamma2 / (1 + t_idx), init_ind = t_idx, dt) ; x1 = sdeint(sde, x1, ts, dt)[-1].  T sequential solves
// of 12 drift evaluations each at 5 trajectories: pure latency.  The throughput kernel (idff_sampler_fused.cu) needs 0.5 ms per
// frame whatever the trajectory count, because one warp owns a particle group for the whole solve.  Here one CTA owns 4
// trajectories for the whole CHAIN and all 512 threads cooperate on every layer:
//   * the network lives in shared memory for the whole launch (W1, W2 once, as stored, with a padded row pitch so that the
//     forward product and the reverse product -- the transposed read of the same copy -- are (nearly) bank-conflict free;
//     W0^T and W3 likewise): 220 KB, loaded once;
//   * a layer pass is a skinny GEMM, 128 units x 8 columns (4 trajectories x 2 jets): one N = 8 tile of mma.sync.m16n8k8 with
//     the 3 x tf32 split (default, eval_mma), or fp32 FMAs with thread = hidden unit (eval, diagnostic); Taylor jets and the
//     reverse sweep as in idff_field_generic.cuh (the SELU derivative data stays in the registers of the unit's thread);
//   * the SRK / Euler update, the per-frame layer-0 bias (folded embedding row) and stage scalars (gamma2 / (1 + frame)) and the
//     hand-over of frame f's result to frame f + 1 all happen inside the kernel.
#include "idff_field_generic.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);

namespace chain {

constexpr int W = 128, LD = 129, LDM = 132, PB = 4, NT = 512, KS = NT / W, DMAX = 64;   // KS: K-slices of a layer pass (FMA path)
// LD / LDM: shared-memory row pitch of the weight matrices for the FMA path (bank = row + col) / the mma.sync path (bank = 4 row + col)

struct Args {
  int n_frames, n_steps, method, D, din, din_pad, dout, reverse;
  float h[kMaxSdeSteps];
  float gval[kMaxSdeSteps][4];
  float w0, w1;            // interpolation weights of the last output (torchsde linear_interp at ts[-1])
  const Stage* stages;     // device [n_frames][n_steps][3]
  const float *x_init, *Ik, *Ik0;
  float* y_hat;
  long K;                  // all trajectories (the row count of x_init / Ik / y_hat): n_sets * Kset
  // parameter sweep (idff_md_generate_sweep): rows [set * Kset, (set + 1) * Kset) run under parameter set `set`, whose stage table
  // is stages[set][n_frames][n_steps][3] and whose diffusion values are gval_sets[set][kMaxSdeSteps][4] (NULL: one set, gval above)
  int n_sets;
  long Kset;
  const float* gval_sets;
};

struct Smem {
  float *W1, *W2, *wt0, *w3, *b0, *b1, *b2, *b3, *u3, *actA, *actB, *part, *opart, *y, *xs, *fs, *k1, *k2, *yprev, *pred, *gout;
  float* gv;   // [kMaxSdeSteps][4] diffusion values of the tile's parameter set
  Stage* st;
};
constexpr int kStateFloats = PB * DMAX;
// rows0 / rows3: rows of the W0^T / W3 images (the mma.sync path rounds them up to 16 and zero-fills); ld: row pitch
__host__ __device__ inline size_t smem_floats(int rows0, int rows3, int ld) {
  return 2 * W * ld + (size_t)(rows0 * ld + 3) / 4 * 4 + (size_t)(rows3 * ld + 3) / 4 * 4 + 4 * W + DMAX +
         (2 + KS - 1) * W * 2 * PB + 2 * NT + 8 * kStateFloats + kMaxSdeSteps * 4;
}
inline size_t smem_bytes(int rows0, int rows3, int ld) { return smem_floats(rows0, rows3, ld) * 4 + kMaxSdeSteps * 3 * sizeof(Stage); }

__device__ __forceinline__ Smem carve(float* m, int rows0, int rows3, int ld) {
  Smem s;
  s.W1 = m; m += W * ld;
  s.W2 = m; m += W * ld;
  s.wt0 = m; m += (rows0 * ld + 3) / 4 * 4;   // keep every region 16-byte aligned
  s.w3 = m; m += (rows3 * ld + 3) / 4 * 4;
  s.b0 = m; m += W; s.b1 = m; m += W; s.b2 = m; m += W; s.u3 = m; m += W;
  s.b3 = m; m += DMAX;
  s.actA = m; m += W * 2 * PB; s.actB = m; m += W * 2 * PB; s.part = m; m += (KS - 1) * W * 2 * PB;
  s.opart = m; m += 2 * NT;
  s.y = m; m += kStateFloats; s.xs = m; m += kStateFloats; s.fs = m; m += kStateFloats; s.k1 = m; m += kStateFloats;
  s.k2 = m; m += kStateFloats; s.yprev = m; m += kStateFloats; s.pred = m; m += kStateFloats; s.gout = m; m += kStateFloats;
  s.gv = m; m += kMaxSdeSteps * 4;
  s.st = reinterpret_cast<Stage*>(m);
  return s;
}

// acc[r] += sum_{k < K} wcol[k * ld] * act[k * RT + r]    (weights and activations in shared memory)
template <int RT>
__device__ __forceinline__ void gemm_col_s(const float* wcol, int ld, const float* act, int K, float (&acc)[RT]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float wv = wcol[k * ld];
    const float4* a4 = reinterpret_cast<const float4*>(act + k * RT);
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

// One evaluation of SDE_cal.f (MD_simulation.py:497-526) for the CTA's 4 trajectories: xs -> fs.
template <int NJA, bool REV>
__device__ __forceinline__ void eval(const Smem& s, const Args& A, const Stage& st) {
  constexpr int RT = NJA * PB;
  const int tid = threadIdx.x, hh = tid >> 7, n = tid & 127;
  const int D = A.D;
  // layer-0 input rows: jet 0 = [x, t]; jet 1 = d = (1..1, 0)
  for (int i = tid; i < A.din_pad * RT; i += NT) {
    const int k = i / RT, r = i - k * RT;
    const int j = r / PB, p = r - j * PB;
    float v = 0.0f;
    if (k < A.din) v = j == 0 ? (k < D ? s.xs[p * D + k] : st.t) : (k < D ? 1.0f : 0.0f);
    s.actA[i] = v;
  }
  __syncthreads();

  float saved[3][REV ? RT : 1];   // kept by the hh == 0 thread of unit n: derivative code and z' per trajectory
  float* ain = s.actA;
  float* aout = s.actB;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const int K = l == 0 ? A.din_pad : W, kh = K / KS, k0 = hh * kh;
    const float* Wl = l == 1 ? s.W1 : s.W2;
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.0f;
    if (l == 0) gemm_col_s<RT>(s.wt0 + k0 * LD + n, LD, ain + k0 * RT, kh, acc);
    else gemm_col_s<RT>(Wl + n * LD + k0, 1, ain + k0 * RT, kh, acc);
    if (hh > 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r) s.part[((hh - 1) * W + n) * RT + r] = acc[r];
    }
    __syncthreads();
    if (hh == 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int q = 0; q < KS - 1; ++q) acc[r] += s.part[(q * W + n) * RT + r];
      const float b = l == 0 ? s.b0[n] : (l == 1 ? s.b1[n] : s.b2[n]);
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        float code, s1, E;
        const float a = selu_fwd(acc[p] + b, code);
        selu_decode(code, s1, E);
        if (REV) saved[l][p] = code;
        aout[n * RT + p] = a;
        if (NJA >= 2) {
          const float zd = acc[PB + p];
          if (REV) saved[l][PB + p] = zd;
          aout[n * RT + PB + p] = s1 * zd;
        }
      }
    }
    __syncthreads();
    float* tmp = ain; ain = aout; aout = tmp;
  }

  // output layer on jet 0: pred[p][d] = b3[d] + sum_n W3[d][n] a3[n][p]   (thread = (trajectory, output))
  {
    const int o = tid & 255, half = tid >> 8;
    if (o < PB * A.dout) {
      const int p = o / A.dout, d = o - p * A.dout;
      const float* wr = s.w3 + d * LD + half * (W / 2);
      const float* ar = ain + half * (W / 2) * RT + p;
      float sum = 0.0f;
#pragma unroll 8
      for (int k = 0; k < W / 2; ++k) sum = fmaf(wr[k], ar[k * RT], sum);
      s.opart[tid] = sum;
    }
    __syncthreads();
    if (tid < PB * A.dout) s.pred[tid] = (s.opart[tid] + s.opart[256 + tid]) + s.b3[tid % A.dout];
  }
  __syncthreads();

  if (REV) {
    if (hh == 0) {   // seeds: adjoints of (a3, a3') are c1 * u3[n], c2 * u3[n]
      const float u = s.u3[n];
      const float A0 = st.c[0] * u, A1 = st.c[1] * u;
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        float P, Q, Rr;
        adjoint_unit<NJA>(saved[2][p], NJA >= 2 ? saved[2][(NJA >= 2 ? PB : 0) + p] : 0.0f, 0.0f, A0, A1, 0.0f, P, Q, Rr);
        aout[n * RT + p] = P;
        if (NJA >= 2) aout[n * RT + PB + p] = Q;
      }
    }
    __syncthreads();
    { float* tmp = ain; ain = aout; aout = tmp; }
#pragma unroll
    for (int l = 1; l >= 0; --l) {
      const float* Wl = l == 1 ? s.W2 : s.W1;   // adjoint of a_l[k] = sum_j W[j][k] P_{l+1}[j]: column k of the stored matrix
      const int k0 = hh * (W / KS);
      float acc[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = 0.0f;
      gemm_col_s<RT>(Wl + k0 * LD + n, LD, ain + k0 * RT, W / KS, acc);
      if (hh > 0) {
#pragma unroll
        for (int r = 0; r < RT; ++r) s.part[((hh - 1) * W + n) * RT + r] = acc[r];
      }
      __syncthreads();
      if (hh == 0) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
          for (int q = 0; q < KS - 1; ++q) acc[r] += s.part[(q * W + n) * RT + r];
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          float P, Q, Rr;
          adjoint_unit<NJA>(saved[l][p], NJA >= 2 ? saved[l][(NJA >= 2 ? PB : 0) + p] : 0.0f, 0.0f, acc[p],
                            NJA >= 2 ? acc[(NJA >= 2 ? PB : 0) + p] : 0.0f, 0.0f, P, Q, Rr);
          aout[n * RT + p] = P;
          if (l == 1 && NJA >= 2) aout[n * RT + PB + p] = Q;
        }
      }
      __syncthreads();
      float* tmp = ain; ain = aout; aout = tmp;
    }
    // gradient wrt x: g[p][d] = sum_n W0[n][d] P1[n][p]   (thread = (trajectory, coordinate))
    {
      const int o = tid & 255, half = tid >> 8;
      if (o < PB * D) {
        const int p = o / D, d = o - p * D;
        const float* wr = s.wt0 + d * LD + half * (W / 2);
        const float* ar = ain + half * (W / 2) * RT + p;
        float sum = 0.0f;
#pragma unroll 8
        for (int k = 0; k < W / 2; ++k) sum = fmaf(wr[k], ar[k * RT], sum);
        s.opart[tid] = sum;
      }
      __syncthreads();
      if (tid < PB * D) s.gout[tid] = s.opart[tid] + s.opart[256 + tid];
    }
    __syncthreads();
  }

  // combine, reference operation order (MD_simulation.py:506-526; the signs of the momentum terms are folded into c)
  if (tid < PB * D) {
    const int p = tid / D, d = tid - p * D;
    const float diff = __fsub_rn(s.pred[p * A.dout + d], s.xs[tid]);
    float f;
    if (st.kind == 0) f = diff;
    else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
    else {
      f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
      if (REV) f = __fadd_rn(f, s.gout[tid]);
    }
    s.fs[tid] = f;
  }
  __syncthreads();
}


// ---- the same evaluation on mma.sync.m16n8k8 (tf32 x 3: hi*hi + lo*hi + hi*lo, fp32 accumulate) --------------------------------
// A layer pass is C[128 units][8 columns] = W[128][128] X[128][8], column = 2 * trajectory + jet: exactly one N = 8 MMA tile, so
// the 16 warps are 8 unit tiles x 2 K halves.  The weights are read from shared memory once per pass (A fragments: 4 loads of 1
// wavefront per k-step at row pitch 132) instead of once per warp and k with broadcast activation loads: 3 x fewer shared-memory
// wavefronts and 3 x fewer instructions than the FMA path.  tcgen05 does not pay here: an N = 8 tile costs it the same ~70
// cycles per instruction as N = 128.  The C fragment gives a thread (unit g, trajectory t) and (unit g + 8, trajectory t) with
// both jets, so the SELU jet algebra and the reverse sweep's derivative data stay thread-local as in the other kernels.
// v = hi + lo with hi = v truncated to tf32 (what the tensor core reads of an fp32 register anyway) and lo = v - hi, exact in
// fp32; the hardware truncates lo to tf32 in turn (relative error 2^-10 of a term that is 2^-10 of v).  Two instructions per
// element: cvt.rna.tf32 is emulated on this part with a dozen.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A X over k-steps [ks0, ks0 + nks): A[m][k] = Wm[m * LDM + k] (TRANS = false) or Wm[k * LDM + m] (TRANS = true), m = m0 + ...
template <bool TRANS>
__device__ __forceinline__ void mma_pass(const float* Wm, const float* X, int m0, int ks0, int nks, int g, int t, float (&c)[4]) {
  // three independent accumulation chains (lo*hi, hi*lo, hi*hi): an mma.sync depends on its accumulator, so one chain of
  // 3 * nks instructions would be latency-bound
  float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int ks = ks0; ks < ks0 + nks; ++ks) {
    const int k0 = 8 * ks;
    float af[4];
    if (!TRANS) {
      const float* r0 = Wm + (m0 + g) * LDM + k0 + t;
      af[0] = r0[0]; af[1] = r0[8 * LDM]; af[2] = r0[4]; af[3] = r0[8 * LDM + 4];
    } else {
      const float* r0 = Wm + (k0 + t) * LDM + m0 + g;
      af[0] = r0[0]; af[1] = r0[8]; af[2] = r0[4 * LDM]; af[3] = r0[4 * LDM + 8];
    }
    const float bf0 = X[(k0 + t) * 8 + g], bf1 = X[(k0 + t + 4) * 8 + g];
    uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(af[i], ah[i], al[i]);
    split_tf32(bf0, bh0, bl0);
    split_tf32(bf1, bh1, bl1);
    mma_tf32(c1, al, bh0, bh1);
    mma_tf32(c2, ah, bl0, bl1);
    mma_tf32(c, ah, bh0, bh1);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] += c1[i] + c2[i];
}

template <bool REV>
__device__ __forceinline__ void eval_mma(const Smem& s, const Args& A, const Stage& st, bool jets) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int mt = warp & 7, kh = warp >> 3, m0 = 16 * mt;
  const int D = A.D, rows0 = (A.din_pad + 15) & ~15;
  // layer-0 input: X[k][2 p + jet]; jet 0 = [x, t], jet 1 = d = (1..1, 0); rows >= din are zero
  for (int i = tid; i < rows0 * 8; i += NT) {
    const int k = i >> 3, r = i & 7, p = r >> 1, j = r & 1;
    float v = 0.0f;
    if (k < A.din) v = j == 0 ? (k < D ? s.xs[p * D + k] : st.t) : ((jets && k < D) ? 1.0f : 0.0f);
    s.actA[i] = v;
  }
  __syncthreads();

  float saved[3][REV ? 4 : 1];   // per layer: (code, z') of (unit m0 + g, trajectory t) and of (unit m0 + g + 8, t); kh == 0 warps
  float* ain = s.actA;
  float* aout = s.actB;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (l == 0) mma_pass<true>(s.wt0, ain, m0, kh * (rows0 / 16), rows0 / 16, g, t, c);
    else mma_pass<false>(l == 1 ? s.W1 : s.W2, ain, m0, kh * 8, 8, g, t, c);
    if (kh == 1) {
      *reinterpret_cast<float2*>(s.part + (m0 + g) * 8 + 2 * t) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(s.part + (m0 + g + 8) * 8 + 2 * t) = make_float2(c[2], c[3]);
    }
    __syncthreads();
    if (kh == 0) {
      const float* bias = l == 0 ? s.b0 : (l == 1 ? s.b1 : s.b2);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n = m0 + g + 8 * q;
        const float2 o = *reinterpret_cast<const float2*>(s.part + n * 8 + 2 * t);
        const float z = c[2 * q] + o.x + bias[n], zd = c[2 * q + 1] + o.y;
        float code, s1, E;
        const float a = selu_fwd(z, code);
        selu_decode(code, s1, E);
        if (REV) { saved[l][2 * q] = code; saved[l][2 * q + 1] = zd; }
        *reinterpret_cast<float2*>(aout + n * 8 + 2 * t) = make_float2(a, s1 * zd);
      }
    }
    __syncthreads();
    float* tmp = ain; ain = aout; aout = tmp;
  }

  // output layer (jet-0 columns): pred[p][d] = b3[d] + sum_n W3[d][n] a3[n][p]; unit tiles 0 .. ceil(dout / 16) - 1
  {
    const int ntile = (A.dout + 15) >> 4;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (mt < ntile) {
      mma_pass<false>(s.w3, ain, m0, kh * 8, 8, g, t, c);
      if (kh == 1) {
        s.opart[(m0 + g) * 4 + t] = c[0];
        s.opart[(m0 + g + 8) * 4 + t] = c[2];
      }
    }
    __syncthreads();
    if (mt < ntile && kh == 0) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int d = m0 + g + 8 * q;
        if (d < A.dout) s.pred[t * A.dout + d] = (c[2 * q] + s.opart[d * 4 + t]) + s.b3[d];
      }
    }
  }
  __syncthreads();

  if (REV) {
    if (kh == 0) {   // seeds: adjoints of (a3, a3') are c1 * u3[n], c2 * u3[n]
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n = m0 + g + 8 * q;
        const float u = s.u3[n];
        float P, Q, Rr;
        adjoint_unit<2>(saved[2][2 * q], saved[2][2 * q + 1], 0.0f, st.c[0] * u, st.c[1] * u, 0.0f, P, Q, Rr);
        *reinterpret_cast<float2*>(aout + n * 8 + 2 * t) = make_float2(P, Q);
      }
    }
    __syncthreads();
    { float* tmp = ain; ain = aout; aout = tmp; }
#pragma unroll
    for (int l = 1; l >= 0; --l) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma_pass<true>(l == 1 ? s.W2 : s.W1, ain, m0, kh * 8, 8, g, t, c);   // adjoint of a_l[k] = sum_j W[j][k] P_{l+1}[j]
      if (kh == 1) {
        *reinterpret_cast<float2*>(s.part + (m0 + g) * 8 + 2 * t) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(s.part + (m0 + g + 8) * 8 + 2 * t) = make_float2(c[2], c[3]);
      }
      __syncthreads();
      if (kh == 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int n = m0 + g + 8 * q;
          const float2 o = *reinterpret_cast<const float2*>(s.part + n * 8 + 2 * t);
          float P, Q, Rr;
          adjoint_unit<2>(saved[l][2 * q], saved[l][2 * q + 1], 0.0f, c[2 * q] + o.x, c[2 * q + 1] + o.y, 0.0f, P, Q, Rr);
          *reinterpret_cast<float2*>(aout + n * 8 + 2 * t) = make_float2(P, l == 1 ? Q : 0.0f);
        }
      }
      __syncthreads();
      float* tmp = ain; ain = aout; aout = tmp;
    }
    // gradient wrt x: g[p][d] = sum_n W0^T[d][n] P1[n][p]   (rows d of the W0^T image; jet-0 columns)
    {
      const int ntile = (D + 15) >> 4;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      if (mt < ntile) {
        mma_pass<false>(s.wt0, ain, m0, kh * 8, 8, g, t, c);
        if (kh == 1) {
          s.opart[(m0 + g) * 4 + t] = c[0];
          s.opart[(m0 + g + 8) * 4 + t] = c[2];
        }
      }
      __syncthreads();
      if (mt < ntile && kh == 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int d = m0 + g + 8 * q;
          if (d < D) s.gout[t * D + d] = c[2 * q] + s.opart[d * 4 + t];
        }
      }
    }
    __syncthreads();
  }

  if (tid < PB * D) {   // combine, as in eval()
    const int p = tid / D, d = tid - p * D;
    const float diff = __fsub_rn(s.pred[p * A.dout + d], s.xs[tid]);
    float f;
    if (st.kind == 0) f = diff;
    else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
    else {
      f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
      if (REV) f = __fadd_rn(f, s.gout[tid]);
    }
    s.fs[tid] = f;
  }
  __syncthreads();
}

template <int NJ, bool MMA>
__global__ void __launch_bounds__(NT, 1) k_md_chain(NetView net, const __grid_constant__ Args A) {
  extern __shared__ __align__(16) float smem[];
  constexpr int ld = MMA ? LDM : LD;
  const int rows0 = MMA ? ((A.din_pad + 15) & ~15) : A.din_pad, rows3 = MMA ? ((A.dout + 15) & ~15) : A.dout;
  const Smem s = carve(smem, rows0, rows3, ld);
  const int tid = threadIdx.x, D = A.D;

  // the network, once (rows beyond din_pad / dout of the padded images are zero)
  for (int i = tid; i < W * W; i += NT) {
    const int r = i >> 7, c = i & 127;
    s.W1[r * ld + c] = __ldg(net.w1 + i);
    s.W2[r * ld + c] = __ldg(net.w2 + i);
  }
  for (int i = tid; i < rows0 * W; i += NT) s.wt0[(i >> 7) * ld + (i & 127)] = i < A.din_pad * W ? __ldg(net.wt0 + i) : 0.0f;
  for (int i = tid; i < rows3 * W; i += NT) s.w3[(i >> 7) * ld + (i & 127)] = i < A.dout * W ? __ldg(net.w3 + i) : 0.0f;
  if (tid < W) { s.b1[tid] = __ldg(net.b1 + tid); s.b2[tid] = __ldg(net.b2 + tid); s.u3[tid] = __ldg(net.u3 + tid); }
  if (tid < A.dout) s.b3[tid] = __ldg(net.b3 + tid);

  const int nstage = A.method == 1 ? 1 : 3;
  const long tps = (A.Kset + PB - 1) / PB;   // tiles per parameter set: a tile never mixes sets
  const long ntiles = tps * A.n_sets;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {   // a CTA runs the whole chain of one group of 4, then the next
  const long set = tile / tps, in_set = (tile % tps) * PB;
  const long base = set * A.Kset + in_set;   // first row of the tile; row base + r is live iff in_set + r < Kset
  const Stage* stages = A.stages + (size_t)set * A.n_frames * A.n_steps * 3;
  __syncthreads();
  for (int o = tid; o < PB * D; o += NT) {
    const long row = base + o / D;
    s.y[o] = in_set + o / D < A.Kset ? A.x_init[row * D + (o % D)] : 0.0f;
  }
  for (int i = tid; i < A.n_steps * 4; i += NT) s.gv[i] = A.gval_sets ? __ldg(A.gval_sets + (size_t)set * kMaxSdeSteps * 4 + i) : A.gval[i >> 2][i & 3];
  __syncthreads();
  for (int f = 0; f < A.n_frames; ++f) {
    // frame constants: folded layer-0 bias b0 + W0[:, D+1:] . emb[f], stage scalars with gamma2 / (1 + f)
    if (tid < W) s.b0[tid] = __ldg(net.b0 + (size_t)f * W + tid);
    {
      const int* src = reinterpret_cast<const int*>(stages + (size_t)f * A.n_steps * 3);
      int* dst = reinterpret_cast<int*>(s.st);
      for (int i = tid; i < A.n_steps * 3 * (int)(sizeof(Stage) / 4); i += NT) dst[i] = __ldg(src + i);
    }
    for (int o = tid; o < PB * D; o += NT) s.xs[o] = s.y[o];
    __syncthreads();
    for (int stp = 0; stp < A.n_steps; ++stp) {
      const float h = A.h[stp];
      const size_t noff = ((size_t)f * A.n_steps + stp) * A.K * D;
      const float rdt = __fdiv_rn(1.0f, h), sq = __fsqrt_rn(h);
#pragma unroll 1
      for (int sg = 0; sg < nstage; ++sg) {
        const Stage& st = s.st[stp * 3 + sg];
        if (MMA) {
          if (st.kind == 2 && A.reverse) eval_mma<true>(s, A, st, NJ >= 2);
          else eval_mma<false>(s, A, st, false);
        } else {
          if (st.kind == 2 && A.reverse) eval<NJ, true>(s, A, st);
          else eval<1, false>(s, A, st);
        }
        // torchsde srk (Roessler SRI2, diagonal noise, state-independent g) / euler: same update as idff_generic_sde.cu
        for (int o = tid; o < PB * D; o += NT) {
          const long row = base + o / D;
          const bool live = in_set + o / D < A.Kset;
          const size_t gi = noff + (size_t)row * D + (o % D);
          const float y = s.y[o], fv = s.fs[o];
          if (A.method == 1) {
            const float ik = live ? A.Ik[gi] : 0.0f;
            const float yn = __fadd_rn(__fadd_rn(y, __fmul_rn(fv, h)), __fmul_rn(s.gv[stp * 4 + 0], ik));
            s.yprev[o] = y; s.y[o] = yn; s.xs[o] = yn;
          } else if (sg == 0) {
            const float ik0 = live ? A.Ik0[gi] : 0.0f;
            s.k1[o] = fv;
            const float t1 = __fmul_rn(__fmul_rn(1.0f, fv), h);
            const float t2 = __fmul_rn(__fmul_rn(__fmul_rn(0.0f, s.gv[stp * 4 + 0]), ik0), rdt);
            s.xs[o] = __fadd_rn(__fadd_rn(y, t1), t2);
          } else if (sg == 1) {
            const float ik0 = live ? A.Ik0[gi] : 0.0f;
            s.k2[o] = fv;
            float v = y;
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, s.k1[o]), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(1.0f, s.gv[stp * 4 + 0]), ik0), rdt));
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, fv), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(0.5f, s.gv[stp * 4 + 1]), ik0), rdt));
            s.xs[o] = v;
          } else {
            const float ik = live ? A.Ik[gi] : 0.0f;
            const float ik0 = live ? A.Ik0[gi] : 0.0f;
            const float ikk = __fmul_rn(__fsub_rn(__fmul_rn(ik, ik), h), 0.5f);
            const float ikkk = __fdiv_rn(__fsub_rn(__fmul_rn(__fmul_rn(ik, ik), ik), __fmul_rn(__fmul_rn(3.0f, h), ik)), 6.0f);
            const float b1[4] = {-1.0f, (float)(4.0 / 3.0), (float)(2.0 / 3.0), 0.0f};
            const float b2[4] = {1.0f, (float)(-4.0 / 3.0), (float)(1.0 / 3.0), 0.0f};
            const float b3[4] = {2.0f, (float)(-4.0 / 3.0), (float)(-2.0 / 3.0), 0.0f};
            const float b4[4] = {-2.0f, (float)(5.0 / 3.0), (float)(-2.0 / 3.0), 1.0f};
            const float al[4] = {(float)(1.0 / 6.0), (float)(1.0 / 6.0), (float)(2.0 / 3.0), 0.0f};
            const float F[4] = {s.k1[o], s.k2[o], fv, 0.0f};
            float v = y;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float gw = __fmul_rn(b1[q], ik);
              gw = __fadd_rn(gw, __fdiv_rn(__fmul_rn(b2[q], ikk), sq));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b3[q], ik0), rdt));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b4[q], ikkk), rdt));
              v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(al[q], F[q]), h)), __fmul_rn(s.gv[stp * 4 + q], gw));
            }
            s.yprev[o] = y; s.y[o] = v; s.xs[o] = v;
          }
        }
        __syncthreads();
      }
    }
    // traj[-1] of this frame (torchsde's linear interpolation at ts[-1]) = y_hat[f] and the next frame's start (:207)
    for (int o = tid; o < PB * D; o += NT) {
      const long row = base + o / D;
      const float v = __fadd_rn(__fmul_rn(A.w0, s.yprev[o]), __fmul_rn(A.w1, s.y[o]));
      s.y[o] = v;
      if (in_set + o / D < A.Kset) A.y_hat[((size_t)f * A.K + row) * D + (o % D)] = v;
    }
    __syncthreads();
  }
  }
}

}  // namespace chain

static thread_local int g_chain_fma = 0;   // diagnostic: 1 = the FMA evaluator instead of the mma.sync one (idff_debug_md_chain_fma)
void md_chain_set_fma(int on) { g_chain_fma = on ? 1 : 0; }

bool md_chain_supports(const idff_weights_t* w, const idff_field_params_t* p, int64_t K) {
  const NetView& v = w->v;
  return p->variant == IDFF_FIELD_MD && v.w == chain::W && v.din == p->dim + 1 && v.din_pad <= chain::DMAX && v.dout <= chain::DMAX &&
         v.dout >= p->dim && p->dim * chain::PB <= 256 && v.dout * chain::PB <= 256 && v.din_pad % chain::KS == 0 && K >= 1;
}

// n_sets parameter sets (1: the plain call): `stages` holds [n_sets][n_frames][n_steps][3], `gvals` (n_sets > 1) [n_sets][kMaxSdeSteps][4];
// K = trajectories per set, rows set-major in x_init / the increments / y_hat; nj / reverse: the jets the most demanding set needs
int md_chain_run(const idff_weights_t* w, const idff_field_params_t* p, int n_frames, const SdeTable& tab0, const Stage* stages,
                 const float* d_x_init, const float* d_I_k, const float* d_I_k0, float* d_y_hat, int64_t K, void* stream,
                 int n_sets, const float* gvals, int nj_sets, int reverse_sets) {
  using namespace chain;
  cudaStream_t s = (cudaStream_t)stream;
  idff_weights_t* wm = const_cast<idff_weights_t*>(w);
  const size_t st_bytes = (size_t)n_sets * n_frames * tab0.n_steps * 3 * sizeof(Stage);
  const size_t gv_bytes = gvals ? (size_t)n_sets * kMaxSdeSteps * 4 * sizeof(float) : 0;
  const size_t bytes = st_bytes + gv_bytes;
  if (wm->sweep_bytes < bytes) {
    if (wm->d_sweep) cudaFree(wm->d_sweep);
    wm->d_sweep = nullptr;
    wm->sweep_bytes = 0;
    IDFF_CUDA(cudaMalloc(&wm->d_sweep, bytes));
    wm->sweep_bytes = bytes;
  }
  IDFF_CUDA(cudaMemcpyAsync(wm->d_sweep, stages, st_bytes, cudaMemcpyHostToDevice, s));
  if (gvals) IDFF_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(wm->d_sweep) + st_bytes, gvals, gv_bytes, cudaMemcpyHostToDevice, s));
  IDFF_CUDA(cudaStreamSynchronize(s));   // `stages` / `gvals` are pageable host memory owned by the caller
  static thread_local Args A;
  memset(&A, 0, sizeof(A));
  int nj;
  jets_needed(*p, &nj, &A.reverse);
  if (n_sets > 1) { nj = nj_sets; A.reverse = reverse_sets; }
  A.n_frames = n_frames; A.n_steps = tab0.n_steps; A.method = tab0.method;
  A.D = p->dim; A.din = w->v.din; A.din_pad = w->v.din_pad; A.dout = w->v.dout;
  for (int i = 0; i < tab0.n_steps; ++i) {
    A.h[i] = tab0.h[i];
    for (int q = 0; q < 4; ++q) A.gval[i][q] = tab0.gval[i][q];
  }
  A.w0 = tab0.n_emit > 0 ? tab0.emit_w0[tab0.n_emit - 1] : 0.0f;
  A.w1 = tab0.n_emit > 0 ? tab0.emit_w1[tab0.n_emit - 1] : 1.0f;
  A.stages = reinterpret_cast<const Stage*>(wm->d_sweep);
  A.x_init = d_x_init; A.Ik = d_I_k; A.Ik0 = d_I_k0; A.y_hat = d_y_hat; A.K = K * n_sets;
  A.n_sets = n_sets; A.Kset = K;
  A.gval_sets = gvals ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(wm->d_sweep) + st_bytes) : nullptr;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long tiles = (K + PB - 1) / PB * n_sets;
  const int grid = (int)(tiles < sms ? tiles : sms);
  const bool mma = g_chain_fma == 0;
  const int rows0 = mma ? ((A.din_pad + 15) & ~15) : A.din_pad, rows3 = mma ? ((A.dout + 15) & ~15) : A.dout;
  const size_t sm = smem_bytes(rows0, rows3, mma ? LDM : LD);
  if (sm > 232448) {
    set_error("idff_md_generate: %zu B of shared memory needed (limit 232448)", sm);
    return IDFF_EUNSUPPORTED;
  }
  auto launch = [&](auto kern) -> int {
    IDFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<grid, NT, sm, s>>>(w->v, A);
    return IDFF_OK;
  };
  int rc;
  if (mma) rc = nj >= 2 ? launch(k_md_chain<2, true>) : launch(k_md_chain<1, true>);
  else rc = nj >= 2 ? launch(k_md_chain<2, false>) : launch(k_md_chain<1, false>);
  if (rc) return rc;
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}

}  // namespace idff
