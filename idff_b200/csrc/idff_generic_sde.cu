// Generic fused SDE sampler: torchsde srk (Roessler SRI2, diagonal noise) / euler with injected Brownian
// increments, for SDE_cal (MD_simulation.py:481-529).  g(t, y) is state independent, so only the drift
// stages H0 are formed; alpha[3] == 0 makes the 4th drift evaluation irrelevant to the result.
#include "idff_field_generic.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);

template <int NJ, int PB, bool MULTI>
__global__ void __launch_bounds__(256) k_sample_sde(NetView net, FieldCfg cfg, const __grid_constant__ SdeTable tab,
                                                    const float* __restrict__ y0, const float* __restrict__ Ik,
                                                    const float* __restrict__ Ik0, float* __restrict__ out, long N) {
  extern __shared__ __align__(16) float smem[];
  SmemPlan s = carve<NJ, PB>(smem, net, cfg.D);
  const int D = cfg.D, tid = threadIdx.x, nth = blockDim.x;
  float* yprev = s.k3;   // previous step's state (for the interpolated outputs)
  for (long g = blockIdx.x; g * PB < N; g += gridDim.x) {
    const long base = g * PB;
    for (int o = tid; o < PB * D; o += nth) {
      const long row = base + o / D;
      const float v = row < N ? y0[row * D + (o % D)] : 0.0f;
      s.y[o] = v;
      s.xs[o] = v;
      if (row < N) out[row * D + (o % D)] = v;   // ys[0] = y0
    }
    __syncthreads();
    int emit = 0;
    while (emit < tab.n_emit && tab.emit_step[emit] < 0) ++emit;
    for (int stp = 0; stp < tab.n_steps; ++stp) {
      const float h = tab.h[stp];
      const size_t noff = (size_t)stp * N * D;
      const float rdt = __fdiv_rn(1.0f, h), sq = __fsqrt_rn(h);
      const int nstage = tab.method == 1 ? 1 : 3;
#pragma unroll 1
      for (int sg = 0; sg < nstage; ++sg) {
        float* kout = sg == 0 ? s.k1 : (sg == 1 ? s.k2 : s.fs);
        eval_dispatch<NJ, PB, MULTI>(net, cfg, tab.st[stp][sg], s.xs, s.x0s, kout, s.actA, s.actB, s.pred, s.gout);
        for (int o = tid; o < PB * D; o += nth) {
          const long row = base + o / D;
          const size_t gi = noff + (size_t)row * D + (o % D);
          const float y = s.y[o];
          if (tab.method == 1) {
            // euler: y + f(t0,y) h + g(t0) I_k
            const float ik = row < N ? Ik[gi] : 0.0f;
            const float yn = __fadd_rn(__fadd_rn(y, __fmul_rn(s.k1[o], h)), __fmul_rn(tab.gval[stp][0], ik));
            yprev[o] = y; s.y[o] = yn; s.xs[o] = yn;
          } else if (sg == 0) {
            // H0_1 = y + A0[1][0] F0 h + B0[1][0] G0 Ik0/h,  A0 = 1, B0 = 0
            const float ik0 = row < N ? Ik0[gi] : 0.0f;
            const float t1 = __fmul_rn(__fmul_rn(1.0f, s.k1[o]), h);
            const float t2 = __fmul_rn(__fmul_rn(__fmul_rn(0.0f, tab.gval[stp][0]), ik0), rdt);
            s.xs[o] = __fadd_rn(__fadd_rn(y, t1), t2);
          } else if (sg == 1) {
            // H0_2 = y + (1/4 F0 h + 1 G0 Ik0/h) + (1/4 F1 h + 1/2 G1 Ik0/h)
            const float ik0 = row < N ? Ik0[gi] : 0.0f;
            float v = y;
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, s.k1[o]), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(1.0f, tab.gval[stp][0]), ik0), rdt));
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, s.k2[o]), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(0.5f, tab.gval[stp][1]), ik0), rdt));
            s.xs[o] = v;
          } else {
            // y1 = y + sum_s alpha_s F_s h + G_s * g_weight_s
            const float ik = row < N ? Ik[gi] : 0.0f;
            const float ik0 = row < N ? Ik0[gi] : 0.0f;
            const float ikk = __fmul_rn(__fsub_rn(__fmul_rn(ik, ik), h), 0.5f);
            const float ikkk = __fdiv_rn(__fsub_rn(__fmul_rn(__fmul_rn(ik, ik), ik), __fmul_rn(__fmul_rn(3.0f, h), ik)), 6.0f);
            const float b1[4] = {-1.0f, (float)(4.0 / 3.0), (float)(2.0 / 3.0), 0.0f};
            const float b2[4] = {1.0f, (float)(-4.0 / 3.0), (float)(1.0 / 3.0), 0.0f};
            const float b3[4] = {2.0f, (float)(-4.0 / 3.0), (float)(-2.0 / 3.0), 0.0f};
            const float b4[4] = {-2.0f, (float)(5.0 / 3.0), (float)(-2.0 / 3.0), 1.0f};
            const float al[4] = {(float)(1.0 / 6.0), (float)(1.0 / 6.0), (float)(2.0 / 3.0), 0.0f};
            const float F[4] = {s.k1[o], s.k2[o], s.fs[o], 0.0f};
            float v = y;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float gw = __fmul_rn(b1[q], ik);
              gw = __fadd_rn(gw, __fdiv_rn(__fmul_rn(b2[q], ikk), sq));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b3[q], ik0), rdt));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b4[q], ikkk), rdt));
              v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(al[q], F[q]), h)), __fmul_rn(tab.gval[stp][q], gw));
            }
            yprev[o] = y; s.y[o] = v; s.xs[o] = v;
          }
        }
        __syncthreads();
      }
      while (emit < tab.n_emit && tab.emit_step[emit] == stp) {
        const float w0 = tab.emit_w0[emit], w1 = tab.emit_w1[emit];
        const size_t ooff = (size_t)tab.emit_out[emit] * N * D;
        for (int o = tid; o < PB * D; o += nth) {
          const long row = base + o / D;
          if (row < N) out[ooff + row * D + (o % D)] = __fadd_rn(__fmul_rn(w0, yprev[o]), __fmul_rn(w1, s.y[o]));
        }
        ++emit;
      }
      __syncthreads();
    }
  }
}

int generic_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, const float* d_y0,
                       const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream) {
  FieldCfg cfg;
  int nj;
  cfg.variant = p->variant; cfg.D = p->dim; cfg.tidx = p->t_idx;
  jets_needed(*p, &nj, &cfg.reverse);
  GenPlan pl;
  int rc = plan_generic(w, p->dim, nj, N, &pl);
  if (rc) return rc;
  IDFF_GEN_DISPATCH(k_sample_sde, pl, stream, w->v, cfg, tab, d_y0, d_Ik, d_Ik0, d_out, (long)N);
}

}  // namespace idff
