// Generic fused RK4 (3/8 rule) sampler: torchdiffeq fixed-grid rk4 == rk_common.rk4_alt_step_func, whole solve
// in one launch, state of the CTA's particles in shared memory.
#include "idff_field_generic.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);

template <int NJ, int PB, bool MULTI>
__global__ void __launch_bounds__(256) k_sample_rk4(NetView net, FieldCfg cfg, const __grid_constant__ Rk4Table tab,
                                                    const float* __restrict__ x0, const float* __restrict__ x0_cfm,
                                                    float* __restrict__ out_final, float* __restrict__ traj, long N) {
  extern __shared__ __align__(16) float smem[];
  SmemPlan s = carve<NJ, PB>(smem, net, cfg.D);
  const int D = cfg.D, tid = threadIdx.x, nth = blockDim.x;
  const float third = (float)(1.0 / 3.0);
  for (long g = blockIdx.x; g * PB < N; g += gridDim.x) {
    const long base = g * PB;
    for (int o = tid; o < PB * D; o += nth) {
      const long row = base + o / D;
      const float v = row < N ? x0[row * D + (o % D)] : 0.0f;
      s.y[o] = v;
      s.xs[o] = v;
      if (x0_cfm != nullptr) s.x0s[o] = row < N ? x0_cfm[row * D + (o % D)] : 0.0f;   // later chunk of a CFM solve
    }
    __syncthreads();
    for (int iv = 0; iv < tab.n_iv; ++iv) {
      const float dt = tab.dt[iv];
#pragma unroll 1
      for (int sg = 0; sg < 4; ++sg) {
        float* kout = sg == 0 ? s.k1 : (sg == 1 ? s.k2 : (sg == 2 ? s.k3 : s.fs));
        eval_dispatch<NJ, PB, MULTI>(net, cfg, tab.st[4 * iv + sg], s.xs, s.x0s, kout, s.actA, s.actB, s.pred, s.gout);
        for (int o = tid; o < PB * D; o += nth) {
          const float y = s.y[o], k1 = s.k1[o];
          float xn;
          if (sg == 0) {          // y + dt*k1*(1/3)
            xn = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, k1), third));
          } else if (sg == 1) {   // y + dt*(k2 - k1*(1/3))
            xn = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(s.k2[o], __fmul_rn(k1, third))));
          } else if (sg == 2) {   // y + dt*(k1 - k2 + k3)
            xn = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1, s.k2[o]), s.k3[o])));
          } else {                // y + (k1 + 3*(k2+k3) + k4)*dt*0.125
            const float sum = __fadd_rn(__fadd_rn(k1, __fmul_rn(3.0f, __fadd_rn(s.k2[o], s.k3[o]))), s.fs[o]);
            xn = __fadd_rn(y, __fmul_rn(__fmul_rn(sum, dt), 0.125f));
            s.y[o] = xn;
            const long row = base + o / D;
            if (traj != nullptr && row < N) traj[((size_t)(iv + 1) * N + row) * D + (o % D)] = xn;
          }
          s.xs[o] = xn;
        }
        __syncthreads();
      }
    }
    for (int o = tid; o < PB * D; o += nth) {
      const long row = base + o / D;
      if (row < N) out_final[row * D + (o % D)] = s.y[o];
    }
    __syncthreads();
  }
}

int generic_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                       const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream) {
  FieldCfg cfg;
  int nj;
  cfg.variant = p->variant; cfg.D = p->dim; cfg.tidx = p->t_idx;
  jets_needed(*p, &nj, &cfg.reverse);
  GenPlan pl;
  int rc = plan_generic(w, p->dim, nj, N, &pl);
  if (rc) return rc;
  static thread_local Rk4Table tab;
  const float* src = d_x0;
  for (int done = 0; done < n_iv; done += kMaxIntervals) {
    const int cnt = n_iv - done < kMaxIntervals ? n_iv - done : kMaxIntervals;
    tab.n_iv = cnt;
    memcpy(tab.dt, dts + done, sizeof(float) * cnt);
    memcpy(tab.st, stages + 4 * (size_t)done, sizeof(Stage) * 4 * cnt);
    float* traj = d_out_traj ? d_out_traj + (size_t)done * N * p->dim : nullptr;
    const float* x0c = (done > 0 && p->variant == IDFF_FIELD_CFM) ? d_x0 : nullptr;
    auto one = [&]() -> int { IDFF_GEN_DISPATCH(k_sample_rk4, pl, stream, w->v, cfg, tab, src, x0c, d_out_final, traj, (long)N); };
    rc = one();
    if (rc) return rc;
    src = d_out_final;
  }
  return IDFF_OK;
}

}  // namespace idff
