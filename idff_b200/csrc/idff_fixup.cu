// Range safety net of the fp16-split tensor-core routes (k_tc16, the layered 3xFP16 engine).
//
// Those kernels carry an fp32 value as two fp16 numbers; an activation, jet or adjoint above 65504 turns into inf/NaN in
// that particle's own columns (columns never mix), and the particle's result -- field value or end point of the solve --
// comes out non-finite where the reference's fp32 path (Autograd_example_order2.py:44-76) is finite.  IDFF_IMPL_AUTO must
// never return a value the reference would not, so every AUTO call on such a route is followed, on the same stream and
// without a host round trip, by k_fixup: it scans the output rows and re-solves exactly the non-finite particles from their
// inputs in fp32 FMA arithmetic (the generic evaluator of idff_field_generic.cuh: same formulation, no fp16 anywhere).
// A particle the reference itself drives to inf/NaN stays non-finite.  On the clean path the cost is one read of the output
// (8 B per particle at D = 2: ~25 us per 2^24 particles, against a 370 ms solve).
//
// A CTA owns a contiguous slice of rows, collects non-finite rows into a shared-memory queue (never across two parameter
// sets of a sweep) and re-solves them in groups of PB.  Particles are independent, so the queue order does not matter.
#include <vector>

#include "idff_field_generic.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);

struct FixArgs {
  NetView net;
  int variant, D, tidx;
  int mode;      // 0: one field evaluation (stage rk.st[0]); 1: rk4 solve; 2: rk4 sweep, n_sets parameter sets over the same x0
  int n_sets, n_iv, reverse;
  long N;                 // rows per set
  const float* x_in;      // [N][D]
  const float* x0_cfm;    // CFM: the x captured at t == 0 when the launch does not start there (eval at t != 0, later chunk), or null
  float* out;             // [n_sets][N][D]
  float* traj;            // mode 1, optional: rows 1.. of this chunk, [n_iv][N][D]
  const SweepTab* tabs;   // mode 2 (fused tensor kernel): [n_sets] in global memory, or null
  const Stage* g_st;      // global tables (layered engine; null: the grid-constant table): [n_sets][4 n_iv]
  const float* g_dt;      // [n_iv]
  unsigned long long* counters;   // [1] += rows re-solved
};

template <int NJ, int PB, bool MULTI>
__global__ void __launch_bounds__(256) k_fixup(const __grid_constant__ FixArgs A, const __grid_constant__ Rk4Table rk) {
  extern __shared__ __align__(16) float smem[];
  __shared__ long q_rows[256 + PB];
  __shared__ int q_n;
  SmemPlan s = carve<NJ, PB>(smem, A.net, A.D);
  const int D = A.D, tid = threadIdx.x, nth = blockDim.x;
  const float third = (float)(1.0 / 3.0);
  const long total = (long)A.n_sets * A.N;
  const long per = (total + gridDim.x - 1) / gridDim.x;
  const long lo = (long)blockIdx.x * per, hi = lo + per < total ? lo + per : total;
  if (tid == 0) q_n = 0;
  __syncthreads();
  long base = lo;
  long q_set = 0;
  while (true) {
    // ---- refill: scan up to nth rows at a time, never past the end of the current parameter set
    while (true) {
      const int n = q_n;
      __syncthreads();   // every thread has read q_n before anyone appends to the queue
      if (n >= PB || base >= hi) break;
      const long set = base / A.N;
      if (n > 0 && set != q_set) break;
      q_set = set;
      const long end0 = (set + 1) * A.N < hi ? (set + 1) * A.N : hi;
      const long end = base + nth < end0 ? base + nth : end0;
      const long idx = base + tid;
      if (idx < end) {
        bool bad = false;
        for (int d = 0; d < D; ++d) bad = bad || !(fabsf(A.out[idx * D + d]) <= 3.0e38f);
        if (bad) q_rows[atomicAdd(&q_n, 1)] = idx;
      }
      base = end;
      __syncthreads();
    }
    const int n = q_n;
    if (n == 0) break;
    const int cnt = n < PB ? n : PB, q0 = n - cnt;
    const int gs = (int)q_set;
    FieldCfg cfg;
    cfg.variant = A.variant; cfg.D = D; cfg.tidx = A.tidx;
    cfg.reverse = (A.mode == 2 && A.tabs) ? A.tabs[gs].reverse : A.reverse;
    auto stage_of = [&](int e) -> const Stage& {
      return A.g_st ? A.g_st[(size_t)gs * 4 * A.n_iv + e] : (A.mode == 2 ? A.tabs[gs].st[e] : rk.st[e]);
    };
    for (int o = tid; o < PB * D; o += nth) {
      const int p = o / D, d = o - p * D;
      const long row = p < cnt ? q_rows[q0 + p] - (long)gs * A.N : -1;
      const float v = row >= 0 ? A.x_in[row * D + d] : 0.0f;
      s.y[o] = v;
      s.xs[o] = v;
      if (A.x0_cfm != nullptr) s.x0s[o] = row >= 0 ? A.x0_cfm[row * D + d] : 0.0f;
    }
    __syncthreads();
    if (A.mode == 0) {
      eval_dispatch<NJ, PB, MULTI>(A.net, cfg, stage_of(0), s.xs, s.x0s, s.fs, s.actA, s.actB, s.pred, s.gout);
      for (int o = tid; o < PB * D; o += nth) {
        const int p = o / D, d = o - p * D;
        if (p < cnt) A.out[q_rows[q0 + p] * D + d] = s.fs[o];
      }
    } else {
      for (int iv = 0; iv < A.n_iv; ++iv) {
        const float dt = A.g_dt ? A.g_dt[iv] : (A.mode == 2 ? A.tabs[gs].dt[iv] : rk.dt[iv]);
#pragma unroll 1
        for (int sg = 0; sg < 4; ++sg) {
          float* kout = sg == 0 ? s.k1 : (sg == 1 ? s.k2 : (sg == 2 ? s.k3 : s.fs));
          eval_dispatch<NJ, PB, MULTI>(A.net, cfg, stage_of(4 * iv + sg), s.xs, s.x0s, kout, s.actA, s.actB, s.pred, s.gout);
          for (int o = tid; o < PB * D; o += nth) {   // torchdiffeq rk4_alt_step_func, as in k_sample_rk4
            const float y = s.y[o], k1 = s.k1[o];
            float xn;
            if (sg == 0) {
              xn = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, k1), third));
            } else if (sg == 1) {
              xn = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(s.k2[o], __fmul_rn(k1, third))));
            } else if (sg == 2) {
              xn = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1, s.k2[o]), s.k3[o])));
            } else {
              const float sum = __fadd_rn(__fadd_rn(k1, __fmul_rn(3.0f, __fadd_rn(s.k2[o], s.k3[o]))), s.fs[o]);
              xn = __fadd_rn(y, __fmul_rn(__fmul_rn(sum, dt), 0.125f));
              s.y[o] = xn;
              const int p = o / D, d = o - p * D;
              if (A.traj != nullptr && p < cnt) A.traj[((size_t)iv * A.N + (q_rows[q0 + p] - (long)gs * A.N)) * D + d] = xn;
            }
            s.xs[o] = xn;
          }
          __syncthreads();
        }
      }
      for (int o = tid; o < PB * D; o += nth) {
        const int p = o / D, d = o - p * D;
        if (p < cnt) A.out[q_rows[q0 + p] * D + d] = s.y[o];
      }
    }
    __syncthreads();
    if (tid == 0) {
      q_n = q0;
      if (A.counters) atomicAdd(A.counters + 1, (unsigned long long)cnt);
    }
    __syncthreads();
  }
}

// Scan `out` and re-solve its non-finite rows in fp32.  p: the parameter set (set 0 of a sweep: variant / dim / t_idx are
// common to the sets).  For mode 0, rk->st[0] is the evaluation's stage.  d_tabs: device SweepTab array (mode 2) or null;
// g_st / g_dt: device tables for solves the grid-constant table cannot hold, or null.
int fixup_nonfinite(const idff_weights_t* w, const idff_field_params_t* p, int mode, int n_sets, int n_iv, const Rk4Table* rk,
                    const SweepTab* d_tabs, const Stage* g_st, const float* g_dt, const float* d_x_in, const float* d_x0_cfm,
                    float* d_out, float* d_traj, int64_t N, void* stream) {
  if (N <= 0 || n_sets <= 0) return IDFF_OK;
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  if (mode == 2) nj = nj < 2 ? 2 : nj;   // the sets of a sweep differ in jet count: unused jets carry zero coefficients
  const int njk = nj <= 2 ? 2 : 3;
  GenPlan pl;
  int rc = plan_generic(w, p->dim, njk, N, &pl);
  if (rc) return rc;
  FixArgs A;
  memset(&A, 0, sizeof(A));
  A.net = w->v;
  A.variant = p->variant; A.D = p->dim; A.tidx = p->t_idx;
  A.mode = mode; A.n_sets = n_sets; A.n_iv = n_iv; A.reverse = rev;
  A.N = (long)N;
  A.x_in = d_x_in; A.x0_cfm = (p->variant == IDFF_FIELD_CFM) ? d_x0_cfm : nullptr;
  A.out = d_out; A.traj = d_traj;
  A.tabs = d_tabs; A.g_st = g_st; A.g_dt = g_dt;
  A.counters = w->d_counters;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long total = (long)n_sets * (long)N;
  const long want = (total + 255) / 256;
  const int grid = (int)(want < 2L * sms ? want : 2L * sms);
  static thread_local Rk4Table rk0;
  const Rk4Table& rkr = rk ? *rk : rk0;
  auto launch = [&](auto kern) -> int {
    IDFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    kern<<<grid, pl.threads, pl.smem, (cudaStream_t)stream>>>(A, rkr);
    IDFF_CUDA(cudaGetLastError());
    count_launch();
    return IDFF_OK;
  };
  if (njk == 2) return pl.multi ? launch(k_fixup<2, 4, true>) : launch(k_fixup<2, 8, false>);
  return pl.multi ? launch(k_fixup<3, 4, true>) : launch(k_fixup<3, 8, false>);
}

// The same safety net behind the PolyALA tensor kernel (k_md16): trajectories whose last output row is non-finite are
// integrated again from y0 with the fp32 generic evaluator under torchsde's srk / euler (the update of idff_generic_sde.cu,
// rows taken from the queue), every requested output row rewritten.
template <int NJ, int PB, bool MULTI>
__global__ void __launch_bounds__(256) k_fixup_sde(NetView net, FieldCfg cfg, const __grid_constant__ SdeTable tab,
                                                   const float* __restrict__ y0, const float* __restrict__ Ik,
                                                   const float* __restrict__ Ik0, float* __restrict__ out, long N, int n_ts,
                                                   unsigned long long* counters) {
  extern __shared__ __align__(16) float smem[];
  __shared__ long q_rows[256 + PB];
  __shared__ int q_n;
  SmemPlan s = carve<NJ, PB>(smem, net, cfg.D);
  const int D = cfg.D, tid = threadIdx.x, nth = blockDim.x;
  float* yprev = s.k3;
  const float* last = out + (size_t)(n_ts - 1) * N * D;
  const long per = (N + gridDim.x - 1) / gridDim.x;
  const long lo = (long)blockIdx.x * per, hi = lo + per < N ? lo + per : N;
  if (tid == 0) q_n = 0;
  __syncthreads();
  long base = lo;
  while (true) {
    while (true) {
      const int n = q_n;
      __syncthreads();
      if (n >= PB || base >= hi) break;
      const long end = base + nth < hi ? base + nth : hi;
      const long idx = base + tid;
      if (idx < end) {
        bool bad = false;
        for (int d = 0; d < D; ++d) bad = bad || !(fabsf(last[idx * D + d]) <= 3.0e38f);
        if (bad) q_rows[atomicAdd(&q_n, 1)] = idx;
      }
      base = end;
      __syncthreads();
    }
    const int n = q_n;
    if (n == 0) break;
    const int cnt = n < PB ? n : PB, q0 = n - cnt;
    auto row_of = [&](int p) -> long { return p < cnt ? q_rows[q0 + p] : -1; };
    for (int o = tid; o < PB * D; o += nth) {
      const long row = row_of(o / D);
      const float v = row >= 0 ? y0[row * D + (o % D)] : 0.0f;
      s.y[o] = v;
      s.xs[o] = v;
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
        for (int o = tid; o < PB * D; o += nth) {   // torchsde srk / euler, as in k_sample_sde
          const long row = row_of(o / D);
          const size_t gi = noff + (size_t)(row >= 0 ? row : 0) * D + (o % D);
          const float y = s.y[o];
          if (tab.method == 1) {
            const float ik = row >= 0 ? Ik[gi] : 0.0f;
            const float yn = __fadd_rn(__fadd_rn(y, __fmul_rn(s.k1[o], h)), __fmul_rn(tab.gval[stp][0], ik));
            yprev[o] = y; s.y[o] = yn; s.xs[o] = yn;
          } else if (sg == 0) {
            const float ik0 = row >= 0 ? Ik0[gi] : 0.0f;
            const float t1 = __fmul_rn(__fmul_rn(1.0f, s.k1[o]), h);
            const float t2 = __fmul_rn(__fmul_rn(__fmul_rn(0.0f, tab.gval[stp][0]), ik0), rdt);
            s.xs[o] = __fadd_rn(__fadd_rn(y, t1), t2);
          } else if (sg == 1) {
            const float ik0 = row >= 0 ? Ik0[gi] : 0.0f;
            float v = y;
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, s.k1[o]), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(1.0f, tab.gval[stp][0]), ik0), rdt));
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, s.k2[o]), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(0.5f, tab.gval[stp][1]), ik0), rdt));
            s.xs[o] = v;
          } else {
            const float ik = row >= 0 ? Ik[gi] : 0.0f;
            const float ik0 = row >= 0 ? Ik0[gi] : 0.0f;
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
            for (int qq = 0; qq < 4; ++qq) {
              float gw = __fmul_rn(b1[qq], ik);
              gw = __fadd_rn(gw, __fdiv_rn(__fmul_rn(b2[qq], ikk), sq));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b3[qq], ik0), rdt));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b4[qq], ikkk), rdt));
              v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(al[qq], F[qq]), h)), __fmul_rn(tab.gval[stp][qq], gw));
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
          const long row = row_of(o / D);
          if (row >= 0) out[ooff + row * D + (o % D)] = __fadd_rn(__fmul_rn(w0, yprev[o]), __fmul_rn(w1, s.y[o]));
        }
        ++emit;
      }
      __syncthreads();
    }
    if (tid == 0) {
      q_n = q0;
      if (counters) atomicAdd(counters + 1, (unsigned long long)cnt);
    }
    __syncthreads();
  }
}

int fixup_nonfinite_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, int n_ts, const float* d_y0,
                        const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream) {
  if (N <= 0 || n_ts < 2 || tab.n_steps == 0) return IDFF_OK;
  FieldCfg cfg;
  int nj;
  cfg.variant = p->variant; cfg.D = p->dim; cfg.tidx = p->t_idx;
  jets_needed(*p, &nj, &cfg.reverse);
  const int njk = nj <= 2 ? 2 : 3;
  GenPlan pl;
  int rc = plan_generic(w, p->dim, njk, N, &pl);
  if (rc) return rc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long want = ((long)N + 255) / 256;
  const int grid = (int)(want < 2L * sms ? want : 2L * sms);
  auto launch = [&](auto kern) -> int {
    IDFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    kern<<<grid, pl.threads, pl.smem, (cudaStream_t)stream>>>(w->v, cfg, tab, d_y0, d_Ik, d_Ik0, d_out, (long)N, n_ts, w->d_counters);
    IDFF_CUDA(cudaGetLastError());
    count_launch();
    return IDFF_OK;
  };
  if (njk == 2) return pl.multi ? launch(k_fixup_sde<2, 4, true>) : launch(k_fixup_sde<2, 8, false>);
  return pl.multi ? launch(k_fixup_sde<3, 4, true>) : launch(k_fixup_sde<3, 8, false>);
}

// AUTO is range safe: a launch must leave its input rows intact for the safety net.  Copies them into handle-owned scratch
// when the call runs in place or continues a chunked solve; *src then points at the copy.
int keep_input(idff_weights_t* wm, const float** src, const float* d_out, bool force, size_t bytes, void* stream) {
  if (!force && *src != d_out) return IDFF_OK;
  if (wm->fix_x_bytes < bytes) {
    if (wm->d_fix_x) cudaFree(wm->d_fix_x);
    wm->d_fix_x = nullptr;
    wm->fix_x_bytes = 0;
    IDFF_CUDA(cudaMalloc(&wm->d_fix_x, bytes));
    wm->fix_x_bytes = bytes;
  }
  IDFF_CUDA(cudaMemcpyAsync(wm->d_fix_x, *src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  *src = wm->d_fix_x;
  return IDFF_OK;
}

// The safety net behind the layered engine, which takes a whole solve (any number of intervals) or a whole sweep in one
// call: stage tables from host arrays -- the grid-constant table when it holds them, a handle-owned device copy otherwise.
// mode 0: stages[0] is the evaluation; mode 1: stages [4 n_iv]; mode 2: stages [n_sets][4 n_iv] (one jet count for all sets).
int fixup_with_tables(const idff_weights_t* w, const idff_field_params_t* p, int mode, int n_sets, const Stage* stages,
                      const float* dts, int n_iv, const float* d_x_in, const float* d_x0_cfm, float* d_out, float* d_traj,
                      int64_t N, void* stream) {
  if (mode == 0 || (mode == 1 && n_iv <= kMaxIntervals)) {
    static thread_local Rk4Table tab;
    tab.n_iv = n_iv;
    if (mode == 0) tab.st[0] = stages[0];
    else {
      memcpy(tab.dt, dts, sizeof(float) * n_iv);
      memcpy(tab.st, stages, sizeof(Stage) * 4 * n_iv);
    }
    return fixup_nonfinite(w, p, mode, 1, n_iv, &tab, nullptr, nullptr, nullptr, d_x_in, d_x0_cfm, d_out, d_traj, N, stream);
  }
  idff_weights_t* wm = const_cast<idff_weights_t*>(w);
  const size_t st_bytes = sizeof(Stage) * 4 * (size_t)n_iv * n_sets, bytes = st_bytes + sizeof(float) * n_iv;
  if (wm->fix_tab_bytes < bytes) {
    if (wm->d_fix_tab) cudaFree(wm->d_fix_tab);
    wm->d_fix_tab = nullptr;
    wm->fix_tab_bytes = 0;
    IDFF_CUDA(cudaMalloc(&wm->d_fix_tab, bytes));
    wm->fix_tab_bytes = bytes;
  }
  // pageable sources: the runtime stages these copies before returning
  unsigned char* d = reinterpret_cast<unsigned char*>(wm->d_fix_tab);
  IDFF_CUDA(cudaMemcpyAsync(d, stages, st_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  IDFF_CUDA(cudaMemcpyAsync(d + st_bytes, dts, sizeof(float) * n_iv, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return fixup_nonfinite(w, p, mode, n_sets, n_iv, nullptr, nullptr, reinterpret_cast<const Stage*>(d),
                         reinterpret_cast<const float*>(d + st_bytes), d_x_in, d_x0_cfm, d_out, d_traj, N, stream);
}

}  // namespace idff
