// extern "C" entry points of the sampler side: stage-table construction on the host (fp32, in the operation
// order of torchdiffeq / torchsde / the reference wrappers) and dispatch to the kernel families.
#include <cmath>
#include <vector>

#include "idff_common.cuh"

namespace idff {
// idff_field_generic.cu
int check_field_args(const idff_weights_t* w, const idff_field_params_t* p, const char* who);
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);
int generic_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x,
                       const float* d_x0, float* d_out, int64_t N, void* stream);
int generic_mlp_forward(const idff_weights_t* w, const float* d_in, float* d_out, int64_t N, int t_idx, void* stream);
int generic_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                       const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream);
int generic_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, const float* d_y0,
                       const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream);
// idff_sampler_fused.cu
bool fused_supports(const idff_weights_t* w, const idff_field_params_t* p);
int fused_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                     const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream);
int fused_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x,
                     const float* d_x0, float* d_out, int64_t N, void* stream);
int fused_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, const float* d_y0,
                     const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream);

// idff_sampler_tc.cu
bool tensor_supports(const idff_weights_t* w, const idff_field_params_t* p);
int tensor_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, const float* d_x0,
                      float* d_out, int64_t N, void* stream);
int tensor_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                      const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream);

int tensor_debug_stamps(long long* h_out, int n);

// idff_sampler_tc16.cu
bool tensor16_supports(const idff_weights_t* w, const idff_field_params_t* p);
int tensor16_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, const float* d_x0,
                        float* d_out, int64_t N, void* stream);
int tensor16_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                        const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream);
int tensor16_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, const float* d_x0,
                              const Stage* stages, const float* dts, int n_iv, float* d_out, int64_t N, void* stream);
int tensor16_debug_stamps(long long* h_out, int n);

// idff_sampler_tc16o3.cu: the fused three-jet kernel (order-3 field)
bool tensor16o3_supports(const idff_weights_t* w, const idff_field_params_t* p);
bool tensor16o3_can_run(const idff_weights_t* w, const idff_field_params_t* p);
int tensor16o3_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, float* d_out,
                          int64_t N, bool safe, void* stream);
int tensor16o3_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                          const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, bool safe, void* stream);
int tensor16o3_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, const float* d_x0,
                                const Stage* stages, const float* dts, int n_iv, float* d_out, int64_t N, bool safe, void* stream);

// idff_md_tc16.cu: the PolyALA tensor kernel
bool md16_supports(const idff_weights_t* w, const idff_field_params_t* p);
int md16_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, float* d_out, int64_t N,
                    void* stream);
int md16_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, const float* d_y0,
                    const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream);

// idff_sampler_wide.cu
bool wide_supports(const idff_weights_t* w, const idff_field_params_t* p);
bool wide_sweep_supports(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, int64_t N);
int wide_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int n_sets, const float* d_x0,
                          const Stage* stages, const float* dts, int n_iv, float* d_out, int64_t N, void* stream);
int wide_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x, const float* d_x0,
                    float* d_out, int64_t N, void* stream);
int wide_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                    const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream);

// idff_fixup.cu: the fp32 safety net behind the fp16-split routes of IDFF_IMPL_AUTO
int keep_input(idff_weights_t* wm, const float** src, const float* d_out, bool force, size_t bytes, void* stream);
int fixup_with_tables(const idff_weights_t* w, const idff_field_params_t* p, int mode, int n_sets, const Stage* stages,
                      const float* dts, int n_iv, const float* d_x_in, const float* d_x0_cfm, float* d_out, float* d_traj,
                      int64_t N, void* stream);

// AUTO on the layered 3xFP16 engine: the engine, then the safety net on the same stream
static int safe_wide_field_eval(const idff_weights_t* w, const idff_field_params_t* q, float t, const float* d_x, const float* d_x0,
                                float* d_out, int64_t N, void* stream) {
  int rc = keep_input(const_cast<idff_weights_t*>(w), &d_x, d_out, false, (size_t)N * q->dim * sizeof(float), stream);
  if (rc) return rc;
  rc = wide_field_eval(w, q, t, d_x, d_x0, d_out, N, stream);
  if (rc) return rc;
  Stage st;
  make_stage(*q, t, &st);
  return fixup_with_tables(w, q, 0, 1, &st, nullptr, 0, d_x, d_x0, d_out, nullptr, N, stream);
}
static int safe_wide_sample_rk4(const idff_weights_t* w, const idff_field_params_t* q, const float* d_x0, const Stage* stages,
                                const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream) {
  int rc = keep_input(const_cast<idff_weights_t*>(w), &d_x0, d_out_final, false, (size_t)N * q->dim * sizeof(float), stream);
  if (rc) return rc;
  rc = wide_sample_rk4(w, q, d_x0, stages, dts, n_iv, d_out_final, d_out_traj, N, stream);
  if (rc) return rc;
  return fixup_with_tables(w, q, 1, 1, stages, dts, n_iv, d_x0, nullptr, d_out_final,
                           d_out_traj ? d_out_traj + (size_t)N * q->dim : nullptr, N, stream);
}

int fixup_nonfinite_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, int n_ts, const float* d_y0,
                        const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream);

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// torchsde BaseSDESolver.integrate: curr_t += dt (fp32 tensor + Python float), clipped to ts[-1].
static std::vector<float> sde_grid(const float* ts, int n_ts, float dt) {
  std::vector<float> g;
  g.push_back(ts[0]);
  const float end = ts[n_ts - 1];
  while (g.back() < end && g.size() < 100000) {
    volatile float nxt = g.back() + dt;
    g.push_back(nxt < end ? nxt : end);
  }
  return g;
}
}  // namespace idff

using namespace idff;

static thread_local int64_t g_small_n = 0;   // solves of at most this many particles take the 32-particle-tile route (0: off)
extern "C" int idff_tune_small_n_route(int64_t n_max) {
  g_small_n = n_max;
  return IDFF_OK;
}

extern "C" int idff_debug_tensor_stamps(int64_t* h_out32) {
  IDFF_CHECK_ARG(h_out32, "idff_debug_tensor_stamps: null pointer");
  return tensor_debug_stamps(reinterpret_cast<long long*>(h_out32), 32);
}

extern "C" int idff_debug_tensor16_stamps(int64_t* h_out32) {
  IDFF_CHECK_ARG(h_out32, "idff_debug_tensor16_stamps: null pointer");
  return tensor16_debug_stamps(reinterpret_cast<long long*>(h_out32), 32);
}

extern "C" int idff_weights_nonfinite(const idff_weights_t* w, uint64_t* h_out2) {
  IDFF_CHECK_ARG(w && h_out2, "idff_weights_nonfinite: null pointer");
  DeviceGuard g(w->device);
  unsigned long long c[2] = {0, 0};
  IDFF_CUDA(cudaMemcpy(c, w->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
  h_out2[0] = (uint64_t)c[0];
  h_out2[1] = (uint64_t)c[1];
  return IDFF_OK;
}

extern "C" int idff_stage_scalars(const idff_field_params_t* p, float t, float* h_out8) {
  IDFF_CHECK_ARG(p && h_out8, "idff_stage_scalars: null pointer");
  Stage s;
  make_stage(*p, t, &s);
  h_out8[0] = s.t; h_out8[1] = (float)s.kind; h_out8[2] = s.a; h_out8[3] = s.den;
  h_out8[4] = s.c[0]; h_out8[5] = s.c[1]; h_out8[6] = s.c[2]; h_out8[7] = s.end_div;
  return IDFF_OK;
}

extern "C" int idff_mlp_forward(const idff_weights_t* w, const float* d_in, float* d_out, int64_t N, int32_t t_idx,
                                void* stream) {
  IDFF_CHECK_ARG(w && d_in && d_out, "idff_mlp_forward: null pointer");
  IDFF_CHECK_ARG(t_idx >= 0 && t_idx < w->v.n_tidx, "idff_mlp_forward: t_idx %d outside the embedding table (%d rows)", t_idx, w->v.n_tidx);
  if (N == 0) return IDFF_OK;
  IDFF_CHECK_ARG(N > 0, "idff_mlp_forward: negative N");
  DeviceGuard g(w->device);
  return generic_mlp_forward(w, d_in, d_out, N, t_idx, stream);
}

extern "C" int idff_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x,
                               const float* d_x0, float* d_out, int64_t N, void* stream) {
  int rc = check_field_args(w, p, "idff_field_eval");
  if (rc) return rc;
  if (N == 0) return IDFF_OK;
  IDFF_CHECK_ARG(N > 0 && d_x && d_out, "idff_field_eval: null pointer or negative N");
  IDFF_CHECK_ARG(p->variant != IDFF_FIELD_CFM || t == 0.0f || d_x0, "idff_field_eval: the CFM field needs x0 (captured at t == 0) when t != 0");
  DeviceGuard g(w->device);
  if (p->impl == IDFF_IMPL_LAYERED) return wide_field_eval(w, p, t, d_x, d_x0, d_out, N, stream);
  if (p->impl == IDFF_IMPL_TENSOR16 && p->variant == IDFF_FIELD_MD) return md16_field_eval(w, p, t, d_x, d_out, N, stream);
  // Single evaluations take the tensor-core kernels at every size they serve (round 2: measured 28 vs 91 us for one order-2
  // evaluation of 256..2048 particles, 23 vs 127 us at order 3, 18.5 vs 41..50 us for the score head --
  // profiles/microbench/field_eval_latency.py; AUTO adds the safety net's scan, ~4 us).
  if (p->impl == IDFF_IMPL_AUTO && p->variant == IDFF_FIELD_MD && N >= 256 && md16_supports(w, p)) {   // tensor kernel + safety net
    int rc2 = keep_input(const_cast<idff_weights_t*>(w), &d_x, d_out, false, (size_t)N * p->dim * sizeof(float), stream);
    if (rc2) return rc2;
    rc2 = md16_field_eval(w, p, t, d_x, d_out, N, stream);
    if (rc2) return rc2;
    Stage st;
    make_stage(*p, t, &st);
    return fixup_with_tables(w, p, 0, 1, &st, nullptr, 0, d_x, nullptr, d_out, nullptr, N, stream);
  }
  // order 3 (three jets) on the 3-256-256-256-2 network: the fused three-jet tcgen05 kernel
  if ((p->impl == IDFF_IMPL_TENSOR16 || p->impl == IDFF_IMPL_AUTO) && tensor16o3_supports(w, p))
    return tensor16o3_field_eval(w, p, t, d_x, d_out, N, p->impl == IDFF_IMPL_AUTO, stream);
  if (p->impl == IDFF_IMPL_AUTO && w->v.w == 256 && N >= 2048 && !tensor16_supports(w, p)) {
    idff_field_params_t q = *p;   // order 3 at width 256: the layered 3xFP16 engine (see idff_sample_rk4)
    q.impl = IDFF_IMPL_LAYERED;
    if (wide_supports(w, &q)) return safe_wide_field_eval(w, &q, t, d_x, d_x0, d_out, N, stream);
  }
  if ((p->impl == IDFF_IMPL_TENSOR || p->impl == IDFF_IMPL_AUTO) && wide_supports(w, p) && (p->impl == IDFF_IMPL_TENSOR || N >= 2048))
    return p->impl == IDFF_IMPL_AUTO ? safe_wide_field_eval(w, p, t, d_x, d_x0, d_out, N, stream)
                                     : wide_field_eval(w, p, t, d_x, d_x0, d_out, N, stream);
  if (p->impl == IDFF_IMPL_TENSOR16 || (p->impl == IDFF_IMPL_AUTO && tensor16_supports(w, p)))
    return tensor16_field_eval(w, p, t, d_x, d_x0, d_out, N, stream);
  if (p->impl == IDFF_IMPL_TENSOR) return tensor_field_eval(w, p, t, d_x, d_x0, d_out, N, stream);
  if (p->impl == IDFF_IMPL_FUSED || (p->impl == IDFF_IMPL_AUTO && fused_supports(w, p))) {
    if (!fused_supports(w, p)) {
      set_error("idff_field_eval: the fused kernel does not support this shape/variant");
      return IDFF_EUNSUPPORTED;
    }
    return fused_field_eval(w, p, t, d_x, d_x0, d_out, N, stream);
  }
  return generic_field_eval(w, p, t, d_x, d_x0, d_out, N, stream);
}

// Would idff_sample_rk4 hand (p, N) to the layered tcgen05 engine?  (mirrors its dispatch; *q = p with the impl that engine expects)
static bool routes_to_wide(const idff_weights_t* w, const idff_field_params_t* p, int64_t N, idff_field_params_t* q) {
  *q = *p;
  if (p->impl == IDFF_IMPL_LAYERED) return wide_supports(w, p);
  if (p->impl == IDFF_IMPL_TENSOR16 || (p->impl == IDFF_IMPL_AUTO && (tensor16_supports(w, p) || tensor16o3_supports(w, p)))) return false;
  if (p->impl == IDFF_IMPL_AUTO && w->v.w == 256 && N >= 2048) {
    q->impl = IDFF_IMPL_LAYERED;
    if (wide_supports(w, q)) return true;
    q->impl = p->impl;
  }
  return (p->impl == IDFF_IMPL_TENSOR || p->impl == IDFF_IMPL_AUTO) && wide_supports(w, p) && (p->impl == IDFF_IMPL_TENSOR || N >= 2048);
}

extern "C" int idff_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0,
                               const float* h_t_grid, int32_t n_grid, float* d_out_final, float* d_out_traj, int64_t N,
                               void* stream) {
  int rc = check_field_args(w, p, "idff_sample_rk4");
  if (rc) return rc;
  IDFF_CHECK_ARG(p->variant != IDFF_FIELD_MD, "idff_sample_rk4: the MD field is an SDE drift; use idff_sample_sde");
  IDFF_CHECK_ARG(h_t_grid && n_grid >= 1 && n_grid <= 100000, "idff_sample_rk4: bad time grid (n_grid=%d)", n_grid);
  for (int i = 1; i < n_grid; ++i)
    IDFF_CHECK_ARG(h_t_grid[i] > h_t_grid[i - 1], "idff_sample_rk4: t_grid must be strictly increasing");
  if (N == 0) return IDFF_OK;
  IDFF_CHECK_ARG(N > 0 && d_x0 && d_out_final, "idff_sample_rk4: null pointer or negative N");
  // CFMModel captures x0 inside forward() at t == 0 only (Autograd_example_order2.py:322-324): a solve that does not start
  // there has no x0 in the reference either (AttributeError); solves longer than one launch re-read the caller's d_x0
  IDFF_CHECK_ARG(p->variant != IDFF_FIELD_CFM || n_grid < 2 || h_t_grid[0] == 0.0f,
                 "idff_sample_rk4: the CFM field captures x0 at t == 0; t_grid[0] = %g", (double)h_t_grid[0]);
  IDFF_CHECK_ARG(p->variant != IDFF_FIELD_CFM || n_grid - 1 <= kMaxIntervals || d_out_final != d_x0,
                 "idff_sample_rk4: a CFM solve of more than %d intervals cannot run in place (x0 is re-read)", kMaxIntervals);
  DeviceGuard g(w->device);
  const size_t row_bytes = (size_t)N * p->dim * sizeof(float);
  if (d_out_traj) IDFF_CUDA(cudaMemcpyAsync(d_out_traj, d_x0, row_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  const int n_iv = n_grid - 1;
  if (n_iv == 0) {
    if (d_out_final != d_x0)
      IDFF_CUDA(cudaMemcpyAsync(d_out_final, d_x0, row_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return IDFF_OK;
  }
  // stage times exactly as torchdiffeq rk_common.rk4_alt_step_func forms them (fp32 tensor * Python float)
  std::vector<Stage> st(4 * (size_t)n_iv);
  std::vector<float> dts(n_iv);
  const float third = (float)(1.0 / 3.0), two_thirds = (float)(2.0 / 3.0);
  for (int i = 0; i < n_iv; ++i) {
    const float t0 = h_t_grid[i], t1 = h_t_grid[i + 1];
    volatile float dt = t1 - t0;
    volatile float a = dt * third, b = dt * two_thirds;
    volatile float tb = t0 + a, tc = t0 + b;
    dts[i] = dt;
    make_stage(*p, t0, &st[4 * i + 0]);
    make_stage(*p, tb, &st[4 * i + 1]);
    make_stage(*p, tc, &st[4 * i + 2]);
    make_stage(*p, t1, &st[4 * i + 3]);
  }
  if (p->impl == IDFF_IMPL_LAYERED)
    return wide_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream);
  if ((p->impl == IDFF_IMPL_TENSOR16 || p->impl == IDFF_IMPL_AUTO) && tensor16o3_supports(w, p))   // order 3: three jets
    return tensor16o3_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, p->impl == IDFF_IMPL_AUTO, stream);
  // opt-in (idff_tune_small_n_route): small solves of the order-1 / order-2 field (the reference's own call: 2048 particles,
  // Autograd_example_order2.py:134-137) on the three-jet kernel's 32-particle tiles -- twice as many SMs as k_tc16's 64-particle
  // tiles; the idle third jet costs less than the idle SMs (128 vs 160 us at nfe 2 up to 4736 particles, slower beyond)
  if (p->impl == IDFF_IMPL_AUTO && N <= g_small_n && tensor16_supports(w, p) && tensor16o3_can_run(w, p))
    return tensor16o3_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, true, stream);
  if (p->impl == IDFF_IMPL_TENSOR16 || (p->impl == IDFF_IMPL_AUTO && tensor16_supports(w, p)))
    return tensor16_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream);
  if (p->impl == IDFF_IMPL_AUTO && w->v.w == 256 && N >= 2048 && !tensor16_supports(w, p)) {
    // three jets do not fit the fused tensor kernel; measured: layered 3xFP16 engine 35 / 81 / 120-135 M particle-steps/s
    // at N = 2048 / 8192 / >= 32768 vs 20 / 41 / 48 M for the FFMA2 kernel
    idff_field_params_t q = *p;
    q.impl = IDFF_IMPL_LAYERED;
    if (wide_supports(w, &q))
      return safe_wide_sample_rk4(w, &q, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream);
  }
  if ((p->impl == IDFF_IMPL_TENSOR || p->impl == IDFF_IMPL_AUTO) && wide_supports(w, p) && (p->impl == IDFF_IMPL_TENSOR || N >= 2048))
    return p->impl == IDFF_IMPL_AUTO ? safe_wide_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream)
                                     : wide_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream);
  if (p->impl == IDFF_IMPL_TENSOR)
    return tensor_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream);
  const bool want_fused = p->impl == IDFF_IMPL_FUSED || (p->impl == IDFF_IMPL_AUTO && fused_supports(w, p));
  if (want_fused) {
    if (!fused_supports(w, p)) {
      set_error("idff_sample_rk4: the fused kernel does not support this shape/variant");
      return IDFF_EUNSUPPORTED;
    }
    return fused_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream);
  }
  return generic_sample_rk4(w, p, d_x0, st.data(), dts.data(), n_iv, d_out_final, d_out_traj, N, stream);
}

extern "C" int idff_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int32_t n_sets,
                                     const float* d_x0, const float* h_t_grid, int32_t n_grid, float* d_out, int64_t N,
                                     void* stream) {
  IDFF_CHECK_ARG(params && n_sets >= 1 && n_sets <= 100000, "idff_sample_rk4_sweep: bad parameter-set array (n_sets=%d)", n_sets);
  for (int g = 0; g < n_sets; ++g) {
    int rc = check_field_args(w, &params[g], "idff_sample_rk4_sweep");
    if (rc) return rc;
    IDFF_CHECK_ARG(params[g].variant != IDFF_FIELD_MD, "idff_sample_rk4_sweep: the MD field is an SDE drift");
    IDFF_CHECK_ARG(params[g].dim == params[0].dim, "idff_sample_rk4_sweep: parameter sets differ in dim");
  }
  IDFF_CHECK_ARG(h_t_grid && n_grid >= 1 && n_grid <= 100000, "idff_sample_rk4_sweep: bad time grid (n_grid=%d)", n_grid);
  for (int i = 1; i < n_grid; ++i)
    IDFF_CHECK_ARG(h_t_grid[i] > h_t_grid[i - 1], "idff_sample_rk4_sweep: t_grid must be strictly increasing");
  if (N == 0) return IDFF_OK;
  IDFF_CHECK_ARG(N > 0 && d_x0 && d_out, "idff_sample_rk4_sweep: null pointer or negative N");
  IDFF_CHECK_ARG(params[0].variant != IDFF_FIELD_CFM || n_grid < 2 || h_t_grid[0] == 0.0f,
                 "idff_sample_rk4_sweep: the CFM field captures x0 at t == 0; t_grid[0] = %g", (double)h_t_grid[0]);
  const int n_iv = n_grid - 1;
  const size_t set_floats = (size_t)N * params[0].dim;
  // one launch when every set is served by the 3xFP16 tensor kernel; otherwise one fused launch per set
  bool batched = n_iv >= 1 && n_iv <= kMaxSweepIntervals;
  for (int g = 0; g < n_sets && batched; ++g)
    batched = (params[g].impl == IDFF_IMPL_AUTO || params[g].impl == IDFF_IMPL_TENSOR16) && tensor16_supports(w, &params[g]) &&
              params[g].variant == params[0].variant && params[g].t_idx == params[0].t_idx;
  // every set needs three jets: one launch of the fused three-jet kernel
  bool batched3 = !batched && n_iv >= 1 && n_iv <= kMaxSweepIntervals;
  for (int g = 0; g < n_sets && batched3; ++g)
    batched3 = (params[g].impl == IDFF_IMPL_AUTO || params[g].impl == IDFF_IMPL_TENSOR16) && tensor16o3_supports(w, &params[g]) &&
               params[g].t_idx == params[0].t_idx;
  // wide networks: one batch on the layered engine when that engine would serve every set anyway
  std::vector<idff_field_params_t> wp;
  bool wide_batched = !batched && !batched3 && n_iv >= 1 && N % 128 == 0 && n_sets >= 2;
  if (wide_batched) {
    wp.resize(n_sets);
    for (int g = 0; g < n_sets && wide_batched; ++g) wide_batched = routes_to_wide(w, &params[g], N, &wp[g]);
    wide_batched = wide_batched && wide_sweep_supports(w, wp.data(), n_sets, N);
  }
  if (!batched && !batched3 && !wide_batched) {
    // Mixed sweep (the order-3 grid holds tuples with gamma3 == 0, which the fused tensor kernel serves, next to tuples that
    // need three jets): split the sets into classes that batch -- class 0: fused tensor kernel; class 1 + 2 nj + rev: layered
    // engine with that jet count -- solve every class of two or more sets as one batch into a staging buffer and scatter the
    // rows to their slots; the rest one call per set.
    std::vector<int> cls(n_sets, -1);
    std::vector<std::vector<int>> members(9);
    if (n_iv >= 1 && n_iv <= kMaxSweepIntervals) {
      for (int g = 0; g < n_sets; ++g) {
        const idff_field_params_t& q = params[g];
        idff_field_params_t qq;
        if ((q.impl == IDFF_IMPL_AUTO || q.impl == IDFF_IMPL_TENSOR16) && tensor16_supports(w, &q) && q.variant == params[0].variant &&
            q.t_idx == params[0].t_idx) {
          cls[g] = 0;
        } else if ((q.impl == IDFF_IMPL_AUTO || q.impl == IDFF_IMPL_TENSOR16) && tensor16o3_supports(w, &q) && q.t_idx == params[0].t_idx) {
          cls[g] = 8;   // fused three-jet kernel
        } else if (N % 128 == 0 && q.variant == IDFF_FIELD_MOMENTUM && routes_to_wide(w, &q, N, &qq)) {
          int nj, rev;
          jets_needed(q, &nj, &rev);
          cls[g] = 1 + 2 * (nj - 1) + (rev ? 1 : 0);
        }
        if (cls[g] >= 0) members[cls[g]].push_back(g);
      }
    }
    for (int c = 0; c < 9; ++c) {
      const std::vector<int>& mem = members[c];
      if (mem.size() < 2 || (int)mem.size() == n_sets) {   // (a class that is the whole sweep already failed to batch above)
        for (int g : mem) cls[g] = -1;
        continue;
      }
      std::vector<idff_field_params_t> sub(mem.size());
      for (size_t k = 0; k < mem.size(); ++k) sub[k] = params[mem[k]];
      idff_weights_t* wm = const_cast<idff_weights_t*>(w);
      const size_t bytes = mem.size() * set_floats * sizeof(float);
      {
        DeviceGuard guard(w->device);
        if (wm->sweep_out_bytes < bytes) {
          if (wm->d_sweep_out) cudaFree(wm->d_sweep_out);
          wm->d_sweep_out = nullptr;
          wm->sweep_out_bytes = 0;
          IDFF_CUDA(cudaMalloc(&wm->d_sweep_out, bytes));
          wm->sweep_out_bytes = bytes;
        }
      }
      int rc = idff_sample_rk4_sweep(w, sub.data(), (int)sub.size(), d_x0, h_t_grid, n_grid, wm->d_sweep_out, N, stream);
      if (rc) return rc;
      DeviceGuard guard(w->device);
      size_t k = 0;
      while (k < mem.size()) {   // scatter, runs of consecutive sets in one copy
        size_t e = k + 1;
        while (e < mem.size() && mem[e] == mem[e - 1] + 1) ++e;
        IDFF_CUDA(cudaMemcpyAsync(d_out + (size_t)mem[k] * set_floats, wm->d_sweep_out + k * set_floats, (e - k) * set_floats * sizeof(float),
                                  cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        k = e;
      }
    }
    for (int g = 0; g < n_sets; ++g) {
      if (cls[g] >= 0) continue;
      int rc = idff_sample_rk4(w, &params[g], d_x0, h_t_grid, n_grid, d_out + (size_t)g * set_floats, nullptr, N, stream);
      if (rc) return rc;
    }
    return IDFF_OK;
  }
  DeviceGuard guard(w->device);
  std::vector<Stage> st((size_t)n_sets * 4 * n_iv);
  std::vector<float> dts(n_iv);
  const float third = (float)(1.0 / 3.0), two_thirds = (float)(2.0 / 3.0);
  for (int i = 0; i < n_iv; ++i) {   // stage times as in idff_sample_rk4 (torchdiffeq rk4_alt_step_func)
    const float t0 = h_t_grid[i], t1 = h_t_grid[i + 1];
    volatile float dt = t1 - t0;
    volatile float a = dt * third, b = dt * two_thirds;
    volatile float tb = t0 + a, tc = t0 + b;
    dts[i] = dt;
    for (int g = 0; g < n_sets; ++g) {
      Stage* s4 = &st[((size_t)g * n_iv + i) * 4];
      make_stage(params[g], t0, &s4[0]);
      make_stage(params[g], tb, &s4[1]);
      make_stage(params[g], tc, &s4[2]);
      make_stage(params[g], t1, &s4[3]);
    }
  }
  if (batched3) {
    bool safe = false;
    for (int g = 0; g < n_sets; ++g) safe = safe || params[g].impl == IDFF_IMPL_AUTO;
    return tensor16o3_sample_rk4_sweep(w, params, n_sets, d_x0, st.data(), dts.data(), n_iv, d_out, N, safe, stream);
  }
  if (wide_batched) {
    int rc = wide_sample_rk4_sweep(w, wp.data(), n_sets, d_x0, st.data(), dts.data(), n_iv, d_out, N, stream);
    bool safe = false;   // range safe unless every set asked for the raw engine
    for (int g = 0; g < n_sets; ++g) safe = safe || params[g].impl == IDFF_IMPL_AUTO;
    if (rc || !safe) return rc;
    return fixup_with_tables(w, &wp[0], 2, n_sets, st.data(), dts.data(), n_iv, d_x0, nullptr, d_out, nullptr, N, stream);
  }
  return tensor16_sample_rk4_sweep(w, params, n_sets, d_x0, st.data(), dts.data(), n_iv, d_out, N, stream);
}

// torchsde's fixed-step grid, the drift stage scalars, g values and output interpolation weights of one sdeint call
static void build_sde_table(const idff_field_params_t* p, const float* h_ts, int n_ts, int method, const std::vector<float>& grid,
                            SdeTable& tab) {
  const int n_steps = (int)grid.size() - 1;
  memset(&tab, 0, sizeof(tab));
  tab.n_steps = n_steps;
  tab.method = method;
  const float sig2f = (float)(p->sigma * p->sigma);
  const float C0[3] = {0.0f, 1.0f, 0.5f}, C1[4] = {0.0f, 0.25f, 1.0f, 0.25f};
  for (int s = 0; s < n_steps; ++s) {
    const float t0 = grid[s];
    volatile float h = grid[s + 1] - t0;
    tab.h[s] = h;
    for (int q = 0; q < 3; ++q) {
      volatile float ch = C0[q] * h;
      volatile float tq = t0 + ch;
      make_stage(*p, tq, &tab.st[s][q]);
    }
    for (int q = 0; q < 4; ++q) {
      // g = ones * 180 * sigma**2 * sqrt(t * (1 - t))      MD_simulation.py:528-529
      volatile float ch = C1[q] * h;
      volatile float tq = t0 + ch;
      volatile float om = 1.0f - tq;
      volatile float pr = tq * om;
      volatile float sq = sqrtf(pr);
      volatile float k = 180.0f * sig2f;
      tab.gval[s][q] = k * sq;
    }
  }
  // outputs: ys[i] = linear_interp(prev_t, prev_y, curr_t, curr_y, ts[i]) after the first step with curr_t >= ts[i]
  int gi = 0, ne = 0;
  for (int i = 1; i < n_ts; ++i) {
    const float out_t = h_ts[i];
    while (grid[gi] < out_t) ++gi;
    if (gi == 0) {   // out_t == ts[0]: the solver returns y0 (no step taken yet)
      tab.emit_step[ne] = -1;
    } else {
      tab.emit_step[ne] = gi - 1;
    }
    const float tp = gi > 0 ? grid[gi - 1] : grid[0], tc = grid[gi];
    float w0 = 0.0f, w1 = 1.0f;
    if (gi > 0 && tc > tp) {   // torchsde interp.linear_interp: (t1-t)/(t1-t0)*y0 + (t-t0)/(t1-t0)*y1
      volatile float span = tc - tp;
      volatile float a = tc - out_t, b = out_t - tp;
      w0 = a / span;
      w1 = b / span;
    }
    tab.emit_out[ne] = i;
    tab.emit_w0[ne] = w0;
    tab.emit_w1[ne] = w1;
    ++ne;
  }
  tab.n_emit = ne;
}

extern "C" int idff_sde_num_steps(const float* h_ts, int32_t n_ts, float dt) {
  IDFF_CHECK_ARG(h_ts && n_ts >= 1 && dt > 0.0f, "idff_sde_num_steps: bad argument");
  return (int)sde_grid(h_ts, n_ts, dt).size() - 1;
}

extern "C" int idff_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const float* d_y0,
                               const float* h_ts, int32_t n_ts, float dt, int32_t method, const float* d_I_k,
                               const float* d_I_k0, float* d_out, int64_t N, void* stream) {
  int rc = check_field_args(w, p, "idff_sample_sde");
  if (rc) return rc;
  IDFF_CHECK_ARG(p->variant == IDFF_FIELD_MD, "idff_sample_sde: only the MD field (SDE_cal) has a diffusion term");
  IDFF_CHECK_ARG(h_ts && n_ts >= 1 && dt > 0.0f, "idff_sample_sde: bad ts/dt");
  IDFF_CHECK_ARG(method == 0 || method == 1, "idff_sample_sde: method %d (0 srk, 1 euler)", method);
  for (int i = 1; i < n_ts; ++i) IDFF_CHECK_ARG(h_ts[i] >= h_ts[i - 1], "idff_sample_sde: ts must be non-decreasing");
  if (N == 0) return IDFF_OK;
  IDFF_CHECK_ARG(N > 0 && d_y0 && d_out, "idff_sample_sde: null pointer or negative N");
  const std::vector<float> grid = sde_grid(h_ts, n_ts, dt);
  const int n_steps = (int)grid.size() - 1;
  IDFF_CHECK_ARG(n_steps == 0 || d_I_k, "idff_sample_sde: Brownian increments d_I_k are required");
  IDFF_CHECK_ARG(n_steps == 0 || method == 1 || d_I_k0, "idff_sample_sde: srk needs the space-time Levy term d_I_k0");
  if (n_steps > kMaxSdeSteps || n_ts - 1 > kMaxSdeEmit) {
    set_error("idff_sample_sde: %d steps / %d outputs exceed the per-call limits (%d / %d)", n_steps, n_ts - 1, kMaxSdeSteps, kMaxSdeEmit);
    return IDFF_EUNSUPPORTED;
  }
  static thread_local SdeTable tab;
  build_sde_table(p, h_ts, n_ts, method, grid, tab);
  const int ne = tab.n_emit;
  DeviceGuard g(w->device);
  // outputs requested at ts[0] itself (emit_step == -1) are y0
  for (int e = 0; e < ne; ++e)
    if (tab.emit_step[e] < 0)
      IDFF_CUDA(cudaMemcpyAsync(d_out + (size_t)tab.emit_out[e] * N * p->dim, d_y0, (size_t)N * p->dim * sizeof(float),
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (p->impl == IDFF_IMPL_TENSOR16) return md16_sample_sde(w, p, tab, d_y0, d_I_k, d_I_k0, d_out, N, stream);
  // default route at batch: the tensor kernel, then the fp32 safety net for trajectories that left the fp16 range
  // (measured: 84 / 362 / 431 M drift evaluations/s at K = 2^10 / 2^13 / 2^17 vs 27 / 112 / 128 M for the FFMA2 kernel)
  if (p->impl == IDFF_IMPL_AUTO && N >= 256 && n_steps > 0 && md16_supports(w, p)) {
    rc = md16_sample_sde(w, p, tab, d_y0, d_I_k, d_I_k0, d_out, N, stream);
    if (rc) return rc;
    return fixup_nonfinite_sde(w, p, tab, n_ts, d_y0, d_I_k, d_I_k0, d_out, N, stream);
  }
  const bool want_fused = p->impl == IDFF_IMPL_FUSED || (p->impl == IDFF_IMPL_AUTO && fused_supports(w, p));
  if (want_fused) {
    if (!fused_supports(w, p)) {
      set_error("idff_sample_sde: the fused kernel does not support this shape/variant");
      return IDFF_EUNSUPPORTED;
    }
    return fused_sample_sde(w, p, tab, d_y0, d_I_k, d_I_k0, d_out, N, stream);
  }
  return generic_sample_sde(w, p, tab, d_y0, d_I_k, d_I_k0, d_out, N, stream);
}

// ---- the autoregressive PolyALA generator in one launch ------------------------------------------------
namespace idff {
bool md_chain_supports(const idff_weights_t* w, const idff_field_params_t* p, int64_t K);
void md_chain_set_fma(int on);
int md_chain_run(const idff_weights_t* w, const idff_field_params_t* p, int n_frames, const SdeTable& tab0, const Stage* stages,
                 const float* d_x_init, const float* d_I_k, const float* d_I_k0, float* d_y_hat, int64_t K, void* stream,
                 int n_sets, const float* gvals, int nj_sets, int reverse_sets);
}

// idff_md_generate (n_sets == 1) and idff_md_generate_sweep: argument checks, per-(set, frame) stage tables, one launch.
// frame f of a set: init_ind = f, gamma2 / (1 + f)   (MD_simulation.py:201); everything else is shared by the frames.
static int md_generate_impl(const char* who, const idff_weights_t* w, const idff_field_params_t* p_sets, int n_sets, int32_t n_frames,
                            const float* d_x_init, const float* h_ts, int32_t n_ts, float dt, int32_t method, const float* d_I_k,
                            const float* d_I_k0, float* d_y_hat, int64_t K, void* stream) {
  for (int i = 0; i < n_sets; ++i) {
    int rc = check_field_args(w, p_sets + i, who);
    if (rc) return rc;
    IDFF_CHECK_ARG(p_sets[i].variant == IDFF_FIELD_MD && p_sets[i].dim == p_sets[0].dim,
                   "%s: the generator integrates SDE_cal (variant IDFF_FIELD_MD), every set in the same dimension", who);
  }
  IDFF_CHECK_ARG(n_frames >= 1 && n_frames <= w->v.n_tidx, "%s: %d frames, the embedding has %d rows", who, n_frames, w->v.n_tidx);
  IDFF_CHECK_ARG(h_ts && n_ts >= 2 && dt > 0.0f, "%s: bad ts/dt", who);
  IDFF_CHECK_ARG(method == 0 || method == 1, "%s: method %d (0 srk, 1 euler)", who, method);
  for (int i = 1; i < n_ts; ++i) IDFF_CHECK_ARG(h_ts[i] >= h_ts[i - 1], "%s: ts must be non-decreasing", who);
  IDFF_CHECK_ARG(h_ts[n_ts - 1] > h_ts[0], "%s: empty time span", who);
  if (K == 0) return IDFF_OK;
  IDFF_CHECK_ARG(K > 0 && d_x_init && d_y_hat && d_I_k && (method == 1 || d_I_k0), "%s: null pointer or negative K", who);
  const std::vector<float> grid = sde_grid(h_ts, n_ts, dt);
  const int n_steps = (int)grid.size() - 1;
  if (n_steps > kMaxSdeSteps || n_ts - 1 > kMaxSdeEmit) {
    set_error("%s: %d steps / %d outputs exceed the per-call limits (%d / %d)", who, n_steps, n_ts - 1, kMaxSdeSteps, kMaxSdeEmit);
    return IDFF_EUNSUPPORTED;
  }
  if (!md_chain_supports(w, p_sets, K)) {
    set_error("%s: needs a width-128 network with at most 63 state dimensions; use one idff_sample_sde call per frame otherwise", who);
    return IDFF_EUNSUPPORTED;
  }
  static thread_local SdeTable tab0, tabf;
  std::vector<Stage> stages((size_t)n_sets * n_frames * n_steps * 3);
  std::vector<float> gvals(n_sets > 1 ? (size_t)n_sets * kMaxSdeSteps * 4 : 0, 0.0f);   // one set: the values travel in the kernel arguments
  int nj_max = 1, reverse_any = 0;
  for (int i = 0; i < n_sets; ++i) {
    int nj, rev;
    jets_needed(p_sets[i], &nj, &rev);
    nj_max = nj > nj_max ? nj : nj_max;
    reverse_any |= rev;
    for (int f = 0; f < n_frames; ++f) {
      idff_field_params_t q = p_sets[i];
      q.t_idx = f;
      q.gamma[1] = p_sets[i].gamma[1] / (double)(1 + f);
      SdeTable& t = (i == 0 && f == 0) ? tab0 : tabf;
      build_sde_table(&q, h_ts, n_ts, method, grid, t);
      for (int s = 0; s < n_steps; ++s)
        for (int g = 0; g < 3; ++g) stages[(((size_t)i * n_frames + f) * n_steps + s) * 3 + g] = t.st[s][g];
      if (f == 0 && n_sets > 1)
        for (int s = 0; s < n_steps; ++s)
          for (int g = 0; g < 4; ++g) gvals[((size_t)i * kMaxSdeSteps + s) * 4 + g] = t.gval[s][g];
    }
  }
  DeviceGuard guard(w->device);
  return md_chain_run(w, p_sets, n_frames, tab0, stages.data(), d_x_init, d_I_k, d_I_k0, d_y_hat, K, stream, n_sets,
                      n_sets > 1 ? gvals.data() : nullptr, nj_max, reverse_any);
}

extern "C" int idff_md_generate(const idff_weights_t* w, const idff_field_params_t* p, int32_t n_frames, const float* d_x_init,
                                const float* h_ts, int32_t n_ts, float dt, int32_t method, const float* d_I_k,
                                const float* d_I_k0, float* d_y_hat, int64_t K, void* stream) {
  IDFF_CHECK_ARG(p, "idff_md_generate: null field parameters");
  return md_generate_impl("idff_md_generate", w, p, 1, n_frames, d_x_init, h_ts, n_ts, dt, method, d_I_k, d_I_k0, d_y_hat, K, stream);
}

// The (sigma, gamma1, gamma2) sweep of TrajectoryGenerator.evaluate_model (MD_simulation.py:159-194) at one nfe: every parameter
// set generates its K trajectories in the SAME launch (a tile of 4 trajectories never mixes sets; per-set stage tables and
// diffusion values in global memory).  Rows are set-major: d_x_init [n_sets * K][D], increments [n_frames][n_steps][n_sets * K][D],
// d_y_hat [n_frames][n_sets * K][D]; set i's rows are exactly what idff_md_generate returns for p_sets[i] on those inputs.
extern "C" int idff_md_generate_sweep(const idff_weights_t* w, const idff_field_params_t* p_sets, int32_t n_sets, int32_t n_frames,
                                      const float* d_x_init, const float* h_ts, int32_t n_ts, float dt, int32_t method,
                                      const float* d_I_k, const float* d_I_k0, float* d_y_hat, int64_t K, void* stream) {
  IDFF_CHECK_ARG(p_sets && n_sets >= 1 && n_sets <= 4096, "idff_md_generate_sweep: %d parameter sets outside [1, 4096]", n_sets);
  return md_generate_impl("idff_md_generate_sweep", w, p_sets, n_sets, n_frames, d_x_init, h_ts, n_ts, dt, method, d_I_k, d_I_k0,
                          d_y_hat, K, stream);
}

namespace idff { void wide_set_dense3(int on); }
extern "C" int idff_debug_wide_dense3(int32_t on) {
  idff::wide_set_dense3(on);
  return IDFF_OK;
}

extern "C" int idff_debug_md_chain_fma(int32_t on) {
  md_chain_set_fma(on);
  return IDFF_OK;
}
