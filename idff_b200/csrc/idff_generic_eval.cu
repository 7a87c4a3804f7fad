// Generic single-evaluation kernel (forward(t, x) / SDE_cal.f(t, y) / plain MLP forward) + shared host planning.
#include "idff_field_generic.cuh"

namespace idff {

template <int NJ, int PB, bool MULTI>
__global__ void __launch_bounds__(256) k_field_eval(NetView net, FieldCfg cfg, Stage st, const float* __restrict__ x,
                                                    const float* __restrict__ x0, float* __restrict__ out, long N) {
  extern __shared__ __align__(16) float smem[];
  SmemPlan s = carve<NJ, PB>(smem, net, cfg.D);
  const int D = cfg.D;
  const int in_cols = cfg.variant == kRaw ? net.din : D;
  const int out_cols = cfg.variant == kRaw ? net.dout : D;
  for (long g = blockIdx.x; g * PB < N; g += gridDim.x) {
    const long base = g * PB;
    for (int o = threadIdx.x; o < PB * in_cols; o += blockDim.x) {
      const long row = base + o / in_cols;
      s.xs[o] = row < N ? x[row * in_cols + (o % in_cols)] : 0.0f;
      if (cfg.variant == IDFF_FIELD_CFM && x0 != nullptr) s.x0s[o] = row < N ? x0[row * D + (o % D)] : 0.0f;
    }
    __syncthreads();
    if (cfg.variant == kRaw) {
      eval_field<1, false, PB, MULTI>(net, cfg, st, s.xs, s.xs, s.x0s, s.fs, s.actA, s.actB, s.pred, s.gout);
      for (int o = threadIdx.x; o < PB * out_cols; o += blockDim.x) {
        const long row = base + o / out_cols;
        if (row < N) out[row * out_cols + (o % out_cols)] = s.pred[o];
      }
    } else {
      eval_dispatch<NJ, PB, MULTI>(net, cfg, st, s.xs, s.x0s, s.fs, s.actA, s.actB, s.pred, s.gout);
      for (int o = threadIdx.x; o < PB * D; o += blockDim.x) {
        const long row = base + o / D;
        if (row < N) out[row * D + (o % D)] = s.fs[o];
      }
    }
    __syncthreads();
  }
}

int plan_generic(const idff_weights_t* w, int D, int nj, int64_t N, GenPlan* pl) {
  const NetView& net = w->v;
  if (net.w > 2048) {
    set_error("generic kernel: hidden width %d > 2048 is not supported", net.w);
    return IDFF_EUNSUPPORTED;
  }
  pl->nj = nj;
  pl->multi = net.w > 256;
  pl->pb = pl->multi ? 4 : 8;
  pl->threads = net.w >= 256 ? 256 : net.w;
  pl->smem = generic_smem_bytes(net, D, nj, pl->pb);
  if (pl->smem > 227 * 1024) {
    set_error("generic kernel: width %d / dim %d needs %zu B of shared memory (limit 232448)", net.w, D, pl->smem);
    return IDFF_EUNSUPPORTED;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  size_t per_sm = 227 * 1024 / (pl->smem + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const long groups = ((long)N + pl->pb - 1) / pl->pb;
  const long cap = (long)sms * (long)per_sm;
  pl->grid = (int)(groups < cap ? (groups > 0 ? groups : 1) : cap);
  return IDFF_OK;
}

// number of jets an interior evaluation needs and whether the reverse sweep runs at all
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse) {
  *nj = 1;
  *reverse = 0;
  if (p.variant == IDFF_FIELD_MOMENTUM || p.variant == IDFF_FIELD_MD) {
    const int order = p.variant == IDFF_FIELD_MD ? 2 : p.order;
    const double c0 = 0.5 * (2.0 * p.gamma[0] - 1.0);
    int top = 0;
    if (c0 != 0.0) top = 1;
    for (int i = 1; i < order; ++i)
      if (p.gamma[i] != 0.0) top = i + 1;
    *nj = top > 0 ? top : 1;
    *reverse = top > 0;
  }
}

int check_field_args(const idff_weights_t* w, const idff_field_params_t* p, const char* who) {
  IDFF_CHECK_ARG(w && p, "%s: null handle or params", who);
  const NetView& v = w->v;
  IDFF_CHECK_ARG(p->dim >= 1 && p->dim <= 256, "%s: dim %d out of range", who, p->dim);
  switch (p->variant) {
    case IDFF_FIELD_MOMENTUM:
      IDFF_CHECK_ARG(p->order >= 1 && p->order <= 3, "%s: order %d not in 1..3", who, p->order);
      /* fallthrough */
    case IDFF_FIELD_CFM:
      // the order-1 script's CFM comparator runs on its score-head network: cat[x, x, t] -> out[:, :dim] - x0
      // (example_order1_with_prediction_head.py:335-356)
      if (v.din == 2 * p->dim + 1 && v.dout == 2 * p->dim) break;
      /* fallthrough */
    case IDFF_FIELD_MD:
      IDFF_CHECK_ARG(v.din == p->dim + 1 && v.dout == p->dim,
                     "%s: network is %d->%d but the field needs %d->%d (cat[x,t] -> x1hat)", who, v.din, v.dout,
                     p->dim + 1, p->dim);
      break;
    case IDFF_FIELD_SCOREHEAD:
      IDFF_CHECK_ARG(v.din == 2 * p->dim + 1 && v.dout == 2 * p->dim,
                     "%s: score-head network is %d->%d but needs %d->%d (cat[x,x,t] -> x1hat|score)", who, v.din,
                     v.dout, 2 * p->dim + 1, 2 * p->dim);
      break;
    default:
      set_error("%s: unknown field variant %d", who, p->variant);
      return IDFF_EINVAL;
  }
  IDFF_CHECK_ARG(p->t_idx >= 0 && p->t_idx < v.n_tidx, "%s: t_idx %d outside the embedding table (%d rows)", who,
                 p->t_idx, v.n_tidx);
  return IDFF_OK;
}

int generic_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x,
                       const float* d_x0, float* d_out, int64_t N, void* stream) {
  FieldCfg cfg;
  int nj;
  cfg.variant = p->variant; cfg.D = p->dim; cfg.tidx = p->t_idx;
  jets_needed(*p, &nj, &cfg.reverse);
  Stage st;
  make_stage(*p, t, &st);
  GenPlan pl;
  int rc = plan_generic(w, p->dim, nj, N, &pl);
  if (rc) return rc;
  IDFF_GEN_DISPATCH(k_field_eval, pl, stream, w->v, cfg, st, d_x, d_x0, d_out, (long)N);
}

int generic_mlp_forward(const idff_weights_t* w, const float* d_in, float* d_out, int64_t N, int t_idx, void* stream) {
  FieldCfg cfg;
  cfg.variant = kRaw; cfg.D = w->v.din; cfg.tidx = t_idx; cfg.reverse = 0;
  Stage st;
  memset(&st, 0, sizeof(st));
  GenPlan pl;
  int rc = plan_generic(w, w->v.din, 1, N, &pl);
  if (rc) return rc;
  IDFF_GEN_DISPATCH(k_field_eval, pl, stream, w->v, cfg, st, d_in, nullptr, d_out, (long)N);
}

}  // namespace idff
