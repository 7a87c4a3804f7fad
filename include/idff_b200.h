/* idff_b200.h -- C-ABI of the B200-native IDFF hot path (libidff_b200.so).
 *
 * The reference (MrRezaeiUofT/IDFF) is pure Python and has no FFI; each entry point below names the
 * reference interface it replaces (file:line under the reference root).  The host-side mirror of the
 * reference's call conventions (torchcfm matchers, forward(t,x) modules, f/g SDE objects) lives in
 * the modules of idff_b200/ and reaches these functions through ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative IDFF_E* code; it never throws and never aborts;
 *     idff_last_error() returns a thread-local message for the last failure on the calling thread;
 *   - pointers named d_* are DEVICE pointers owned by the caller (torch tensors), fp32 row-major unless
 *     stated; pointers named h_* are HOST pointers;  `stream` is a cudaStream_t passed as void*;
 *   - the only persistent allocation is the weights handle.  A handle also owns grow-only scratch that some routes size on
 *     first use (the order-3 stash of the FFMA2 kernel, the buffers of the layered engine, the stage tables of the last
 *     sweep, the input copy / tables of the AUTO route's safety net): those calls may cudaMalloc once and are not
 *     CUDA-graph-capturable the first time; steady-state calls allocate nothing;
 *   - a handle serves ONE in-flight call at a time (one stream, one host thread): the scratch above and the per-handle
 *     counters are not keyed by stream.  Use one handle per stream for concurrent sampling with the same network;
 *   - no process-global state on the product path: counters live in the handle (idff_weights_nonfinite).  The idff_debug_*
 *     switches (alternative evaluators / algorithms) are thread-local: they affect calls made by the thread that set them and
 *     nothing else; the idff_debug_*_stamps read-outs return device-side phase timestamps of the last launch of that kernel in
 *     the process, and idff_launch_count() is one process-wide atomic counter -- diagnostics for tests and profiling only.
 */
#ifndef IDFF_B200_H
#define IDFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDFF_OK 0
#define IDFF_EINVAL (-1)   /* bad argument (shape, null pointer, unsupported size) */
#define IDFF_ECUDA (-2)    /* CUDA runtime error (message in idff_last_error) */
#define IDFF_ENOMEM (-3)
#define IDFF_EUNSUPPORTED (-4)

/* field variants (the reference's four wrappers around the MLP) */
#define IDFF_FIELD_MOMENTUM 0 /* CustomNetModel, 2D_examples/Autograd_example_order2.py:36-76 (order 2),
                                 Autograd_example_order3.py:34-77 (order 3); order 1 = g2 = g3 = 0 */
#define IDFF_FIELD_SCOREHEAD 1 /* CustomNetModel, 2D_examples/example_order1_with_prediction_head.py:64-96 */
#define IDFF_FIELD_CFM 2       /* CFMModel, 2D_examples/Autograd_example_order2.py:310-332; on a score-head network (din = 2 dim + 1,
                                  dout = 2 dim) the order-1 script's variant, example_order1_with_prediction_head.py:335-356 */
#define IDFF_FIELD_MD 3        /* SDE_cal.f, timeseris_example/MD_simulation.py:497-526 */

/* which kernel family serves a sampler call */
#define IDFF_IMPL_AUTO 0    /* fastest kernel that serves the shape, and RANGE SAFE: where that kernel carries fp32 values as fp16
                               pairs (TENSOR16, LAYERED) the call also runs the fp32 safety net (idff_fixup.cu) on the same
                               stream -- particles whose result came out non-finite (an activation, jet or adjoint beyond 65504)
                               are re-solved from their inputs in fp32 FMA arithmetic, so AUTO never returns inf/NaN where the
                               reference's fp32 path is finite.  No host synchronisation; idff_weights_nonfinite() counts them */
#define IDFF_IMPL_GENERIC 1 /* any width / dimension: one thread per hidden unit, weights from L2 */
#define IDFF_IMPL_FUSED 2   /* persistent warp-tiled kernel, weights streamed by TMA bulk copies */
#define IDFF_IMPL_TENSOR 3  /* tcgen05: fused 3xTF32 kernel (width 256, dim 2, order <= 2) or the layered 3xFP16 engine (width >= 512) */
#define IDFF_IMPL_LAYERED 4 /* force the layer-by-layer tcgen05 3xFP16 engine (width >= 256, dim 2, orders 1..3) */
#define IDFF_IMPL_TENSOR16 5 /* the raw tcgen05 3xFP16 fused kernels (width 256, dim 2: k_tc16 for up to two jets -- orders 1, 2, and
                                order 3 with gamma3 == 0 -- and the three-jet k_tc16o3 for order 3), WITHOUT the safety net:
                                values beyond the fp16 range (|activation|, |jet|, |adjoint| > 65504) make that particle's
                                output NaN (counted by idff_weights_nonfinite()).  AUTO = this kernel + the safety net */

typedef struct idff_weights idff_weights_t; /* opaque: packed device copies of one 4-layer SELU MLP */

typedef struct idff_field_params {
  int32_t variant;   /* IDFF_FIELD_* */
  int32_t order;     /* 1..3: number of gamma terms (MOMENTUM); 2 for MD; ignored otherwise */
  int32_t dim;       /* state dimension D of one particle */
  int32_t t_idx;     /* MD: embedding row (SDE_cal init_ind, MD_simulation.py:501); else 0 */
  int32_t impl;      /* IDFF_IMPL_* */
  int32_t reserved;
  double sigma;      /* sigma of the probability path (order2.py:30; MD_simulation.py:542) -- Python floats */
  double gamma[3];   /* gamma1..gamma3 */
  double end_div;    /* divisor of the t == 1 branch: 1e-2 (order2.py:60) or SDE_cal.dt (MD_simulation.py:510) */
} idff_field_params_t;

int idff_version(void);
/* first 16 hex digits of the sha256 over the library's sources (the .cu and .cuh files of csrc/ in sorted order, this header, the
 * Makefile) as they were when the library was built: lets a caller prove the loaded binary matches its checkout */
const char* idff_source_hash(void);
const char* idff_last_error(void);
/* number of kernels this library has launched from the calling process so far (bench.py's gpu_launches) */
uint64_t idff_launch_count(void);

/* ---- a1/a2: the networks, torchcfm/models/models.py:4-21 (MLP) and :23-42 (MLP_Embedding) ------------
 * h_W[l] is nn.Linear.weight of layer l, row-major [dims[l+1]][dims[l]] (+emb_dim extra input columns on
 * layer 0 when an embedding is given); h_b[l] its bias.  n_layers must be 4 and dims[1]==dims[2]==dims[3].
 * h_emb (optional) is nn.Embedding.weight [emb_rows][emb_dim]; because every row of a batch shares one
 * time index during generation (MD_simulation.py:501) it is folded into a per-index layer-0 bias table. */
int idff_weights_create(const float* const* h_W, const float* const* h_b, const int32_t* dims, int32_t n_layers,
                        const float* h_emb, int32_t emb_rows, int32_t emb_dim, int32_t device,
                        idff_weights_t** out);
int idff_weights_destroy(idff_weights_t* w);

/* plain MLP forward on rows of cat[x, t] (models.py:20-21 / :40-42 with the folded embedding): d_in [N][dims[0]],
 * d_out [N][dims[4]]. */
int idff_mlp_forward(const idff_weights_t* w, const float* d_in, float* d_out, int64_t N, int32_t t_idx,
                     void* stream);

/* ---- a3..a7: one evaluation of the IDFF field f(t, x); backs forward(t, x) / SDE_cal.f(t, y) ------------
 * d_x [N][dim] -> d_out [N][dim].  d_x0 is only read by IDFF_FIELD_CFM at t != 0 (the x captured at t == 0). */
int idff_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x,
                    const float* d_x0, float* d_out, int64_t N, void* stream);

/* Host-only diagnostic: the per-evaluation scalars the kernels receive for time t, computed in fp32 in the
 * reference's operation order.  out[0..7] = {t, kind (0: t==0, 1: t==1, 2: interior), a = sqrt(1-Gamma sigma_t^2),
 * 1-t, c1, c2, c3, end_div}.  Needs no GPU. */
int idff_stage_scalars(const idff_field_params_t* p, float t, float* h_out8);

/* Diagnostic: clock64 phase timestamps of the last IDFF_IMPL_TENSOR launch (CTA 0, first tile, 2nd evaluation):
 * [0..12] epilogue warp 0 (layer-1 start, publish, accumulator-ready / publish of every layer pass, reductions,
 * stage update), [16+2g], [17+2g] MMA issue start / commit of layer pass g.  h_out32: 32 int64 on the host. */
int idff_debug_tensor_stamps(int64_t* h_out32);
/* same for the last IDFF_IMPL_TENSOR16 launch */
int idff_debug_tensor16_stamps(int64_t* h_out32);
/* Per-handle counters of the fp16-split kernels: h_out2[0] = particle-evaluations whose field value was not finite (fp16
 * operand overflow or non-finite input), h_out2[1] = rows the safety net of IDFF_IMPL_AUTO re-solved in fp32.  A blocking
 * 16-byte device-to-host copy; never called on the product path. */
int idff_weights_nonfinite(const idff_weights_t* w, uint64_t* h_out2);

/* ---- a8 + a3..a6 fused: torchdiffeq.odeint(model, x0, t_grid, method='rk4') (3/8 rule, grid == t_grid;
 * call sites Autograd_example_order2.py:167,213,291).  h_t_grid: n_grid fp32 times on the host.
 * d_out_final [N][dim] = traj[-1]; d_out_traj (optional) [n_grid][N][dim] = the whole traj. */
int idff_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0,
                    const float* h_t_grid, int32_t n_grid, float* d_out_final, float* d_out_traj, int64_t N,
                    void* stream);

/* ---- the gamma sweeps that call the sampler (find_best_gammas_mmd, Autograd_example_order2.py:281-303,
 * 8x8 tuples; Autograd_example_order3.py:288-314, 10^3 tuples): the same x0 [N][dim] and t grid solved under n_sets
 * parameter sets (params[g]: same variant and dim, any gammas / order / sigma).  d_out [n_sets][N][dim] = traj[-1] of every
 * set.  One launch for all sets where the 3xFP16 tensor kernel serves them (the reference's 2048 particles fill 32 of 148
 * SMs per tuple; 64 tuples fill the machine); sets that need three jets are solved as one batch on the layered engine (N a
 * multiple of 128); a mixed sweep is split into such classes and the rows scattered to their slots; whatever remains takes one
 * launch per set -- results are identical to calling idff_sample_rk4 per set in every case. */
int idff_sample_rk4_sweep(const idff_weights_t* w, const idff_field_params_t* params, int32_t n_sets, const float* d_x0,
                          const float* h_t_grid, int32_t n_grid, float* d_out, int64_t N, void* stream);

/* ---- a8 + a7 fused: torchsde.sdeint(sde, y0, ts, dt) for SDE_cal (MD_simulation.py:209,228); method 0 = srk
 * (torchsde default for ito/diagonal), 1 = euler.  Brownian increments are injected: d_I_k, d_I_k0
 * [n_steps][N][dim] with n_steps = idff_sde_num_steps(h_ts, n_ts, dt); d_I_k0 may be NULL for euler.
 * d_out [n_ts][N][dim] (linear interpolation at ts, as torchsde does). g(t,y) = 180 sigma^2 sqrt(t(1-t)). */
int idff_sde_num_steps(const float* h_ts, int32_t n_ts, float dt);
int idff_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const float* d_y0, const float* h_ts,
                    int32_t n_ts, float dt, int32_t method, const float* d_I_k, const float* d_I_k0, float* d_out,
                    int64_t N, void* stream);

/* ---- a9/a10: ConditionalFlowMatcher.sample_xt + compute_conditional_flow,
 * torchcfm/conditional_flow_matching.py:98-148 (= conditional_flow_matching_dynamic.py:101-151):
 *   xt = t*x1 + (1-t)*x0 + sigma_t*eps,  ut = x1 - x0      (bit-exact: no FMA contraction)
 * sigma_mode 0: sigma_t = sigma (:79-96); 1: sigma_t = sigma*sqrt(t(1-t)) (ExactOT :231-233).
 * d_idx0/d_idx1 (optional int64 [B]) re-pair rows first: x0[idx0[b]], x1[idx1[b]] (the OT plan's output,
 * conditional_flow_matching_dynamic.py:268 / optimal_transport.py:140-142).  d_t [B], others [B][D]. */
int idff_path_sample(const float* d_x0, const float* d_x1, const float* d_t, const float* d_eps,
                     const int64_t* d_idx0, const int64_t* d_idx1, float sigma, int32_t sigma_mode, float* d_xt,
                     float* d_ut, int64_t B, int64_t D, void* stream);

/* The same with eps = torch.randn_like(x0) (conditional_flow_matching.py:150-151,187) drawn INSIDE the kernel: Philox4x32-10
 * in ATen's element layout (distribution_elementwise_grid_stride_kernel: block 256, grid = min(SMs * maxThreadsPerSM / 256,
 * ceil(B D / 256)), thread idx = subsequence idx, curand_normal4 per iteration, element idx + S (4 j + ii) gets component ii),
 * so for the torch CUDA generator's current (seed, philox_offset) every element receives exactly the number randn_like would
 * have written -- eps never goes through HBM (36 instead of 44 B per row at D = 2) unless d_eps_out (optional [B][D]) asks for
 * it.  *h_offset_increment (optional) = what the caller must add to the generator's offset afterwards, as ATen would have
 * (idff_philox_plan returns it, and the grid, without launching).  B D < 2^31 (beyond that ATen splits the tensor into several
 * RNG launches: IDFF_EUNSUPPORTED); no index gather (the _dynamic matcher re-pairs first and injects eps). */
int idff_philox_plan(int64_t numel, int32_t device, int32_t* h_grid, uint64_t* h_offset_increment);
int idff_path_sample_philox(const float* d_x0, const float* d_x1, const float* d_t, uint64_t seed, uint64_t philox_offset,
                            float sigma, int32_t sigma_mode, float* d_xt, float* d_ut, float* d_eps_out, int64_t B, int64_t D,
                            uint64_t* h_offset_increment, void* stream);

/* ---- a11: compute_lambda, conditional_flow_matching.py:195-211 (dynamic = 0: 2 sigma_t^2/(sigma^2+1e-8))
 * and conditional_flow_matching_dynamic.py:198-214 (dynamic = 1: 2 sigma_t/(sigma^2+1e-8)). */
int idff_compute_lambda(const float* d_t, float sigma, int32_t sigma_mode, int32_t dynamic, float* d_out,
                        int64_t B, void* stream);

/* ---- a12: losses.  kind 0: flow loss mean(((vt-x1)/(1e-2+1-t))^2), Autograd_example_order2.py:105;
 * kind 1: score loss mean((st-logp)^2) with logp of example_order1_with_prediction_head.py:26-51,124-132
 *         (a = st, needs x1,x0,xt,t, sigma0_sq = the value passed at :126);
 * kind 2: PolyALA periodic loss sum_k mean((vt-(x1+k))^2), k in {0,+360,-360}, MD_simulation.py:137-141.
 * d_loss: one fp32.  d_grad (optional) [B][D]: d loss / d a, scaled by grad_scale.
 * d_scratch: >= idff_loss_scratch_bytes() bytes of device scratch (partial sums). */
int64_t idff_loss_scratch_bytes(void);
int idff_loss(int32_t kind, const float* d_a, const float* d_x1, const float* d_x0, const float* d_xt,
              const float* d_t, float sigma0_sq, float grad_scale, float* d_loss, float* d_grad, int64_t B,
              int64_t D, void* d_scratch, void* stream);

/* ---- a13: compute_mmd_rbf, Autograd_example_order2.py:260-278 (= compute_mmd, MD_simulation.py:50-86):
 * biased V-statistic with K = exp(-|a-b|^2 / (2 sigma_k^2)).  Accumulates into d_sums[3] (fp64, caller
 * zeroes them): sum of Kxx over rows [row_begin,row_end) x all x, sum of Kyy over the matching share of y
 * rows, sum of Kxy over x rows [row_begin,row_end) x all y -- so ranks can split the work by rows and add. */
int idff_mmd_rbf(const float* d_x, int64_t N, const float* d_y, int64_t M, int32_t D, float sigma_k,
                 double* d_sums, int64_t x_row_begin, int64_t x_row_end, int64_t y_row_begin, int64_t y_row_end,
                 void* stream);

/* The MMD half of a gamma sweep (find_best_gammas_mmd, Autograd_example_order2.py:293-295): one fixed sample set d_x [N][D]
 * against n_sets generated sets d_y [n_sets][M][D] (the output of idff_sample_rk4_sweep) in ONE launch, D <= 4.
 * d_sums [1 + 2 n_sets] fp64, zeroed by the caller: [0] = sum Kxx, [1 + 2g] = sum Kyy of set g, [2 + 2g] = sum Kxy of set g;
 * MMD_g = sums[0]/N^2 + sums[1+2g]/M^2 - 2 sums[2+2g]/(N M). */
int idff_mmd_rbf_sweep(const float* d_x, int64_t N, const float* d_y, int64_t M, int32_t n_sets, int32_t D, float sigma_k,
                       double* d_sums, void* stream);

/* ---- the 2-D training step (Autograd_example_order2.py:89-112 = Autograd_example_order3.py:90-109)
 *   t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1)   (path sample, conditional_flow_matching.py:153-193)
 *   vt = base_model(cat[xt, t])                                         (MLP 3-256-256-256-2, models.py:4-21)
 *   loss = mean(((1/(1e-2+1-t)) * (vt - FM.x1))**2); loss.backward()    (:105, :109)
 *   clip_grad_norm_(params, max_grad_norm); AdamW.step()                (:110-112)
 * as ONE persistent cooperative kernel that runs n_steps consecutive steps per launch: path sample, forward, loss,
 * backward, gradient-norm clipping and the AdamW update, no torch kernel in between.
 * The caller owns the state: d_state = idff_train_flow_state_floats() fp32 (16-byte aligned), zero-initialised, holding
 * parameters, AdamW moments and scratch.  idff_train_flow_offsets fills h_out12 = {offset of 0.weight [256][3], 0.bias,
 * 2.weight [256][256], 2.bias, 4.weight, 4.bias, 6.weight [2][256], 6.bias within a parameter block, parameter count,
 * state offset of the parameter block, of exp_avg, of exp_avg_sq} (floats; tensors in nn.Linear's own layout, so torch
 * parameters can be views of the state).  After writing the parameters call idff_train_flow_init (builds the transposed
 * forward copies); call it again whenever the parameters are changed from outside.
 * Inputs of step s: d_x0, d_x1, d_eps [n_steps][B][2] (d_eps may be NULL: sigma_t * eps = 0), d_t [n_steps][B].
 * B: multiple of 4 in [4, IDFF_TRAIN_MAX_BATCH].  steps_done: optimizer steps taken before this call (AdamW bias
 * correction).  max_grad_norm <= 0 disables clipping.  Outputs (each optional): d_loss [n_steps], d_gnorm [n_steps] (total
 * gradient norm before clipping), d_grad_dump [parameter count] (unclipped gradient of the last step, parameter-block
 * layout).  Other networks / losses train through torch (idff_b200/train.py GraphedFlowTrainer). */
#define IDFF_TRAIN_MAX_BATCH 512
typedef struct idff_adamw {
  double lr, beta1, beta2, eps, weight_decay; /* torch.optim.AdamW defaults: 1e-3 (the scripts use 2e-4), 0.9, 0.999, 1e-8, 1e-2 */
  double max_grad_norm;                       /* clip_grad_norm_ threshold: 1 (Autograd_example_order2.py:111) */
} idff_adamw_t;
int64_t idff_train_flow_state_floats(void);
int idff_train_flow_offsets(int64_t* h_out12);
int idff_train_flow_init(float* d_state, void* stream);
/* Diagnostic: clock64 stamps of the last step of the last idff_train_flow_steps launch, h_out [n_ctas][8]: per CTA
 * {step start, rows phase done, barrier 1 passed, weight-gradient job done, barrier 2 passed, AdamW done, barrier 3 passed, 0}. */
int idff_debug_train_stamps(int64_t* h_out, int32_t n_ctas);
/* Diagnostic (thread-local): on == 0 makes the layered engine run three-jet solves in the 32-particle column-tile geometry
 * (96 of 128 columns used) instead of the default 40-particle one (120 of 128). */
int idff_debug_wide_dense3(int32_t on);
/* Tuning knob (thread-local, default 0 = off): idff_sample_rk4 under IDFF_IMPL_AUTO sends order-1 / order-2 solves of at most n_max
 * particles on the 3-256-256-256-2 network to the 32-particle-tile kernel (the three-jet kernel with an idle jet) instead of
 * k_tc16's 64-particle tiles: twice as many SMs for a small batch -- measured 128 vs 160 us for the reference's 2048-particle call
 * (nfe 2), a gain up to 4736 particles (one wave of 148 CTAs), a loss beyond.  Off by default because the two kernels round
 * differently: with the knob on, a particle's result depends (in its last bits) on how many particles share the call, and a
 * sweep is no longer bit-identical to per-tuple solves.  Same parity bounds either way. */
int idff_tune_small_n_route(int64_t n_max);
/* Thread-local switch, default on: every group of 4 batch rows runs on a CTA pair (cluster of 2 that splits the hidden units,
 * activation halves exchanged through distributed shared memory with st.async hand-overs) whenever the batch allows it (B <= 296
 * on 148 SMs); on == 0 forces one CTA per row group.  The two modes sum the K-slices of a layer in different (fixed) orders, so
 * they agree to fp32 rounding, not bit for bit; each is bit-reproducible in itself. */
int idff_debug_train_pair_mode(int32_t on);
int idff_train_flow_steps(float* d_state, const float* d_x0, const float* d_x1, const float* d_t, const float* d_eps,
                          int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode, const idff_adamw_t* opt,
                          int64_t steps_done, float* d_loss, float* d_gnorm, float* d_grad_dump, void* stream);
/* The same persistent kernel for either network the 2-D scripts train (`net`):
 *   IDFF_TRAIN_NET_FLOW       MLP(dim=2): cat[xt, t] -> 256 -> 256 -> 256 -> 2, flow loss (the idff_train_flow_* entry points)
 *   IDFF_TRAIN_NET_SCOREHEAD  the order-1 script's step (example_order1_with_prediction_head.py:111-137): MLP(dim=4):
 *                             outs = model(cat[xt, xt, t]) (5 -> 256 -> 256 -> 256 -> 4), vt = outs[:, :2], st = outs[:, 2:],
 *                             loss = mean(((vt - x1) / (1e-2 + 1 - t))^2) + mean((st - logp)^2),
 *                             logp = log_prob_xt_given_x1_dimwise(xt, x1, x0, t, score_sigma) (:26-51; the script passes sigma
 *                             where sigma^2 is named, :126), clip_grad_norm_, AdamW.
 * State size, offsets (0.weight is [256][5], 6.weight [4][256] for the score head) and init per network; the inputs,
 * outputs and conventions are those of idff_train_flow_steps; score_sigma is ignored by IDFF_TRAIN_NET_FLOW. */
#define IDFF_TRAIN_NET_FLOW 0
#define IDFF_TRAIN_NET_SCOREHEAD 1
int64_t idff_train_state_floats(int32_t net);
int idff_train_offsets(int32_t net, int64_t* h_out12);
int idff_train_init(int32_t net, float* d_state, void* stream);
int idff_train_steps(int32_t net, float* d_state, const float* d_x0, const float* d_x1, const float* d_t, const float* d_eps,
                     int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode, float score_sigma, const idff_adamw_t* opt,
                     int64_t steps_done, float* d_loss, float* d_gnorm, float* d_grad_dump, void* stream);

/* ---- the PolyALA training step (TrajectoryGenerator.train_model, timeseris_example/MD_simulation.py:117-144):
 *   idx = randint(1, T, (B,)); x0 = dataset[idx - 1]; x1 = dataset[idx]                                   (:128-130)
 *   t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1)     (ExactOT matcher sigma 0.5, :121,132)
 *   vt = model(cat[xt, t], idx)       MLP_Embedding(dim=D, w=128, time_embed=T): cat[xt, t, Embedding[idx]] -> 128 -> 128 -> 128 -> D
 *   loss = mean((vt-x1)^2) + mean((vt-(x1+360))^2) + mean((vt-(x1-360))^2); loss.backward(); Adam().step()   (:137-144)
 * n_steps consecutive steps in ONE launch on a cluster of 8 CTAs whose shared memories hold every parameter and both Adam
 * moments for the whole launch (idff_trainer_md.cu).  The caller owns d_state = idff_train_md_state_floats(D, T) fp32,
 * zero-initialised then filled with the parameters: three blocks [parameters | exp_avg | exp_avg_sq];
 * idff_train_md_offsets fills h_out13 = {offsets of time_index_embedding.weight [T][128], net.0.weight [128][D+1+128],
 * net.0.bias, net.2.weight, net.2.bias, net.4.weight, net.4.bias, net.6.weight [D][128], net.6.bias inside a block (the
 * module's parameters() order, nn layouts), parameter count, state offsets of the three blocks}.  D <= 47, 2 <= T <=
 * IDFF_TRAIN_MD_MAX_FRAMES, 1 <= B <= IDFF_TRAIN_MD_MAX_BATCH (the reference: D 46, T 200, B 10).
 * Inputs of step s: d_idx [n_steps][B] int64 frame indices in [1, T) (clamped), d_t [n_steps][B], d_eps [n_steps][B][D] (NULL:
 * sigma_t * eps = 0); d_dataset [T][D].  opt: torch.optim.Adam's lr / betas / eps (weight_decay must be 0, max_grad_norm <= 0:
 * the loop neither decays nor clips).  Outputs (optional): d_loss [n_steps], d_grad_dump [parameter count] (gradient of the last
 * step, block layout). */
#define IDFF_TRAIN_MD_MAX_BATCH 16
#define IDFF_TRAIN_MD_MAX_FRAMES 224
/* Diagnostic: clock64 stamps of the last step of the last idff_train_md_steps launch (CTA 0), h_out32: step start, then the
 * arrival at / the release from each of the 8 cluster barriers, then step end (18 stamps). */
int idff_debug_train_md_stamps(int64_t* h_out32);
int64_t idff_train_md_state_floats(int32_t D, int32_t T);
int idff_train_md_offsets(int32_t D, int32_t T, int64_t* h_out13);
int idff_train_md_steps(float* d_state, const float* d_dataset, int32_t T, int32_t D, const int64_t* d_idx, const float* d_t,
                        const float* d_eps, int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode, const idff_adamw_t* opt,
                        int64_t steps_done, float* d_loss, float* d_grad_dump, void* stream);
/* The same loop with every batch RE-PAIRED before its path sample, as the `_dynamic` matcher's sample_plan does
 * (torchcfm/conditional_flow_matching_dynamic.py:268): x0 = dataset[d_row0], x1 = dataset[d_row1] (int64 [n_steps][B] each; NULL =
 * d_idx - 1 / d_idx, i.e. idff_train_md_steps), while the embedding sees index_sel = d_idx in its original order
 * (MD_simulation.py:135).  The rows come from idff_ot_assign_batch on the gathered frames + one index draw per step. */
int idff_train_md_steps_paired(float* d_state, const float* d_dataset, int32_t T, int32_t D, const int64_t* d_idx,
                               const int64_t* d_row0, const int64_t* d_row1, const float* d_t, const float* d_eps, int32_t n_steps,
                               int32_t B, float sigma, int32_t sigma_mode, const idff_adamw_t* opt, int64_t steps_done,
                               float* d_loss, float* d_grad_dump, void* stream);

/* ---- OTPlanSampler.get_map for the reference's use (torchcfm/optimal_transport.py:59-94: uniform marginals over
 * two minibatches of EQUAL size n, squared-Euclidean cost, exact plan from pot.emd -- POT's network simplex, un-vendored).  The
 * exact plan of uniform equal-size marginals is 1/n times a permutation matrix, so the plan is returned as the permutation:
 * d_perm [n] int64, plan[i][perm[i]] = 1/n; sample_plan (:120-142) then draws i uniformly and pairs x0[i] with x1[perm[i]]
 * (the gather inputs d_idx0 / d_idx1 of idff_path_sample).  d_x0, d_x1 [n][D] fp32; n <= 1024; d_cost (optional): the plan's
 * total cost sum_i |x0[i] - x1[perm[i]]|^2 in fp64.  fp64 throughout, one launch: shortest augmenting paths for n <= 64, an
 * epsilon-scaling auction above (eps down to max(c) 2^-40: the optimum unless another plan is within n max(c) 2^-40 of it). */
/* Diagnostic: on != 0 makes idff_ot_assign use the augmenting-path kernel for every size (default: n <= 64 only; above that an
 * epsilon-scaling auction, exact up to plans whose costs differ by less than n max(c) 2^-40). */
int idff_debug_ot_sap(int32_t on);
/* Diagnostic (calling thread): block size of the auction kernel, a multiple of 32 in [64, 512]; 0 = default (512 threads up to one
 * problem per SM, 128 (n <= 512) or 256 for larger batches).  The result does not depend on it. */
int idff_debug_ot_threads(int32_t threads);
int idff_ot_assign(const float* d_x0, const float* d_x1, int32_t n, int32_t D, int64_t* d_perm, double* d_cost, void* stream);
/* n_problems independent couplings in ONE launch, one CTA per problem: d_x0, d_x1 [n_problems][n][D] -> d_perm [n_problems][n],
 * d_cost [n_problems] (may be NULL); every problem gets exactly the result idff_ot_assign returns for it.  The solvers are
 * latency-bound one-CTA kernels, so the minibatches of the next S training steps -- they do not depend on the parameters -- are
 * coupled together and fill the machine (n = 256: 1024 problems in one launch cost what 5 single problems cost).  n_problems <= 2^20. */
int idff_ot_assign_batch(const float* d_x0, const float* d_x1, int32_t n_problems, int32_t n, int32_t D, int64_t* d_perm,
                         double* d_cost, void* stream);

/* ---- TrajectoryGenerator.generate_trajectories, timeseris_example/MD_simulation.py:197-213 -- n_frames
 * sequential solves  x1 = sdeint(SDE_cal(model, gamma2 = gamma2 / (1 + f), init_ind = f, dt, ...), x1, ts, dt)[-1],  f = 0 ..
 * n_frames - 1, each starting from the previous frame's result -- in ONE launch: one CTA owns 4 trajectories for the whole
 * chain, the network stays in shared memory.  p: variant IDFF_FIELD_MD with the BASE gamma2 in gamma[1] (divided by 1 + f per
 * frame here, as :201 does) and end_div = dt; t_idx is ignored (frame f uses embedding row f).  d_x_init [K][D];
 * d_I_k, d_I_k0 [n_frames][n_steps][K][D] Brownian increments per frame (n_steps = idff_sde_num_steps; d_I_k0 may be NULL for
 * euler); d_y_hat [n_frames][K][D] = traj[-1] of every frame.  Serves hidden width 128, D <= 63 (IDFF_EUNSUPPORTED otherwise).  It is
 * the latency path (the reference generates 5 trajectories): 592 trajectories run concurrently, more in waves; above a few
 * thousand trajectories one idff_sample_sde call per frame (the throughput kernel) is faster. */
/* Diagnostic: on != 0 makes idff_md_generate evaluate the layers with the fp32 FMA loop instead of mma.sync (3 x tf32). */
int idff_debug_md_chain_fma(int32_t on);
int idff_md_generate(const idff_weights_t* w, const idff_field_params_t* p, int32_t n_frames, const float* d_x_init,
                     const float* h_ts, int32_t n_ts, float dt, int32_t method, const float* d_I_k, const float* d_I_k0,
                     float* d_y_hat, int64_t K, void* stream);
/* The (sigma, gamma1, gamma2) sweep of TrajectoryGenerator.evaluate_model (MD_simulation.py:159-194) at one nfe / dt: n_sets
 * parameter sets (p_sets[i]: variant IDFF_FIELD_MD, same dim), each generating K trajectories, in ONE launch -- a generation uses
 * ceil(K / 4) CTAs (2 of 148 SMs for the reference's K = 5), so up to 74 sets run in the time of one.  Rows are set-major:
 * d_x_init [n_sets * K][D]; d_I_k, d_I_k0 [n_frames][n_steps][n_sets * K][D]; d_y_hat [n_frames][n_sets * K][D].  Set i's rows are
 * bit for bit what idff_md_generate returns for p_sets[i] on those rows.  n_sets <= 4096. */
int idff_md_generate_sweep(const idff_weights_t* w, const idff_field_params_t* p_sets, int32_t n_sets, int32_t n_frames,
                           const float* d_x_init, const float* h_ts, int32_t n_ts, float dt, int32_t method, const float* d_I_k,
                           const float* d_I_k0, float* d_y_hat, int64_t K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDFF_B200_H */
