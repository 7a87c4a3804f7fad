"""The 2-D IDFF training step (2D_examples/Autograd_example_order2.py:89-125, = order3 :90-121) as a replayable
CUDA graph.

    x1 = sample_checkerboard(B); x0 = randn_like(x1)
    t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, return_noise=True)     # idff_path_sample
    vt = base_model(cat[xt, t])                                                              # torch (cuBLAS) fwd/bwd
    loss = mean(((vt - FM.x1) / (1e-2 + 1 - t))**2)                                          # idff_loss (value + grad)
    loss.backward(); clip_grad_norm_(params, 1); AdamW(lr = 2e-4).step()

At the reference's batch size (256) this step is ~60 tiny kernels and purely launch-bound, so `GraphedFlowTrainer`
captures it once -- data draw, path-sample kernel, MLP forward/backward, loss kernel, clipping and a capturable AdamW
-- and replays it with one `cudaGraphLaunch` per step.  `flow_train_step` is the same step in eager form with the
random draws passed in (used by the parity test against the reference's formula).

`FusedFlowTrainer` goes one step further for the network the 2-D scripts train (MLP 3-256-256-256-2, flow loss): the
whole step -- path sample, forward, loss, backward, gradient-norm clipping, AdamW -- is ONE hand-written persistent
kernel (`idff_train_flow_steps`, csrc/idff_trainer.cu) that runs many consecutive steps per launch; no torch kernel
is involved and the model's parameters are views of the kernel's state buffer.  SURVEY.md section 8(f) row 3."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, losses
from .torchcfm import conditional_flow_matching as cfm


def checkerboard_device(n: int, device, generator=None) -> torch.Tensor:
    """torchcfm/utils.py:74-78 with the draws taken on the device (same distribution, not the numpy stream)."""
    u = torch.rand(n, device=device, generator=generator) * 4 - 2
    v = torch.rand(n, device=device, generator=generator) - torch.randint(0, 2, (n,), device=device, generator=generator) * 2
    v = v + torch.floor(u) % 2
    return torch.stack([u, v], 1) * 2


_CENTERS = {}


def eight_gaussians_device(n: int, device, generator=None) -> torch.Tensor:
    """sample_8gaussians (torchcfm/utils.py:12-33,85-86) with the draws taken on the device: 8 modes on a circle of radius 5,
    noise of covariance sqrt(0.1) I -- i.e. std 0.1**0.25, the reference passes sqrt(var) I as the covariance.  (The mode table is
    cached per device: a host-to-device copy cannot be captured into a CUDA graph.)"""
    key = str(torch.device(device))
    if key not in _CENTERS:
        r = 0.5 ** 0.5
        _CENTERS[key] = torch.tensor([(1, 0), (-1, 0), (0, 1), (0, -1), (r, r), (r, -r), (-r, r), (-r, -r)], device=device,
                                     dtype=torch.float32) * 5.0
    which = torch.randint(0, 8, (n,), device=device, generator=generator)
    return _CENTERS[key][which] + torch.randn(n, 2, device=device, generator=generator) * (0.1 ** 0.25)


def flow_train_step(model, optimizer, FM, x0, x1, t=None, clip: float = 1.0):
    """One reference training step on given samples; returns the loss (0-d device tensor).  t = None draws it like the
    reference does (torch.rand on the CPU generator)."""
    optimizer.zero_grad(set_to_none=False)
    t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, t=t, return_noise=True)
    vt = model(torch.cat([xt, t[:, None]], dim=-1))
    loss = losses.flow_loss(vt, FM.x1, t)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    optimizer.step()
    return loss.detach()


def scorehead_train_step(model, optimizer, FM, x0, x1, sigma0, t=None, clip: float = 1.0):
    """One training step of the order-1 script with the score head (example_order1_with_prediction_head.py:111-137):
    outs = model(cat[xt, xt, t]); vt, st = outs[:, :D], outs[:, D:]; loss = flow_loss(vt) + mean((st - logp)^2) with
    logp = log_prob_xt_given_x1_dimwise(xt, x1, x0, t, sigma0) (:26-51, :126); clip; AdamW.  The path sample and both loss
    terms (value + gradient, one pass each) are kernels; the MLP runs in torch.  Returns the loss (0-d device tensor)."""
    optimizer.zero_grad(set_to_none=False)
    t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, t=t, return_noise=True)     # :116
    outs = model(torch.cat([xt, xt, t[:, None]], dim=-1))                                         # :119
    D = x1.shape[1]
    vt, st = outs[:, :D].contiguous(), outs[:, D:].contiguous()                                   # :120-121
    loss = losses.flow_loss(vt, FM.x1, t) + losses.score_loss(st, xt, FM.x1, FM.x0, t, sigma0)    # :126-132
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    optimizer.step()
    return loss.detach()


class GraphedFlowTrainer:
    """Captures `flow_train_step` (with on-device data and time draws) into a CUDA graph; `score_sigma` not None captures
    `scorehead_train_step` instead (the order-1 script: MLP(dim=4, w=256, time_varying=True), AdamW lr 1e-3, sigma 0.2).

    trainer = GraphedFlowTrainer(model, batch_size=256); for _ in range(n): loss = trainer.step()
    `loss` is a device tensor that the next replay overwrites; read it with .item() only when needed."""

    def __init__(self, model: torch.nn.Module, batch_size: int = 256, sigma: float = 0.0, lr: float = 2e-4,
                 clip: float = 1.0, data=checkerboard_device, device=None, warmup: int = 3, score_sigma=None):
        self.model = model
        self.score_sigma = score_sigma
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("GraphedFlowTrainer needs a CUDA device: idff_b200 has no CPU path")
        self.B = int(batch_size)
        self.FM = cfm.ExactOptimalTransportConditionalFlowMatcher(sigma=sigma)       # order2.py:81
        self.optimizer = torch.optim.AdamW(model.parameters(), lr=lr, capturable=True)  # order2.py:80
        self.clip, self.data = clip, data
        self.loss = torch.zeros((), device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._eager()
        self.steps = 0
        self._params = list(model.parameters())

    def _eager(self):
        x1 = self.data(self.B, self.device)
        x0 = torch.randn_like(x1)
        t = torch.rand(self.B, device=self.device)          # on the device: a host draw cannot be captured
        if self.score_sigma is None:
            self.loss.copy_(flow_train_step(self.model, self.optimizer, self.FM, x0, x1, t=t, clip=self.clip))
        else:
            self.loss.copy_(scorehead_train_step(self.model, self.optimizer, self.FM, x0, x1, self.score_sigma, t=t,
                                                 clip=self.clip))

    def step(self) -> torch.Tensor:
        self.graph.replay()
        # a graph replay writes the parameters without touching their version counters: bump them, so that caches keyed on
        # _version (the sampler's packed weights, weights.packed_for) see the update
        torch.autograd.graph.increment_version(self._params)
        self.steps += 1
        return self.loss


def md_train_step(model, optimizer, FM, dataset, index_sel, t=None):
    """One PolyALA training step (TrajectoryGenerator.train_model, MD_simulation.py:126-144) on given frame indices:
    x0 = dataset[idx - 1], x1 = dataset[idx]; path sample (idff_path_sample); vt = model(cat[xt, t], idx) (torch);
    periodic loss (idff_loss kind 2, value + gradient in one pass); Adam.  Returns the loss (0-d device tensor)."""
    optimizer.zero_grad(set_to_none=False)
    x0, x1 = dataset[index_sel - 1], dataset[index_sel]                                    # :129-130
    t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, t=t, return_noise=True)   # :132
    vt = model(torch.cat([xt, t[:, None]], dim=-1), index_sel)                                 # :135
    loss = losses.md_periodic_loss(vt, x1)                                                      # :137-141
    loss.backward()
    optimizer.step()
    return loss.detach()


class GraphedMDTrainer:
    """`md_train_step` with on-device index / time draws captured as one CUDA graph (the reference's batch is 10:
    ~70 launch-bound kernels per step).  dataset: (T, D) CUDA tensor; model: MLP_Embedding(dim=D, w, time_embed=T,
    time_varying=True); optimizer: torch.optim.Adam with defaults (MD_simulation.py:119), sigma :121."""

    def __init__(self, model: torch.nn.Module, dataset: torch.Tensor, batch_size: int = 10, sigma: float = 0.5, warmup: int = 3):
        if not dataset.is_cuda:
            raise RuntimeError("GraphedMDTrainer needs a CUDA data set: idff_b200 has no CPU path")
        self.model, self.dataset, self.B = model, dataset, int(batch_size)
        self.device = dataset.device
        self.FM = cfm.ExactOptimalTransportConditionalFlowMatcher(sigma=sigma)
        self.optimizer = torch.optim.Adam(model.parameters(), capturable=True)
        self.loss = torch.zeros((), device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._eager()
        self.steps = 0
        self._params = list(model.parameters())

    def _eager(self):
        idx = torch.randint(1, self.dataset.shape[0], (self.B,), device=self.device)      # :128
        t = torch.rand(self.B, device=self.device)
        self.loss.copy_(md_train_step(self.model, self.optimizer, self.FM, self.dataset, idx, t=t))

    def step(self) -> torch.Tensor:
        self.graph.replay()
        # a graph replay writes the parameters without touching their version counters: bump them, so that caches keyed on
        # _version (the sampler's packed weights, weights.packed_for) see the update
        torch.autograd.graph.increment_version(self._params)
        self.steps += 1
        return self.loss


class FusedFlowTrainer:
    """The reference's 2-D training loops on the persistent kernel `idff_train_steps`: Autograd_example_order2.py:89-112
    (MLP(dim=2), flow loss) and, with `score_sigma` given, the order-1 script's step with the score head
    (example_order1_with_prediction_head.py:111-137: MLP(dim=4, w=256) fed cat[xt, xt, t], flow loss + score regression against
    log_prob_xt_given_x1_dimwise(.., score_sigma); the script uses data=eight_gaussians_device, lr=1e-3, sigma=0).

    model: torchcfm MLP(dim=2 | 4, w=256, time_varying=True) on a CUDA device.  Its parameters are re-pointed at the
    kernel's state buffer, so the module (forward, state_dict, PackedWeights.from_module) always sees the trained weights.
        tr = FusedFlowTrainer(model, batch_size=256)
        losses = tr.steps(x0, x1, t, eps)          # [S][B][2], [S][B][2], [S][B], [S][B][2] (eps optional) -> [S] losses
        losses = tr.run(n_steps)                   # draws checkerboard / normal / uniform samples on the device first
    """

    SHAPES = {0: [(256, 3), (256,), (256, 256), (256,), (256, 256), (256,), (2, 256), (2,)],      # MLP(dim=2): flow loss
              1: [(256, 5), (256,), (256, 256), (256,), (256, 256), (256,), (4, 256), (4,)]}      # MLP(dim=4): score head

    def __init__(self, model: torch.nn.Module, batch_size: int = 256, sigma: float = 0.0, lr: float = 2e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, clip: float = 1.0,
                 data=checkerboard_device, exact_ot_sigma: bool = True, score_sigma=None, ot_pairing: bool = False):
        L = _lib.lib()
        self.ot_pairing = bool(ot_pairing)   # draw(): couple every step's (x0, x1) minibatch by the exact OT plan, as the
        # `_dynamic` matcher's sample_plan does per step (conditional_flow_matching_dynamic.py:268) -- all steps in one launch
        params = [p for p in model.parameters()]
        shapes = [tuple(p.shape) for p in params]
        self.net = 0 if score_sigma is None else 1      # IDFF_TRAIN_NET_FLOW / IDFF_TRAIN_NET_SCOREHEAD
        if shapes != self.SHAPES[self.net]:
            raise _lib.IdffError(f"FusedFlowTrainer serves MLP(dim=2, w=256, time_varying=True) with the flow loss and MLP(dim=4, "
                                 f"w=256, time_varying=True) with score_sigma (the order-1 script's step), got {shapes}; "
                                 "use GraphedFlowTrainer for other networks")
        self.score_sigma = 0.0 if score_sigma is None else float(score_sigma)
        self.device = params[0].device
        if self.device.type != "cuda":
            raise _lib.IdffError("FusedFlowTrainer needs a CUDA device: idff_b200 has no CPU path")
        self.B = int(batch_size)
        off = (C.c_int64 * 12)()
        _lib.check(L.idff_train_offsets(self.net, off), "idff_train_offsets")
        self.offsets = [int(v) for v in off]
        self.n_params = self.offsets[8]
        self.state = torch.zeros(int(L.idff_train_state_floats(self.net)), device=self.device, dtype=torch.float32)
        base = self.offsets[9]
        with torch.no_grad():
            for p, o in zip(params, self.offsets[:8]):
                view = self.state[base + o: base + o + p.numel()].view(p.shape)
                view.copy_(p)
                p.data = view
        self.model = model
        self.opt = _lib.AdamW(lr, betas[0], betas[1], eps, weight_decay, clip)
        self.sigma, self.sigma_mode = float(sigma), 1 if exact_ot_sigma else 0   # ExactOT matcher: sigma*sqrt(t(1-t))
        self.data = data
        self.steps_done = 0
        self.sync_params()

    def state_dict(self):
        """model.state_dict() with every tensor cloned out of the kernel's state buffer -- what the reference saves when the
        loss improves (Autograd_example_order2.py:115-117).  (The module's own state_dict() returns views of one 13 MB buffer,
        which torch.save would write whole.)"""
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()}

    def sync_params(self):
        """Rebuild the kernel's streaming weight copies after the parameters were written from outside (load_state_dict,
        manual edits).  `steps` calls it by itself when a parameter's version counter moved since the last launch."""
        _lib.check(_lib.lib().idff_train_init(self.net, _lib.ptr(self.state), _lib.stream_of(self.state)), "idff_train_init")
        self._versions = [p._version for p in self.model.parameters()]

    def _block(self, name):   # parameter-block shaped views: "param", "exp_avg", "exp_avg_sq"
        base = self.offsets[{"param": 9, "exp_avg": 10, "exp_avg_sq": 11}[name]]
        return self.state[base: base + self.n_params]

    def named_block(self, name):
        """The 8 tensors (nn.Linear layout) of the parameter / exp_avg / exp_avg_sq block."""
        flat, shapes = self._block(name), self.SHAPES[self.net]
        out = []
        for o, sh in zip(self.offsets[:8], shapes):
            n = 1
            for v in sh:
                n *= v
            out.append(flat[o: o + n].view(sh))
        return out

    def steps(self, x0, x1, t, eps=None, return_gnorm: bool = False, grad_dump: torch.Tensor | None = None):
        for name, v in (("x0", x0), ("x1", x1), ("t", t)) + ((("eps", eps),) if eps is not None else ()):
            _lib.require_cuda(v, name)
            if not v.is_contiguous():
                raise _lib.IdffError(f"{name} must be contiguous")
        if x0.dim() == 2:
            x0, x1, t = x0[None], x1[None], t[None]
            eps = eps[None] if eps is not None else None
        S, B = int(t.shape[0]), int(t.shape[1])
        if tuple(x0.shape) != (S, B, 2) or tuple(x1.shape) != (S, B, 2) or (eps is not None and tuple(eps.shape) != (S, B, 2)):
            raise _lib.IdffError(f"expected x0, x1[, eps] of shape ({S}, {B}, 2)")
        if [p._version for p in self.model.parameters()] != self._versions:
            self.sync_params()                 # somebody wrote the parameters in place since the last launch
        loss = torch.empty(S, device=self.device, dtype=torch.float32)
        gnorm = torch.empty(S, device=self.device, dtype=torch.float32) if return_gnorm else None
        _lib.check(_lib.lib().idff_train_steps(
            self.net, _lib.ptr(self.state), _lib.ptr(x0), _lib.ptr(x1), _lib.ptr(t), _lib.ptr(eps), S, B, self.sigma, self.sigma_mode,
            self.score_sigma, C.byref(self.opt), self.steps_done, _lib.ptr(loss), _lib.ptr(gnorm), _lib.ptr(grad_dump),
            _lib.stream_of(self.state)), "idff_train_steps")
        self.steps_done += S
        torch.autograd.graph.increment_version(list(self.model.parameters()))   # the kernel wrote them in place
        self._versions = [p._version for p in self.model.parameters()]
        return (loss, gnorm) if return_gnorm else loss

    def draw(self, n_steps: int):
        """Device draws for n_steps steps: x1 ~ data, x0 = randn_like, t ~ U(0,1) (order2.py:92-95; t on the device)."""
        x1 = self.data(n_steps * self.B, self.device).view(n_steps, self.B, 2).contiguous()
        x0 = torch.randn_like(x1)
        if self.ot_pairing:
            from .torchcfm.optimal_transport import OTPlanSampler
            x0, x1 = OTPlanSampler().sample_plan_batch(x0, x1)
        t = torch.rand(n_steps, self.B, device=self.device)
        eps = torch.randn_like(x1) if self.sigma != 0.0 else None
        return x0, x1, t, eps

    def run(self, n_steps: int) -> torch.Tensor:
        x0, x1, t, eps = self.draw(n_steps)
        return self.steps(x0, x1, t, eps)


class FusedMDTrainer:
    """TrajectoryGenerator.train_model's loop (timeseris_example/MD_simulation.py:117-144) on `idff_train_md_steps`: every step --
    frame-pair gather, path sample, MLP_Embedding forward, periodic loss, backward, Adam -- inside one persistent kernel on a
    cluster of 8 CTAs that keeps the parameters and both Adam moments in shared memory (csrc/idff_trainer_md.cu); many steps per
    launch, no torch kernel involved.

    model: torchcfm MLP_Embedding(dim=D, w=128, time_embed=T, time_varying=True) on a CUDA device (D <= 47, T <= 224); its parameters
    are re-pointed at the kernel's state buffer.  dataset: (T, D) CUDA tensor.
        tr = FusedMDTrainer(model, dataset, batch_size=10, sigma=0.5)
        losses = tr.steps(idx, t, eps)          # [S][B] int64 in [1, T), [S][B], [S][B][D] -> [S] losses
        losses = tr.run(n_steps)                # draws idx / t / eps on the device first
    """

    def __init__(self, model: torch.nn.Module, dataset: torch.Tensor, batch_size: int = 10, sigma: float = 0.5, lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8, exact_ot_sigma: bool = True, ot_pairing: bool = False):
        L = _lib.lib()
        self.ot_pairing = bool(ot_pairing)   # run(): re-pair every batch by its exact OT plan first (the `_dynamic` matcher)
        _lib.require_cuda(dataset, "dataset")
        self.dataset = dataset.contiguous().float()
        self.T, self.D = int(dataset.shape[0]), int(dataset.shape[1])
        params = [p for p in model.parameters()]
        shapes = [tuple(p.shape) for p in params]
        D, T, w = self.D, self.T, 128
        want = [(T, w), (w, D + 1 + w), (w,), (w, w), (w,), (w, w), (w,), (D, w), (D,)]
        if shapes != want:
            raise _lib.IdffError(f"FusedMDTrainer serves MLP_Embedding(dim={D}, w=128, time_embed={T}, time_varying=True), got {shapes}; "
                                 "use GraphedMDTrainer for other networks")
        self.shapes = want
        self.device = params[0].device
        if self.device != self.dataset.device:
            raise _lib.IdffError("model and dataset live on different devices")
        self.B = int(batch_size)
        n = int(L.idff_train_md_state_floats(D, T))
        if n < 0:
            raise _lib.IdffError("idff_train_md_state_floats: " + L.idff_last_error().decode())
        off = (C.c_int64 * 13)()
        _lib.check(L.idff_train_md_offsets(D, T, off), "idff_train_md_offsets")
        self.offsets = [int(v) for v in off]
        self.n_params = self.offsets[9]
        self.state = torch.zeros(n, device=self.device, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, self.offsets[:9]):
                view = self.state[o: o + p.numel()].view(p.shape)
                view.copy_(p)
                p.data = view
        self.model = model
        self.opt = _lib.AdamW(lr, betas[0], betas[1], eps, 0.0, 0.0)     # torch.optim.Adam defaults: no decay; the loop does not clip
        self.sigma, self.sigma_mode = float(sigma), 1 if exact_ot_sigma else 0
        self.steps_done = 0

    def named_block(self, name):
        """The 9 tensors (module.parameters() order) of the "param" / "exp_avg" / "exp_avg_sq" block."""
        base = self.offsets[{"param": 10, "exp_avg": 11, "exp_avg_sq": 12}[name]]
        out = []
        for o, sh in zip(self.offsets[:9], self.shapes):
            n = 1
            for v in sh:
                n *= v
            out.append(self.state[base + o: base + o + n].view(sh))
        return out

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()}

    def steps(self, idx, t, eps=None, grad_dump: torch.Tensor | None = None, rows=None):
        """rows: optional (row0, row1) int64 (S, B) dataset rows of x0 / x1 when the batches were re-paired (the `_dynamic` matcher's
        sample_plan, conditional_flow_matching_dynamic.py:268); default idx - 1 / idx."""
        if not idx.is_cuda or idx.dtype != torch.int64:
            raise _lib.IdffError("idx must be a CUDA int64 tensor (torch.randint's dtype): idff_b200 has no CPU path")
        for name, v in (("t", t),) + ((("eps", eps),) if eps is not None else ()):
            _lib.require_cuda(v, name)
        for name, v in (("idx", idx), ("t", t)) + ((("eps", eps),) if eps is not None else ()):
            if not v.is_contiguous():
                raise _lib.IdffError(f"{name} must be contiguous")
        if idx.dim() == 1:
            idx, t = idx[None], t[None]
            eps = eps[None] if eps is not None else None
        S, B = int(idx.shape[0]), int(idx.shape[1])
        if tuple(t.shape) != (S, B) or (eps is not None and tuple(eps.shape) != (S, B, self.D)):
            raise _lib.IdffError(f"expected t of shape ({S}, {B}) and eps of shape ({S}, {B}, {self.D})")
        loss = torch.empty(S, device=self.device, dtype=torch.float32)
        r0 = r1 = None
        if rows is not None:
            r0, r1 = (v.contiguous().view(S, B) for v in rows)
            for v in (r0, r1):
                if not v.is_cuda or v.dtype != torch.int64:
                    raise _lib.IdffError("rows must be CUDA int64 tensors")
        _lib.check(_lib.lib().idff_train_md_steps_paired(
            _lib.ptr(self.state), _lib.ptr(self.dataset), self.T, self.D, _lib.ptr(idx), _lib.ptr(r0), _lib.ptr(r1), _lib.ptr(t),
            _lib.ptr(eps), S, B, self.sigma, self.sigma_mode, C.byref(self.opt), self.steps_done, _lib.ptr(loss), _lib.ptr(grad_dump),
            _lib.stream_of(self.state)), "idff_train_md_steps")
        self.steps_done += S
        torch.autograd.graph.increment_version(list(self.model.parameters()))   # the kernel wrote them in place
        return loss

    def draw(self, n_steps: int):
        """Device draws for n_steps steps: idx = randint(1, T) (MD_simulation.py:128), t ~ U(0, 1), eps ~ N(0, 1)."""
        idx = torch.randint(1, self.T, (n_steps, self.B), device=self.device)
        t = torch.rand(n_steps, self.B, device=self.device)
        eps = torch.randn(n_steps, self.B, self.D, device=self.device) if self.sigma != 0.0 else None
        return idx, t, eps

    def ot_rows(self, idx):
        """The `_dynamic` matcher's re-pairing for every step at once (conditional_flow_matching_dynamic.py:268 on x0 = dataset[idx - 1],
        x1 = dataset[idx]): one batched exact-OT launch, index pairs drawn with replacement per step -> (row0, row1) for `steps`."""
        from .torchcfm.optimal_transport import ot_assign_batch
        perm = ot_assign_batch(self.dataset[idx - 1], self.dataset[idx])               # (S, B)
        i = torch.randint(0, idx.shape[1], idx.shape, device=self.device)
        j = torch.gather(perm, 1, i)
        return torch.gather(idx, 1, i) - 1, torch.gather(idx, 1, j)

    def run(self, n_steps: int) -> torch.Tensor:
        idx, t, eps = self.draw(n_steps)
        return self.steps(idx, t, eps, rows=self.ot_rows(idx) if self.ot_pairing else None)
