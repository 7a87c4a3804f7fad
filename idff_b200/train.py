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
random draws passed in (used by the parity test against the reference's formula).  The MLP's own GEMMs stay in
torch: training the network is SURVEY.md section 8(f) row 3, not the sampler hot path."""
from __future__ import annotations

import torch

from . import losses
from .torchcfm import conditional_flow_matching as cfm


def checkerboard_device(n: int, device, generator=None) -> torch.Tensor:
    """torchcfm/utils.py:74-78 with the draws taken on the device (same distribution, not the numpy stream)."""
    u = torch.rand(n, device=device, generator=generator) * 4 - 2
    v = torch.rand(n, device=device, generator=generator) - torch.randint(0, 2, (n,), device=device, generator=generator) * 2
    v = v + torch.floor(u) % 2
    return torch.stack([u, v], 1) * 2


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


class GraphedFlowTrainer:
    """Captures `flow_train_step` (with on-device data and time draws) into a CUDA graph.

    trainer = GraphedFlowTrainer(model, batch_size=256); for _ in range(n): loss = trainer.step()
    `loss` is a device tensor that the next replay overwrites; read it with .item() only when needed."""

    def __init__(self, model: torch.nn.Module, batch_size: int = 256, sigma: float = 0.0, lr: float = 2e-4,
                 clip: float = 1.0, data=checkerboard_device, device=None, warmup: int = 3):
        self.model = model
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

    def _eager(self):
        x1 = self.data(self.B, self.device)
        x0 = torch.randn_like(x1)
        t = torch.rand(self.B, device=self.device)          # on the device: a host draw cannot be captured
        self.loss.copy_(flow_train_step(self.model, self.optimizer, self.FM, x0, x1, t=t, clip=self.clip))

    def step(self) -> torch.Tensor:
        self.graph.replay()
        self.steps += 1
        return self.loss
