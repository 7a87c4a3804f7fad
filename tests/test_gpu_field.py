"""-m gpu: one evaluation of the IDFF field (forward(t, x) / SDE_cal.f) and the plain MLP forward against the golden
vectors produced by the reference's own nested-autograd code.  Tolerance: 1e-5 * max|f| per evaluation."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_common import assert_field_close, assert_field_vs_refs, packed
from idff_b200 import _lib, fields
from idff_b200.torchcfm.models.models import MLP, MLP_Embedding
from make_golden_consts import FIELD_TIMES, GAMMAS2, GAMMAS3

pytestmark = pytest.mark.gpu

IMPLS = [_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_TENSOR, _lib.IDFF_IMPL_TENSOR16, _lib.IDFF_IMPL_LAYERED,
         _lib.IDFF_IMPL_AUTO]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("which", ["checkerboard", "8gaussians"])
def test_field2d_order2_order3(cuda_device, which, impl):
    g = load_golden("field2d")
    x = torch.tensor(g["x"]).to(cuda_device)
    pw = packed(which, cuda_device)
    worst = 0.0
    for gi, gam in enumerate(GAMMAS2):
        m = fields.CustomNetModel(pw, 0.2, *gam)
        m.impl = impl
        for ti, tv in enumerate(FIELD_TIMES):
            out = m(torch.tensor(tv, dtype=torch.float32), x).cpu().numpy()
            worst = max(worst, assert_field_vs_refs(out, g[f"{which}_o2_g{gi}"][ti], g[f"{which}_o2_g{gi}_f64"][ti],
                                                    1e-5, f"o2 g{gi} t={tv}"))
    for gi, gam in enumerate(GAMMAS3):
        m = fields.CustomNetModelOrder3(pw, 0.2, *gam)
        m.impl = impl
        if impl == _lib.IDFF_IMPL_TENSOR and gam[2] != 0:      # the 3xTF32 tcgen05 kernel carries two jets: no silent fallback
            with pytest.raises(_lib.IdffError):
                m(torch.tensor(0.5), x)
            continue
        for ti, tv in enumerate(FIELD_TIMES):
            out = m(torch.tensor(tv, dtype=torch.float32), x).cpu().numpy()
            worst = max(worst, assert_field_vs_refs(out, g[f"{which}_o3_g{gi}"][ti], g[f"{which}_o3_g{gi}_f64"][ti],
                                                    1e-5, f"o3 g{gi} t={tv}"))
    print(f"[{which} impl={impl}] worst per-eval error / max|f| = {worst:.2e}")


def test_mlp_forward_and_module_surface(cuda_device):
    g = load_golden("field2d")
    w = load_golden("weights_checkerboard")
    net = MLP(dim=2, w=256, time_varying=True)
    sd = {f"net.{2 * i}.weight": torch.tensor(w[f"W{i}"]) for i in range(4)}
    sd.update({f"net.{2 * i}.bias": torch.tensor(w[f"b{i}"]) for i in range(4)})
    net.load_state_dict(sd)                      # the reference's state_dict keys
    net = net.to(cuda_device)
    x = torch.tensor(g["x"]).to(cuda_device)
    inp = torch.cat([x, torch.full((x.shape[0], 1), 0.37, device=cuda_device)], -1)
    before = _lib.launch_count()
    with torch.no_grad():
        out = net(inp)
    assert _lib.launch_count() == before + 1     # served by the CUDA kernel, not by torch
    assert_field_close(out.cpu().numpy(), g["checkerboard_mlp"], 2e-6, "mlp forward")
    # under autograd the module is an ordinary differentiable torch network (training side)
    out2 = net(inp)
    assert out2.requires_grad
    assert_field_close(out2.detach().cpu().numpy(), g["checkerboard_mlp"], 1e-5, "mlp torch path")
    # the field wrapper accepts the module as base_model, like the reference's CustomNetModel(base_model, ...)
    m = fields.CustomNetModel(net, sigma=0.2, gamma1=0.2, gamma2=0.2)
    out3 = m(torch.tensor(0.37), x).cpu().numpy()
    assert_field_close(out3, g["checkerboard_o2_g2"][3], 1e-5, "module-backed field")
    # in-place weight update invalidates the packed copy
    with torch.no_grad():
        net.net[6].bias.add_(1.0)
    out4 = m(torch.tensor(0.0), x).cpu().numpy()
    assert_field_close(out4, g["checkerboard_o2_g2"][0] + 1.0, 1e-5, "repacked after update")


def test_scorehead_and_cfm_fields(cuda_device):
    g = load_golden("scorehead")
    x = torch.tensor(g["x"]).to(cuda_device)
    pw = packed("scorehead", cuda_device)
    for gi, gam in enumerate(g["gammas"]):
        m = fields.CustomNetModelScoreHead(pw, 0.2, float(gam[0]), float(gam[1]))
        for ti, tv in enumerate(FIELD_TIMES):
            out = m(torch.tensor(tv, dtype=torch.float32), x).cpu().numpy()
            assert_field_close(out, g[f"field_g{gi}"][ti], 1e-5, f"scorehead g{gi} t={tv}")
    # the order-1 script's own network (w = 256): generic fp32 kernel and the tcgen05 kernel k_tc16<score head>, against the
    # reference's fp32 and fp64 results
    g = load_golden("scorehead256")
    pw = packed("scorehead256", cuda_device)
    for gi, gam in enumerate(g["gammas"]):
        for impl in (_lib.IDFF_IMPL_AUTO, _lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_TENSOR16):
            m = fields.CustomNetModelScoreHead(pw, 0.2, float(gam[0]), float(gam[1]))
            m.impl = impl
            for ti, tv in enumerate(FIELD_TIMES):
                out = m(torch.tensor(tv, dtype=torch.float32), x).cpu().numpy()
                assert_field_vs_refs(out, g[f"field_g{gi}"][ti], g[f"field_g{gi}_f64"][ti], 1e-5, f"scorehead256 g{gi} t={tv} impl={impl}")
    c = fields.CFMModel(packed("checkerboard", cuda_device))
    with pytest.raises(_lib.IdffError):
        c(torch.tensor(0.5), x)                  # no x0 captured yet
    f0 = c(torch.tensor(0.0), x)
    f1 = c(torch.tensor(0.5), x + 0.25)
    m2 = fields.CustomNetModel(packed("checkerboard", cuda_device), 0.2, 0.5, 0.0)   # c1 = c2 = 0: f(0) = pred - x
    assert torch.equal(f0, m2(torch.tensor(0.0), x))
    assert f1.shape == x.shape and torch.isfinite(f1).all()


@pytest.mark.parametrize("impl", [_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_TENSOR16, _lib.IDFF_IMPL_AUTO])
def test_md_field_with_folded_embedding(cuda_device, impl):
    g = load_golden("md")
    y = torch.tensor(g["y"]).to(cuda_device)
    pw = packed("polyala", cuda_device)
    for ii in g["init_inds"]:
        sde = fields.SDE_cal(pw, input_dim=46, sigma=0.5, init_ind=int(ii), max_length=200, dt=0.3, gamma1=0.2,
                             gamma2=5.0 / (1 + int(ii)))
        sde.impl = impl
        assert sde.noise_type == "diagonal" and sde.sde_type == "ito"
        for ti, tv in enumerate((0.0, 0.3, 0.6, 1.0)):
            t = torch.tensor(tv)
            f = sde.f(t, y).cpu().numpy()
            # folding the embedding into the bias re-associates a 175-term fp32 dot product: 3e-5 abs on O(100)
            # outputs on top of the 1e-5 relative budget
            assert_field_close(f, g[f"f_i{ii}"][ti], 2e-5, f"SDE_cal.f init_ind={ii} t={tv}")
            np.testing.assert_array_equal(sde.g(t.to(cuda_device), y).cpu().numpy(), g[f"g_i{ii}"][ti])
    with pytest.raises(_lib.IdffError):
        fields.SDE_cal(pw, 46, sigma=0.5, init_ind=200).f(torch.tensor(0.5), y)    # outside Embedding(200, .)


def test_mlp_embedding_module(cuda_device):
    w = load_golden("weights_polyala")
    net = MLP_Embedding(dim=46, w=128, time_embed=200, time_varying=True)
    sd = {f"net.{2 * i}.weight": torch.tensor(w[f"W{i}"]) for i in range(4)}
    sd.update({f"net.{2 * i}.bias": torch.tensor(w[f"b{i}"]) for i in range(4)})
    sd["time_index_embedding.weight"] = torch.tensor(w["emb"])
    net.load_state_dict(sd)
    net = net.to(cuda_device)
    y = torch.randn(33, 46, device=cuda_device) * 50
    inp = torch.cat([y, torch.full((33, 1), 0.4, device=cuda_device)], -1)
    ti = torch.full((33, 1), 7.0, device=cuda_device)
    ref = net(inp, ti)                                      # autograd on: torch path
    with torch.no_grad():
        before = _lib.launch_count()
        out = net(inp, ti)                                  # kernel path with the folded bias of row 7
        assert _lib.launch_count() == before + 1
    assert_field_close(out.cpu().numpy(), ref.detach().cpu().numpy(), 2e-6, "MLP_Embedding kernel vs torch")
    # against the reference's own module: outputs of torchcfm/models/models.py:23-42 on the shipped PolyALA
    # checkpoint, written by oracle/make_golden.py.  Folding the embedding into the layer-0 bias re-associates a 175-term
    # fp32 dot product: 3e-5 abs on O(100) outputs
    g = load_golden("mlp_modules")
    yg = torch.tensor(g["y"]).to(cuda_device)
    with torch.no_grad():
        for tidx in (0, 7, 199):
            for tv in (0.0, 0.4, 1.0):
                inp = torch.cat([yg, torch.full((33, 1), tv, device=cuda_device)], -1)
                out = net(inp, torch.full((33, 1), float(tidx), device=cuda_device))
                assert_field_close(out.cpu().numpy(), g[f"emb_ti{tidx}_t{tv}"], 2e-6, f"MLP_Embedding vs reference module ti={tidx} t={tv}")


def test_field_shape_errors(cuda_device):
    pw = packed("checkerboard", cuda_device)
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    with pytest.raises(_lib.IdffError):
        m(torch.tensor(0.5), torch.randn(8, 3, device=cuda_device))      # network is 3->2, x has 3 columns
    with pytest.raises(_lib.IdffError):
        m(torch.tensor(0.5), torch.randn(8, 2))                          # CPU tensor: no CPU path
    assert m(torch.tensor(0.5), torch.empty(0, 2, device=cuda_device)).shape == (0, 2)
    out1 = m(torch.tensor(0.5), torch.ones(1, 2, device=cuda_device))
    out9 = m(torch.tensor(0.5), torch.ones(9, 2, device=cuda_device))
    assert torch.equal(out9[:1], out1) and torch.equal(out9[8:], out1)   # rows are independent
