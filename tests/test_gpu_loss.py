"""Fused MonoSDFLoss (csrc/loss.cu) against the same loss in torch ops (oracle/loss_torch.py forward_torch,
itself equal to the reference's loss on the golden training fixtures via port.monosdf_loss): the seven scalars and
the gradients with respect to every renderer output, for every loss variant, masked / unmasked rays and empty masks."""
import pytest
import torch

from monosdf_b200.model.loss import MonoSDFLoss
from oracle import loss_torch, port
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(n, S=98, seed=0, fg_frac=0.7, gt_frac=0.9, eik=True):
    g = torch.Generator().manual_seed(seed)
    out = {
        "rgb_values": torch.rand(n, 3, generator=g),
        "depth_values": torch.rand(n, 1, generator=g) * 3 + 0.2,
        "normal_map": torch.randn(n, 3, generator=g) * 0.7,
        "sdf": torch.randn(n, S, generator=g).abs() + 0.01,
    }
    cross = torch.rand(n, generator=g) < fg_frac                       # rays whose sdf row changes sign
    out["sdf"][cross, S // 2:] *= -1
    if eik:
        out["grad_theta"] = torch.randn(2 * n, 3, generator=g)
        out["grad_theta_nei"] = out["grad_theta"] + 0.05 * torch.randn(2 * n, 3, generator=g)
    gt = {"rgb": torch.rand(1, n, 3, generator=g), "depth": torch.rand(1, n, 1, generator=g) * 0.06 + 0.02,
          "normal": torch.randn(1, n, 3, generator=g), "mask": (torch.rand(1, n, 1, generator=g) < gt_frac).float()}
    return out, gt


VARIANTS = [dict(), dict(rgb_loss="torch.nn.MSELoss"), dict(if_gamma_loss=True), dict(if_scale_invariant_depth=False),
            dict(end_step=100)]


@pytest.mark.parametrize("kw", VARIANTS)
@pytest.mark.parametrize("n,fg,gtf,eik", [(1000, 0.7, 0.9, True), (37, 0.5, 1.0, True), (4096, 0.0, 1.0, True), (512, 1.0, 1.0, False)])
def test_fused_loss_matches_torch(kw, n, fg, gtf, eik):
    out, gt = _inputs(n, fg_frac=fg, gt_frac=gtf, eik=eik)
    res = {}
    for which in ("torch", "fused"):
        loss_fn = MonoSDFLoss(**kw)
        loss_fn.step = 7
        # the torch evaluation runs in float64: in fp32 the autograd gradient of the scale/shift-invariant depth term
        # carries ~3e-4 of cancellation noise (the terms through scale and shift sum to zero analytically)
        dt = torch.float64 if which == "torch" else torch.float32
        o = {k: v.clone().to(DEV, dt).requires_grad_(k != "sdf") for k, v in out.items()}
        g = {k: v.to(DEV, dt) for k, v in gt.items()}
        r = loss_torch.forward_torch(loss_fn, o, g, True) if which == "torch" else loss_fn(o, g, True)
        r["loss"].backward()
        res[which] = (r, {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in o.items() if k != "sdf"})
    for k in ("loss", "rgb_loss", "eikonal_loss", "smooth_loss", "depth_loss", "normal_l1", "normal_cos"):
        a, b = float(res["fused"][0][k]), float(res["torch"][0][k])
        assert a == pytest.approx(b, rel=2e-5, abs=1e-7), k
    for k in res["torch"][1]:
        a, b = res["fused"][1][k], res["torch"][1][k]
        if float(b.abs().max()) == 0.0:
            assert float(a.abs().max()) == 0.0, k
        else:
            assert rel_err(a, b) < 1e-4, (k, rel_err(a, b))


def test_fused_loss_equals_oracle_loss():
    out, gt = _inputs(800, seed=3)
    o = {k: v.to(DEV) for k, v in out.items()}
    g = {k: v.to(DEV) for k, v in gt.items()}
    a = MonoSDFLoss()(o, g, True)
    b = port.monosdf_loss(out, gt)
    for k in ("loss", "rgb_loss", "eikonal_loss", "smooth_loss", "depth_loss", "normal_l1", "normal_cos"):
        assert float(a[k]) == pytest.approx(float(b[k]), rel=2e-5, abs=1e-7), k
