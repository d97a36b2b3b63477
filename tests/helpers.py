"""Shared helpers of the test-suite (CPU and GPU)."""
import torch

from oracle import port
from oracle.ref_shim import to_conf


def build_model(fx, device="cpu"):
    """monosdf_b200 model for a golden fixture: constructor under the fixture's seed, beta set like make_golden."""
    from monosdf_b200.model.network import MonoSDFNetwork
    torch.manual_seed(fx["seed"])
    model = MonoSDFNetwork(to_conf(fx["conf"]), if_hdr=fx.get("if_hdr", False))
    with torch.no_grad():
        model.density.beta.fill_(fx["beta"])
    return model.to(device)


def params_of(model, requires_grad=False):
    out = {}
    for k, v in model.state_dict().items():
        t = v.detach().cpu().clone()
        if requires_grad and t.is_floating_point():
            t.requires_grad_(True)
        out[k] = t
    return out


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def frac_within(a, b, tol):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b).abs() <= tol * (1.0 + b.abs())).double().mean())


def oracle_forward(fx, params, rays, training, seed=None, eik_points=None, uv=False):
    cfg = port.cfg_from_conf(fx["conf"], fx.get("if_hdr", False))
    if seed is not None:
        torch.manual_seed(seed)
    n = rays["uv"].shape[1] if uv else rays["ray_dirs"].shape[0]
    idx = torch.zeros(1 if uv else n, dtype=torch.long)
    trace = {}
    out = port.model_forward(params, cfg, rays, idx, if_pixel_input=not uv, training=training, trace=trace, eik_points=eik_points)
    return out, trace
