"""BASELINE.json's full sizes (65 536 rays / step on one B200): the sampler stays bit-exact against the C oracle, and the
full pipeline is checked through size-independent properties (the oracle's torch port would need minutes and tens of
GB at this size): sortedness and range of the samples, partition-of-unity of the compositing weights, bounded outputs,
invariance to how the batch is chunked, determinism, and finite gradients with the norm ratio bf16 : fp32 near 1."""
import pytest
import torch

from oracle import port
from oracle.sampler_oracle import OracleSampler
from tests.helpers import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda"
N = 65536


def _cuda(d):
    return {k: v.to(DEV) for k, v in d.items()}


@pytest.fixture(autouse=True)
def _release_device_memory():
    """These tests hold tens of GB (saved activations, workspaces): start each one from an empty allocator."""
    from monosdf_b200 import _lib
    yield
    _lib.saved_pool.clear()
    _lib._workspaces.clear()
    torch.cuda.empty_cache()


def test_sampler_bit_exact_at_65536_rays(golden):
    fx = golden("mlp_full")
    model = build_model(fx, DEV).eval()
    model.ray_sampler.rng = "reference"
    rays = port.synthetic_rays(N, seed=9)
    d, o = rays["ray_dirs"], rays["ray_cam_loc"]
    with model.implicit_network.cached_weights():
        z_gpu, _ = model.ray_sampler.get_z_vals(d.to(DEV), o.to(DEV), model)
        iters = model.ray_sampler.last_total_iters
        cfg = port.cfg_from_conf(fx["conf"])
        sc = cfg.sampler
        smp = OracleSampler(cfg.scene_bounding_sphere, sc.near, sc.N_samples, sc.N_samples_eval, sc.N_samples_extra, sc.eps,
                            sc.beta_iters, sc.max_total_iters, sc.add_tiny)
        trace = {}
        with torch.no_grad():
            z_cpu, _ = smp.get_z_vals(d, o, lambda p: model.implicit_network.get_sdf_vals(p.to(DEV)).cpu(),
                                      float(model.density.get_beta().detach().cpu()), False, trace)
    assert trace["total_iters"] == iters and iters >= 2
    assert torch.equal(z_gpu.cpu(), z_cpu)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_render_properties(golden, precision):
    fx = golden("mlp_full")
    model = build_model(fx, DEV).eval()
    model.set_precision(precision)
    rays = _cuda(port.synthetic_rays(N, seed=1))
    idx = torch.zeros(N, dtype=torch.long, device=DEV)
    with torch.no_grad():
        out = model(rays, idx, if_pixel_input=True)
        again = model(rays, idx, if_pixel_input=True)
    z, w = out["z_vals"], out["weights"]
    assert z.shape == (N, 98)
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and float(z.min()) >= 0.0 and float(z.max()) <= 2 * 1.1 * 1.75 + 1e-6
    assert bool((w >= 0).all()) and float(w.sum(-1).max()) <= 1.0 + 1e-4          # alpha compositing: partition of unity
    # rays start inside the closed init sphere: every ray hits the surface, so (almost) all of the weight is spent
    assert float((w.sum(-1) > 0.99).float().mean()) > 0.99
    assert float(out["rgb_values"].min()) >= 0.0 and float(out["rgb_values"].max()) <= 1.0 + 1e-5
    assert bool(torch.isfinite(out["depth_values"]).all()) and bool(torch.isfinite(out["normal_map"]).all())
    assert float(out["normal_map"].norm(dim=-1).max()) <= 1.0 + 1e-3              # convex combination of unit normals
    for k in ("rgb_values", "depth_values", "normal_map", "z_vals"):              # deterministic eval path
        assert torch.equal(out[k], again[k]), k
    # chunk invariance: the same rays rendered 4096 at a time (every chunk needs the same number of sampler rounds here)
    parts = []
    with torch.no_grad():
        for s in range(0, N, 4096):
            part = {k: v[s:s + 4096] for k, v in rays.items()}
            parts.append(model(part, idx[s:s + 4096], if_pixel_input=True))
    tol = 1e-5 if precision == "fp32" else 0.0      # bf16 GEMMs are tile-order independent; fp32 SIMT reductions too, but
    for k in ("rgb_values", "depth_values", "normal_map"):   # the split-K of nothing here changes with M: expect equality
        cat = torch.cat([p[k] for p in parts], 0)
        assert float((cat - out[k]).abs().max()) <= tol + 1e-6, k


def test_full_size_training_step_gradients(golden):
    """One 65 536-ray training step per precision mode: finite gradients for every parameter, saved-activation path in use
    (the step's activations fit in HBM), and the bf16 gradient close to the fp32 one in norm and direction."""
    from monosdf_b200 import _lib
    fx = golden("mlp_full")
    rays, gt = _cuda(port.synthetic_rays(N, seed=1)), _cuda(port.synthetic_gt(N, seed=2))
    idx = torch.zeros(N, dtype=torch.long, device=DEV)
    grads = {}
    for precision in ("fp32", "bf16"):
        model = build_model(fx, DEV).train()
        model.set_precision(precision)
        with torch.no_grad():
            model.density.beta.fill_(0.05)          # a beta at which the bf16 sdf error is small against beta
        torch.manual_seed(11)
        out = model(rays, idx, if_pixel_input=True)
        loss = port.monosdf_loss(out, gt)["loss"]
        loss.backward()
        assert bool(torch.isfinite(loss))
        g = []
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            assert bool(torch.isfinite(p.grad).all()), k
            g.append(p.grad.flatten().double())
        grads[precision] = torch.cat(g)
        del model, out, loss
        _lib.saved_pool.clear()
        torch.cuda.empty_cache()
    a, b = grads["bf16"], grads["fp32"]
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    assert cos > 0.99, cos
    assert abs(float(a.norm() / b.norm()) - 1.0) < 0.05


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_subsample_matches_oracle(golden, precision):
    """65 536-ray training forward at the bench's beta = 0.01: a 256-ray subsample of its outputs against the oracle on the
    SAME sample positions (z_vals injected, like tests/test_gpu_bf16.py) -- values, not just properties, at full size."""
    from tests.helpers import params_of, rel_err
    fx = golden("mlp_full")
    model = build_model(fx, DEV).train()
    model.set_precision(precision)
    with torch.no_grad():
        model.density.beta.fill_(0.01)
    rays = port.synthetic_rays(N, seed=1)
    idx = torch.zeros(N, dtype=torch.long, device=DEV)
    torch.manual_seed(11)
    with torch.no_grad():
        out = model(_cuda(rays), idx, if_pixel_input=True)
    pick = torch.arange(0, N, N // 256)[:256]
    sub = {k: v[pick] for k, v in rays.items()}
    params = params_of(model)
    cfg = port.cfg_from_conf(fx["conf"])
    out_o = port.model_forward(params, cfg, sub, torch.zeros(256, dtype=torch.long), if_pixel_input=True, training=False,
                               z_vals=out["z_vals"][pick.to(DEV)].cpu())      # (autograd.grad inside: no no_grad here)
    tol = 5e-4 if precision == "fp32" else 2e-2
    for k in ["sdf", "rgb_values", "depth_values", "normal_map"]:
        e = rel_err(out[k][pick.to(DEV)], out_o[k])
        print("REPORT full size %s subsample %s %.3e" % (precision, k, e))
        assert e < tol, (k, e)
    wa, wb = out["weights"][pick.to(DEV)].double().cpu(), out_o["weights"].double()
    l2 = float((wa - wb).norm() / wb.norm())
    print("REPORT full size %s subsample weights l2 %.3e max-norm %.3e" % (precision, l2, rel_err(wa, wb)))
    assert l2 < tol * (1 if precision == "bf16" else 4)
