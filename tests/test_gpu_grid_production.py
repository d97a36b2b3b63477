"""ImplicitNetworkGrid + HashEncoder at PRODUCTION geometry (kitchen_HDR_grids: 16 levels x 2 features, 2^19 entries per
level, 16 -> 2048, 2 x 256 MLP; reference confs/mi.conf:83-132, network.py:247-309, hashencoder/hashgrid.py:154-166) in
both precision modes, against the oracle (oracle/port.py, whose hash_encode restates hashencoder.cu:35-254; the CUDA hash
operators themselves are pinned to the reference's own kernels by tests/test_gpu_hashgrid_reference.py).

The table is drawn U(-0.5, 0.5) (seed 3) and every weight_v is perturbed: with the untouched geometric
init, layer 0 ignores every input column but xyz (network.py:218-237) and the hash features would not be tested.
Bars: fp32 mode 1e-4 (outputs) / 2e-4 (gradients, atomics); tensor-core mode 2e-2."""
import copy

import pytest
import torch

from monosdf_b200 import confs
from oracle import port
from tests.helpers import build_model, params_of, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"fp32": (1e-4, 2e-4), "bf16": (2e-2, 2e-2)}


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(beta=0.02):
    conf = copy.deepcopy(confs.KITCHEN_GRIDS)
    conf["rendering_network"]["per_image_code"] = False
    fx = {"conf": conf, "seed": 0, "beta": beta}
    model = build_model(fx, DEV)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        emb = model.implicit_network.encoding.embeddings
        emb.copy_((torch.rand(emb.shape, generator=g) - 0.5).to(DEV))
        for name, p in model.named_parameters():
            if name.endswith("weight_v"):
                p.add_((torch.randn(p.shape, generator=g) * 0.02).to(DEV))
    return model, conf


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_grid_field_forward_backward_production(precision):
    """sdf, grad_x sdf (gradient_sdf / get_sdf_vals: the kinds that keep the tensor-core engine) and the parameter + table
    gradients of random upstream adjoints, 4096 + 77 points, some outside the grid's [-1.1, 1.1]^3 domain."""
    model, conf = _model()
    model.set_precision(precision)
    cfg = port.cfg_from_conf(conf)
    n = 4096 + 77
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * 1.2
    w_grad = torch.randn(n, 3, generator=g)
    params = params_of(model, requires_grad=True)
    sdf_o = port.sdf_vals(params, cfg, x)
    grad_o = port.sdf_gradient(params, cfg, x)
    (grad_o * w_grad).sum().backward()
    inet = model.implicit_network
    with torch.no_grad():
        sdf = inet.get_sdf_vals(x.to(DEV))
    grad = inet.gradient_sdf(x.to(DEV))
    (grad * w_grad.to(DEV)).sum().backward()
    t_out, t_grad = TOL[precision]
    print("REPORT grid production %s: sdf %.3e grad %.3e" % (precision, rel_err(sdf, sdf_o), rel_err(grad, grad_o)))
    assert rel_err(sdf, sdf_o) < t_out
    if precision == "fp32":
        # grad_x at level 15 (resolution 2048): fp32 resolves the position inside a cell to 2048 * 2^-24 = 1.2e-4, so
        # two fp32 evaluations of d feature / d x only agree to a few 1e-4 unless every rounding matches (the CUDA
        # operators are pinned to the reference's kernels at 1e-6 on identical inputs by test_gpu_hashgrid_reference).
        # The bar is therefore set against the oracle in DOUBLE precision: our fp32 result must be as close to it as the
        # reference's fp32 arithmetic class (the oracle in fp32) is.
        p64 = {k: (v.detach().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in params.items()}
        grad_64 = port.sdf_gradient(p64, cfg, x.double())
        (grad_64 * w_grad.double()).sum().backward()
        e_ours, e_port = rel_err(grad, grad_64), rel_err(grad_o, grad_64)
        print("REPORT grid production fp32: grad vs fp64 oracle: ours %.3e, fp32 oracle %.3e" % (e_ours, e_port))
        assert e_ours < max(1e-4, 3.0 * e_port), (e_ours, e_port)
    else:
        assert rel_err(grad, grad_o) < t_out
    worst = 0.0
    for k, p in model.named_parameters():
        if not k.startswith("implicit_network.") or params[k].grad is None:
            continue
        assert p.grad is not None, k
        if precision == "fp32":       # same criterion as for grad_x: as close to the fp64 oracle as the fp32 oracle is
            e, e_port = rel_err(p.grad, p64[k].grad), rel_err(params[k].grad, p64[k].grad)
            assert e < max(3e-4, 3.0 * e_port), (k, e, e_port)      # the same error class as the fp32 oracle's
        else:
            e = rel_l2(p.grad, params[k].grad)
            assert e < t_grad, (k, e)
        worst = max(worst, e)
    print("REPORT grid production %s: worst parameter-gradient error %.3e (table included)" % (precision, worst))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_grid_train_step_production(precision):
    """one training step (64 rays x 98 samples = 6272 render points + eikonal points) on the oracle's sample positions
    (z_vals injected: the sampler's own parity is the bit-exact sampler test)."""
    model, conf = _model()
    model.train()
    model.rng = "reference"
    model.set_precision(precision)
    n = 64
    rays, gt = port.synthetic_rays(n, seed=1), port.synthetic_gt(n, seed=2)
    torch.manual_seed(9)
    out = model({k: v.to(DEV) for k, v in rays.items()}, torch.zeros(n, dtype=torch.long, device=DEV), if_pixel_input=True)
    loss = port.monosdf_loss(out, {k: v.to(DEV) for k, v in gt.items()})
    loss["loss"].backward()
    params = params_of(model, requires_grad=True)
    cfg = port.cfg_from_conf(conf)
    torch.manual_seed(9)
    out_o = port.model_forward(params, cfg, rays, torch.zeros(n, dtype=torch.long), if_pixel_input=True, training=True,
                               eik_points=model._last_eikonal_points.cpu(), z_vals=out["z_vals"].detach().cpu())
    loss_o = port.monosdf_loss(out_o, gt)
    loss_o["loss"].backward()
    t_out, t_grad = TOL[precision]
    t_out = max(t_out, 5e-4)      # composited outputs: 1 / beta amplification of the last-bit sdf differences (as in test_gpu_parity)
    for k in ["rgb_values", "depth_values", "normal_map", "sdf", "grad_theta"]:
        print("REPORT grid production train %s %s %.3e" % (precision, k, rel_err(out[k], out_o[k])))
        assert rel_err(out[k], out_o[k]) < t_out, (k, rel_err(out[k], out_o[k]))
    assert float(loss["loss"]) == pytest.approx(float(loss_o["loss"]), rel=2e-3 if precision == "fp32" else 2e-2)
    ours = torch.cat([p.grad.flatten().cpu() for k, p in model.named_parameters() if params[k].grad is not None]).double()
    ref = torch.cat([params[k].grad.flatten() for k, p in model.named_parameters() if params[k].grad is not None]).double()
    cos = float((ours * ref).sum() / (ours.norm() * ref.norm()))
    print("REPORT grid production train %s: e2e gradient cosine %.6f rel-l2 %.3e" % (precision, cos, float((ours - ref).norm() / ref.norm())))
    assert float((ours - ref).norm() / ref.norm()) < 2e-2
    tab = model.implicit_network.encoding.embeddings
    assert rel_l2(tab.grad, params["implicit_network.encoding.embeddings"].grad) < 2e-2
