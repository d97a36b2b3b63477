"""GPU parity of the hash-grid encoder (operator boundary and inside the field) against the CPU restatement
oracle/port.py:hash_encode (hashencoder.cu:35-93,104-254 restated in differentiable torch ops).

The reference's own hash kernels are CUDA-only and cannot run in the build container; they are compiled there
(oracle/build_ref_hashencoder.py) and run on the GPU box by tests/test_gpu_hashgrid_reference.py, which pins
msdf_hash_encode_* to them by execution.  This file pins the restatement (used by the oracle's model_forward for Grid
nets) to the same CUDA path, and covers what the reference kernels cannot (C = 1, the fused row variants).
"""
import copy

import numpy as np
import pytest
import torch

from oracle import port
from oracle.make_golden import SMALL_CONF
from tests.helpers import params_of, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

GRID_CONF = copy.deepcopy(SMALL_CONF)
GRID_CONF["Grid_MLP"] = True
GRID_CONF["implicit_network"].update(dims=[64, 64], skip_in=[4], use_grid_feature=True, divide_factor=1.1, num_levels=8,
                                     level_dim=2, base_size=4, end_size=96, logmap=12)


def _encoder(levels=8, C=2, base=4, end=96, logmap=12, seed=3):
    from monosdf_b200.hashencoder import HashEncoder
    enc = HashEncoder(3, levels, C, 2, base, logmap, end).to(DEV)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) - 0.5).to(DEV))
    return enc


@pytest.mark.parametrize("C", [1, 2, 4])
def test_hash_encode_forward_and_dydx(C):
    enc = _encoder(C=C)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(3000, 3, generator=g) * 2.4 - 1.2          # some points outside [-1,1] -> zero features
    x[0] = torch.tensor([1.0, -1.0, 1.0])                     # exactly on the boundary
    offsets = enc.offsets.cpu().numpy()
    emb = enc.embeddings.detach().cpu()
    x01 = (x + 1) / 2
    feat_o, dydx_o = port.hash_encode(x01, emb, offsets, enc.per_level_scale, enc.base_resolution, want_dy_dx=True)
    xin = x.to(DEV).requires_grad_(True)
    out = enc(xin)
    assert rel_err(out, feat_o) < 1e-5
    # dy_dx through the operator: d out[:, j] / d x  (chain factor 1/2 of the [-1,1] -> [0,1] map)
    for j in [0, 3, out.shape[1] - 1]:
        (gx,) = torch.autograd.grad(out[:, j].sum(), xin, retain_graph=True)
        l, c = divmod(j, C)
        assert rel_err(gx, dydx_o[:, l, :, c] * 0.5) < 1e-5


def test_hash_encode_backward_and_second_backward():
    enc = _encoder()
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2000, 3, generator=g) * 2.2 - 1.1
    w = torch.randn(2000, enc.output_dim, generator=g)
    v = torch.randn(2000, 3, generator=g)
    offsets = enc.offsets.cpu().numpy()
    # oracle: first-order table gradient, and the eikonal-style double backward  d/d emb < d out . w / dx , v >
    emb = enc.embeddings.detach().cpu().clone().requires_grad_(True)
    xo = ((x + 1) / 2).requires_grad_(True)
    out_o = port.hash_encode(xo, emb, offsets, enc.per_level_scale, enc.base_resolution)
    (g_emb_o,) = torch.autograd.grad((out_o * w).sum(), emb, retain_graph=True)
    (gx_o,) = torch.autograd.grad((out_o * w).sum(), xo, create_graph=True)
    (g2_emb_o,) = torch.autograd.grad((gx_o * v).sum(), emb)
    # CUDA operator
    xin = x.to(DEV).requires_grad_(True)
    out = enc(xin)
    (g_emb,) = torch.autograd.grad((out * w.to(DEV)).sum(), enc.embeddings, retain_graph=True)
    (gx,) = torch.autograd.grad((out * w.to(DEV)).sum(), xin, create_graph=True)
    (g2_emb,) = torch.autograd.grad((gx * (2.0 * v).to(DEV)).sum(), enc.embeddings)   # x01 = (x+1)/2
    assert rel_err(g_emb, g_emb_o) < 1e-4
    assert rel_err(gx * 2.0, gx_o) < 1e-4
    assert rel_err(g2_emb, g2_emb_o) < 1e-4


def test_hash_levels_follow_reference_geometry():
    """Offsets / dense-vs-hashed levels of the kitchen_HDR_grids-shaped encoder (16 x 2, 2^19, 16 -> 2048)."""
    from monosdf_b200.hashencoder.hashgrid import level_offsets
    offs_o, pls = port.hash_offsets(16, 16, 2048, 19)
    offs = level_offsets(16, 16, pls, 19)
    assert np.array_equal(offs, offs_o)
    assert int(offs[-1]) == 6098120 or int(offs[-1]) > 6.0e6


def _grid_model():
    from monosdf_b200.model.network import MonoSDFNetwork
    from oracle.ref_shim import to_conf
    torch.manual_seed(0)
    model = MonoSDFNetwork(to_conf(GRID_CONF)).to(DEV)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        e = model.implicit_network.encoding.embeddings
        e.copy_(((torch.rand(e.shape, generator=g) - 0.5) * 0.2).to(DEV))
        model.density.beta.fill_(0.02)
    return model


def test_grid_field_forward_backward_matches_oracle():
    model = _grid_model()
    cfg = port.cfg_from_conf(GRID_CONF)
    n = 900
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * 1.25
    F = cfg.feature_vector_size
    w_sdf, w_feat, w_grad = torch.randn(n, 1, generator=g), torch.randn(n, F, generator=g) * 0.1, torch.randn(n, 3, generator=g)
    params = params_of(model, requires_grad=True)
    sdf_o, feat_o, grad_o = port.sdf_outputs(params, cfg, x)
    ((sdf_o * w_sdf).sum() + (feat_o * w_feat).sum() + (grad_o * w_grad).sum()).backward()
    sdf, feat, grad = model.implicit_network.get_outputs(x.to(DEV))
    assert rel_err(sdf, sdf_o) < 1e-4
    assert rel_err(feat, feat_o) < 1e-4
    assert rel_err(grad, grad_o) < 1e-4
    ((sdf * w_sdf.to(DEV)).sum() + (feat * w_feat.to(DEV)).sum() + (grad * w_grad.to(DEV)).sum()).backward()
    for k, p in model.named_parameters():
        if not k.startswith("implicit_network."):
            continue
        assert p.grad is not None, k
        assert rel_err(p.grad, params[k].grad) < 2e-4, k


def test_grid_model_train_step_matches_oracle():
    model = _grid_model().train()
    model.rng = "reference"
    n = 40
    rays = port.synthetic_rays(n, seed=1)
    gt = port.synthetic_gt(n, seed=2)
    torch.manual_seed(9)
    out = model({k: v.to(DEV) for k, v in rays.items()}, torch.zeros(n, dtype=torch.long, device=DEV), if_pixel_input=True)
    loss = port.monosdf_loss(out, {k: v.to(DEV) for k, v in gt.items()})
    loss["loss"].backward()
    params = params_of(model, requires_grad=True)
    cfg = port.cfg_from_conf(GRID_CONF)
    torch.manual_seed(9)
    out_o = port.model_forward(params, cfg, rays, torch.zeros(n, dtype=torch.long), if_pixel_input=True, training=True,
                               eik_points=model._last_eikonal_points.cpu())
    loss_o = port.monosdf_loss(out_o, gt)
    loss_o["loss"].backward()
    assert float(loss["loss"]) == pytest.approx(float(loss_o["loss"]), rel=2e-3)
    for k, p in model.named_parameters():
        assert p.grad is not None and params[k].grad is not None, k
        assert rel_err(p.grad, params[k].grad) < 2e-2, k
