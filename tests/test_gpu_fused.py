"""The fused sdf-only kernel (csrc/fused_mlp.cuh: the whole SDF network as one persistent tcgen05 kernel, activations
in shared / tensor memory) against the fp32 path of the same library (itself 1e-4 from the oracle) and against the
per-layer tensor-core sweep it replaces.  Tolerance of the tensor-core mode: 2e-2 (BASELINE.json north_star); the
measured distance is printed (pytest -s)."""
import pytest
import torch

from monosdf_b200 import confs
from tests.helpers import build_model, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_TOL = 2e-2


def _model(conf, seed=0, table_scale=None):
    fx = {"conf": conf, "seed": seed, "beta": 0.01}
    model = build_model(fx, DEV)
    if table_scale is not None:       # non-trivial hash content (U(-0.5, 0.5), seed 3)
        g = torch.Generator().manual_seed(3)
        with torch.no_grad():
            emb = model.implicit_network.encoding.embeddings
            emb.copy_(((torch.rand(emb.shape, generator=g) - 0.5) * 2 * table_scale).to(DEV))
    return model


def _points(n, seed, spread=1.3):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(n, 3, generator=g) * 2 - 1) * spread).to(DEV)


def _sdf(model, x, precision, fused):
    from monosdf_b200 import _lib
    model.set_precision(precision)
    _lib.lib().msdf_set_fused(1 if fused else 0)
    try:
        with torch.no_grad():
            return model.implicit_network.get_sdf_vals(x).clone()
    finally:
        _lib.lib().msdf_set_fused(1)


@pytest.mark.parametrize("n", [1, 127, 256, 257, 40000 + 77, 148 * 256 * 3 + 129])
def test_fused_sdf_mlp_conf(n):
    """scannet-MLP-shaped net (8 x 256, skip at 4, PE 6), ragged sizes: one row, less than a sub-tile, exactly a tile,
    one row into the next, several tiles per CTA."""
    model = _model(confs.SCANNET_MLP)
    x = _points(n, seed=n)
    ref = _sdf(model, x, "fp32", False)
    swp = _sdf(model, x, "bf16", False)
    fus = _sdf(model, x, "bf16", True)
    print("REPORT fused n=%d vs fp32 %.3e (per-layer sweep vs fp32 %.3e), fused vs sweep %.3e" %
          (n, rel_err(fus, ref), rel_err(swp, ref), rel_err(fus, swp)))
    assert torch.isfinite(fus).all()
    assert rel_err(fus, ref) < BF16_TOL
    assert rel_err(fus, swp) < 5e-3


def test_fused_sdf_far_points_and_clamp():
    """points far outside the unit sphere (sampler reaches z = 3.85): large PE arguments and the bounding-sphere clamp
    of ImplicitNetwork.get_sdf_vals (network.py:134-136)."""
    model = _model(confs.SCANNET_MLP)
    x = _points(30000, seed=5, spread=4.5)
    ref = _sdf(model, x, "fp32", False)
    fus = _sdf(model, x, "bf16", True)
    print("REPORT fused far points vs fp32 %.3e" % rel_err(fus, ref))
    assert rel_err(fus, ref) < BF16_TOL


@pytest.mark.parametrize("table_scale", [1e-4, 0.5])
def test_fused_sdf_grid_conf(table_scale):
    """kitchen_HDR_grids-shaped net at production geometry (16 x 2 hash grid, 2^19, 16 -> 2048, 2 x 256 MLP,
    mi.conf:87-103): 71-wide encoded input, no sphere clamp."""
    model = _model(confs.KITCHEN_GRIDS, table_scale=table_scale)
    x = _points(50000 + 3, seed=7, spread=1.2)
    ref = _sdf(model, x, "fp32", False)
    swp = _sdf(model, x, "bf16", False)
    fus = _sdf(model, x, "bf16", True)
    print("REPORT fused grid conf (table %.0e) vs fp32 %.3e (sweep %.3e), fused vs sweep %.3e" %
          (table_scale, rel_err(fus, ref), rel_err(swp, ref), rel_err(fus, swp)))
    assert rel_err(fus, ref) < BF16_TOL
    assert rel_err(fus, swp) < 5e-3


def test_fused_sdf_gridmlp_conf():
    """the fork's "MLP" confs: Grid_MLP=True with use_grid_feature=False (71-wide input with 32 zero columns, 8 x 256,
    skip at 4 with 185 outputs before it; mp_jh4fc5c5qoQ_undist_scannetMLP.conf:82-126)."""
    import copy
    conf = copy.deepcopy(confs.KITCHEN_GRIDS)
    conf["implicit_network"].update(dims=[256] * 8, skip_in=[4], use_grid_feature=False)
    model = _model(conf)
    x = _points(20000 + 11, seed=9)
    ref = _sdf(model, x, "fp32", False)
    fus = _sdf(model, x, "bf16", True)
    swp = _sdf(model, x, "bf16", False)
    print("REPORT fused grid-MLP conf vs fp32 %.3e, fused vs sweep %.3e" % (rel_err(fus, ref), rel_err(fus, swp)))
    assert rel_err(fus, ref) < BF16_TOL
    assert rel_err(fus, swp) < 5e-3


def test_fused_kernel_is_what_runs():
    """the sampler's sdf passes in tensor-core mode launch ONE field kernel (plus the weight pack), not a per-layer chain."""
    from monosdf_b200 import _lib
    model = _model(confs.SCANNET_MLP)
    x = _points(4096, seed=1)
    _sdf(model, x, "bf16", True)
    before = _lib.launch_count()
    _sdf(model, x, "bf16", True)
    fused_launches = _lib.launch_count() - before
    before = _lib.launch_count()
    _sdf(model, x, "bf16", False)
    sweep_launches = _lib.launch_count() - before
    assert fused_launches < sweep_launches and fused_launches <= 12, (fused_launches, sweep_launches)


def test_effective_weight_cache_follows_parameter_updates(golden):
    """no-grad calls reuse the weight-normed weights while nothing changed, and see every kind of update: in-place torch
    ops (version counters), load_state_dict, and the fused Adam kernel that writes the flat arena directly."""
    from monosdf_b200 import training
    from oracle import port
    fx = golden("mlp_small")
    model = build_model(fx, DEV).eval()
    x = _points(2000, seed=2, spread=0.6)
    with torch.no_grad():
        a = model.implicit_network.get_sdf_vals(x).clone()
        b = model.implicit_network.get_sdf_vals(x).clone()
        assert torch.equal(a, b)
        model.implicit_network.lin1.weight_g.mul_(1.5)                       # in-place update
        c = model.implicit_network.get_sdf_vals(x).clone()
    assert not torch.equal(a, c)
    arena, opt = training.build_optimizer(model)
    arena.zero_grad()
    for p in model.parameters():
        p.grad.fill_(1.0e-2)
    opt.step()                                                              # raw kernel on the arena
    with torch.no_grad():
        d = model.implicit_network.get_sdf_vals(x).clone()
    assert not torch.equal(c, d)
    model2 = build_model(fx, DEV).eval()
    model2.load_state_dict(model.state_dict(), strict=True)
    with torch.no_grad():
        e = model2.implicit_network.get_sdf_vals(x).clone()
    assert torch.equal(d, e)


@pytest.mark.parametrize("conf_name", ["mlp", "grid"])
def test_backward_sweep_engines_agree(conf_name):
    """The three implementations of the backward-side sweeps -- the reverse sweep as one chained launch (tc_chain.cuh,
    default), per-layer kernels with TMA-fed operands (tc_stream.cuh) and the round-1 per-layer engine (tc_gemm.cuh) --
    are the same arithmetic on the same 16-bit operands: one training step on identical rays and sample positions gives
    the same outputs and parameter gradients (to accumulation order)."""
    from monosdf_b200 import _lib
    from oracle import port
    conf = confs.SCANNET_MLP if conf_name == "mlp" else confs.KITCHEN_GRIDS
    model = _model(conf, table_scale=0.01 if conf_name == "grid" else None).train()
    model.set_precision("bf16")
    n = 2048 + 37          # several 128-row tiles per chunk and a ragged tail
    rays = {k: v.to(DEV) for k, v in port.synthetic_rays(n, seed=1).items()}
    gt = {k: v.to(DEV) for k, v in port.synthetic_gt(n, seed=2).items()}
    idx = torch.zeros(n, dtype=torch.long, device=DEV)

    def step(stream, chain):
        _lib.lib().msdf_set_sweeps(stream, chain)
        try:
            torch.manual_seed(11)
            out = model(rays, idx, if_pixel_input=True)
            loss = port.monosdf_loss({k: v for k, v in out.items()}, gt)["loss"]
            model.zero_grad()
            loss.backward()
            return ({k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)},
                    {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
        finally:
            _lib.lib().msdf_set_sweeps(1, 1)

    before = _lib.launch_count()
    out_c, g_c = step(1, 1)
    n_chain = _lib.launch_count() - before
    before = _lib.launch_count()
    out_s, g_s = step(1, 0)
    n_stream = _lib.launch_count() - before
    out_g, g_g = step(0, 0)
    assert n_chain < n_stream, (n_chain, n_stream)          # the chained sweep replaces one launch per layer
    for k in ("rgb_values", "depth_values", "normal_map", "grad_theta"):
        assert rel_err(out_c[k], out_s[k]) < 2e-3, k
        assert rel_err(out_c[k], out_g[k]) < 2e-3, k
    worst = 0.0
    for k in g_c:
        for other in (g_s, g_g):
            d = float((g_c[k].double() - other[k].double()).norm() / other[k].double().norm().clamp_min(1e-20))
            worst = max(worst, d)
            assert d < 5e-3, (k, d)
    print("REPORT backward sweep engines (%s conf): worst parameter-gradient distance between engines %.2e" % (conf_name, worst))


@pytest.mark.parametrize("n", [1, 127, 128, 129, 255, 257, 148 * 256 + 1])
def test_chained_reverse_sweep_ragged_sizes(n):
    """grad_x sdf (ImplicitNetwork.gradient_sdf, network.py:98-109) through the chained reverse sweep at point counts around
    the 128-row sub-tile / 256-row pair boundaries: against the fp32 path and against the per-layer engines."""
    from monosdf_b200 import _lib
    model = _model(confs.SCANNET_MLP)
    inet = model.implicit_network
    x = _points(n, seed=100 + n, spread=0.5)
    model.set_precision("fp32")
    with torch.no_grad():
        ref = inet.gradient_sdf(x).clone()
    model.set_precision("bf16")
    out = {}
    for name, (stream, chain) in {"chain": (1, 1), "stream": (1, 0), "gemm": (0, 0)}.items():
        _lib.lib().msdf_set_sweeps(stream, chain)
        try:
            with torch.no_grad():
                out[name] = inet.gradient_sdf(x).clone()
        finally:
            _lib.lib().msdf_set_sweeps(1, 1)
    assert torch.isfinite(out["chain"]).all()
    assert rel_err(out["chain"], ref) < BF16_TOL, rel_err(out["chain"], ref)
    assert rel_err(out["chain"], out["stream"]) < 1e-3
    assert rel_err(out["chain"], out["gemm"]) < 1e-3
