"""The CUDA sampler against the REFERENCE's torch sampler run on the same GPU (model/ray_sampler.py:ErrorBoundSampler from
the staged reference files), fed by the same SDF network (ours, fp32 mode): exact-match / ulp / index statistics of the
sample positions ("report ulp / idx mismatch statistics vs the reference torch path").  The bit-exact
contract itself is against oracle/sampler_oracle.c (tests/test_gpu_parity.py::test_sampler_bit_exact): torch's CUDA
cumsum / sum association orders are its own, so agreement with the reference is statistical, and measured here."""
import json
import os

import pytest
import torch

from oracle import port, ref_shim
from tests.helpers import build_model

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="reference files not staged (baseline/_ref)")]
DEV = "cuda"


def _ulps(a, b):
    ia, ib = a.contiguous().view(torch.int32).long(), b.contiguous().view(torch.int32).long()
    return (ia - ib).abs()


@pytest.mark.parametrize("n,beta", [(4096, 0.01), (65536, 0.01), (4096, 0.001)])
def test_sampler_statistics_vs_reference_torch_sampler(golden, n, beta):
    net = ref_shim.load_reference()
    import model.ray_sampler as ref_rs
    fx = dict(golden("mlp_full"))
    fx["beta"] = beta
    model = build_model(fx, DEV).eval()
    sc = fx["conf"]["ray_sampler"]
    ref_sampler = ref_rs.ErrorBoundSampler(fx["conf"]["scene_bounding_sphere"], **sc)
    rays = port.synthetic_rays(n, seed=9)
    d, o = rays["ray_dirs"].to(DEV), rays["ray_cam_loc"].to(DEV)
    with torch.no_grad(), model.implicit_network.cached_weights():
        z_ours, _ = model.ray_sampler.get_z_vals(d, o, model)
        k_ours = model.ray_sampler.last_total_iters
        z_ref, _ = ref_sampler.get_z_vals(d, o, model)          # the reference's loop, our SDF network underneath
    assert z_ours.shape == z_ref.shape
    u = _ulps(z_ours, z_ref)
    close = (z_ours - z_ref).abs() <= 1e-4 * (1 + z_ref.abs())
    stats = {
        "rays": n, "beta": beta, "rounds_ours": int(k_ours), "samples": int(z_ref.numel()),
        "bit_equal_fraction": float((u == 0).float().mean()),
        "within_1_ulp": float((u <= 1).float().mean()), "within_4_ulp": float((u <= 4).float().mean()),
        "within_1e-4": float(close.float().mean()),
        "rays_with_any_sample_off_by_more_than_1e-4": float((~close).any(-1).float().mean()),
        "max_abs_diff": float((z_ours - z_ref).abs().max()),
    }
    print("REPORT sampler vs reference torch sampler: " + json.dumps(stats))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "sampler_vs_reference_%d_%g.json" % (n, beta)), "w") as f:
            json.dump(stats, f)
    assert stats["within_1e-4"] > 0.99
