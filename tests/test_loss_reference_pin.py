"""CPU: the loss oracle (oracle/port.py:monosdf_loss) and the host module's torch path (oracle/loss_torch.py) against
the REFERENCE's own model/loss.py:MonoSDFLoss on the reference's model outputs stored in the golden fixtures -- pins
SURVEY section 8 row f1 by execution.  (The fused CUDA loss is compared with forward_torch in tests/test_gpu_loss.py.)"""
import os

import pytest
import torch

from oracle import loss_torch, port

REF = "/root/reference/code"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
KEYS = ["loss", "rgb_loss", "eikonal_loss", "smooth_loss", "depth_loss", "normal_l1", "normal_cos"]


def _reference_loss():
    from oracle import ref_shim
    ref_shim.load_reference()
    from model.loss import MonoSDFLoss
    return MonoSDFLoss(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.05, smooth_weight=0.005, depth_weight=0.1,
                       normal_l1_weight=0.05, normal_cos_weight=0.05)


@pytest.mark.parametrize("case", ["mlp_small", "gridmlp_small", "mlp_full"])
@pytest.mark.parametrize("masked", ["none", "some", "all"])
def test_loss_oracle_and_host_module_match_reference_loss(golden, case, masked):
    fx = golden(case)
    n = fx["n_rays"]
    out = {k: v.clone() for k, v in fx["train"].items()}
    gt = port.synthetic_gt(n, seed=2)
    if masked == "some":
        gt["mask"][0, ::3] = 0.0
    elif masked == "all":                         # det == 0 in compute_scale_and_shift (loss.py:29-49)
        gt["mask"].zero_()
    ref = _reference_loss()(out, {k: v.clone() for k, v in gt.items()}, if_pixel_input=True)
    ours = port.monosdf_loss(out, gt)
    from monosdf_b200.model.loss import MonoSDFLoss
    host = loss_torch.forward_torch(MonoSDFLoss(), out, gt, if_pixel_input=True)
    for k in KEYS:
        r = float(ref[k])
        assert float(ours[k]) == pytest.approx(r, rel=1e-6, abs=1e-9), (k, "oracle")
        assert float(host[k]) == pytest.approx(r, rel=1e-6, abs=1e-9), (k, "host module")
