"""The drop-in really drops in: the reference's OWN glue code drives monosdf_b200's model.

Uses the reference files oracle/stage_reference.py stages under baseline/_ref/ (or /root/reference in the build
container):
  * utils/general.py:get_class resolves the conf string `train.model_class = monosdf_b200.model.network.MonoSDFNetwork`
    (the plug-in mechanism, monosdf_train.py:199, general.py:14-20);
  * the module is wrapped in DistributedDataParallel(device_ids=[rank], broadcast_buffers=False,
    find_unused_parameters=True) exactly as monosdf_train.py:228-229 does (NCCL process group of one rank);
  * the reference's model/loss.py:MonoSDFLoss consumes its live outputs, loss.backward() runs through DDP's hooks, and
    torch.optim.Adam with the trainer's parameter groups (:210-221) + ExponentialLR (:223-226) step it;
  * a checkpoint written by the REFERENCE model in the trainer's format (monosdf_train.py:277-281) loads with
    strict=True the way eval.py:55-65 loads it ('module.' prefix stripped), and both models render the same image.
"""
import os

import pytest
import torch

from oracle import port, ref_shim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="reference files not staged (baseline/_ref)")]
DEV = "cuda"
CLASS = "monosdf_b200.model.network.MonoSDFNetwork"


def _cuda(d):
    return {k: v.to(DEV) for k, v in d.items()}


def _reference():
    net = ref_shim.load_reference()
    import utils.general as ref_utils
    from model.loss import MonoSDFLoss as RefLoss
    return net, ref_utils, RefLoss


def test_get_class_ddp_reference_loss_adam():
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel
    net, ref_utils, RefLoss = _reference()
    Model = ref_utils.get_class(CLASS)                                  # monosdf_train.py:199
    from monosdf_b200.model.network import MonoSDFNetwork
    assert Model is MonoSDFNetwork
    conf = ref_shim.to_conf(ref_shim.MLP_CONF)
    torch.manual_seed(0)
    model = Model(conf=conf, if_hdr=False)
    assert model.Grid_MLP is False                                      # :200
    model.cuda()                                                        # :201-202
    with torch.no_grad():
        model.density.beta.fill_(0.05)
    loss_fn = RefLoss(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.05, smooth_weight=0.005, depth_weight=0.1,
                      normal_l1_weight=0.05, normal_cos_weight=0.05)
    opt = torch.optim.Adam(model.parameters(), lr=5.0e-4)              # :221
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, 0.1 ** (1.0 / 1000))
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29561", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    try:
        ddp = DistributedDataParallel(model, device_ids=[0], broadcast_buffers=False, find_unused_parameters=True)   # :229
        n = 512
        rays, gt = _cuda(port.synthetic_rays(n, seed=1)), _cuda(port.synthetic_gt(n, seed=2))
        idx = torch.zeros(n, dtype=torch.long, device=DEV)
        before = {k: v.detach().clone() for k, v in model.state_dict().items()}
        losses = []
        import contextlib
        for it in range(3):                                             # the loop body of monosdf_train.py:425-432,480
            opt.zero_grad()
            out = ddp(rays, idx, if_pixel_input=True)
            with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):
                loss_out = loss_fn(out, gt, if_pixel_input=True)
            loss_out["loss"].backward()
            opt.step()
            sched.step()
            losses.append(float(loss_out["loss"]))
            assert float(ddp.module.density.get_beta()) > 0              # the trainer logs it (:454)
        assert all(torch.isfinite(torch.tensor(losses))), losses
        moved = [k for k, v in model.state_dict().items() if v.is_floating_point() and not torch.equal(v, before[k])]
        # every learnable tensor received a gradient through DDP's hooks and was stepped (the per-image code table of the
        # colour net does not exist in this conf)
        assert len(moved) == len([1 for _ in model.parameters()]), (len(moved), losses)
        # the reference loss on our outputs equals our fused loss on the same outputs
        from monosdf_b200.model.loss import MonoSDFLoss
        with torch.no_grad():
            out = ddp(rays, idx, if_pixel_input=True)
            with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):
                a = loss_fn(out, gt, if_pixel_input=True)["loss"]
            b = MonoSDFLoss()(out, gt, if_pixel_input=True)["loss"]
        assert float(a) == pytest.approx(float(b), rel=1e-5)
        print("REPORT drop-in: 3 DDP steps with the reference's MonoSDFLoss + torch.optim.Adam, losses %s" % losses)
    finally:
        if created:
            dist.destroy_process_group()


def test_reference_checkpoint_loads_strict_and_renders_the_same(tmp_path):
    """the reference model (its own torch CUDA path, on this GPU) writes a checkpoint; ours loads it strict=True and the two
    render the same 1024 rays in eval mode."""
    net, ref_utils, _ = _reference()
    conf = ref_shim.to_conf(ref_shim.MLP_CONF)
    torch.manual_seed(7)
    ref_model = net.MonoSDFNetwork(conf=conf).cuda()
    g = torch.Generator().manual_seed(8)
    with torch.no_grad():
        ref_model.density.beta.fill_(0.03)
        for name, p in ref_model.named_parameters():     # not the init: a "trained-looking" state
            if name.endswith("weight_v"):
                p.add_((torch.randn(p.shape, generator=g) * 0.01).to(DEV))
    path = str(tmp_path / "latest.pth")
    sd = {"module." + k: v for k, v in ref_model.state_dict().items()}          # as saved from a DDP-wrapped model
    torch.save({"epoch": 3, "iter_step": 1234, "model_state_dict": sd}, path)    # monosdf_train.py:277-281
    Model = ref_utils.get_class(CLASS)
    ours = Model(conf=conf)                                                      # eval.py:42
    ours.cuda()
    saved = torch.load(path)
    if list(saved["model_state_dict"].keys())[0].startswith("module."):          # eval.py:58-63
        saved["model_state_dict"] = {k[7:]: v for k, v in saved["model_state_dict"].items()}
    ours.load_state_dict(saved["model_state_dict"], strict=True)                 # eval.py:65
    ours.eval(), ref_model.eval()
    n = 1024
    rays = _cuda(port.synthetic_rays(n, seed=4))
    idx = torch.zeros(n, dtype=torch.long, device=DEV)
    out = ours({k: v.clone() for k, v in rays.items()}, idx, if_pixel_input=True)
    ref = ref_model({k: v.clone() for k, v in rays.items()}, idx, if_pixel_input=True)
    from tests.helpers import frac_within
    assert frac_within(out["z_vals"], ref["z_vals"], 1e-4) > 0.995
    for k in ["rgb_values", "depth_values", "normal_map"]:
        f = frac_within(out[k], ref[k], 2e-3)
        print("REPORT checkpoint round trip: %s within 2e-3 of the reference's CUDA render: %.4f" % (k, f))
        assert f > 0.99, (k, f)
