"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/code) on CPU.

Container-only (the reference tree does not exist on the GPU box); the fixtures it writes are committed.
TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/make_golden.py

Each fixture holds: the model conf, the seed the reference constructor was run under (monosdf_b200's constructor
consumes torch's RNG identically, checked by tests/test_reference_pin.py, so weights are reproducible from the
seed without shipping them), the synthetic rays, and the reference's outputs / loss / parameter gradients.
"""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import port, ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

SMALL_CONF = copy.deepcopy(ref_shim.MLP_CONF)
SMALL_CONF["feature_vector_size"] = 32
SMALL_CONF["implicit_network"].update(dims=[64, 64, 64, 64], skip_in=[2], multires=4)
SMALL_CONF["rendering_network"].update(dims=[64, 64], multires_view=2)
SMALL_CONF["ray_sampler"].update(N_samples=16, N_samples_eval=32, N_samples_extra=8)

GRIDMLP_CONF = copy.deepcopy(SMALL_CONF)     # the fork's "MLP" confs: Grid class with zero grid features
GRIDMLP_CONF["Grid_MLP"] = True
GRIDMLP_CONF["implicit_network"].update(use_grid_feature=False, divide_factor=1.1, num_levels=4, level_dim=2,
                                        base_size=4, end_size=32, logmap=10)

# RenderingNetwork(spec=True): diffuse/specular split after layer 2, HDR only (network.py:376-380, 427-454; dims as in
# confs/archive/kitchen_hdr_est_grids_spec.conf:106)
SPEC_SMALL_CONF = copy.deepcopy(SMALL_CONF)
SPEC_SMALL_CONF["rendering_network"].update(dims=[64, 64, 67, 64], spec=True)
SPEC_FULL_CONF = copy.deepcopy(ref_shim.MLP_CONF)
SPEC_FULL_CONF["rendering_network"].update(dims=[256, 256, 259, 256], spec=True)

CASES = {
    # name: (conf, n_rays, beta, store_full_grads[, if_hdr])
    "mlp_full": (ref_shim.MLP_CONF, 24, 0.01, False),
    "mlp_small": (SMALL_CONF, 48, 0.02, True),
    "gridmlp_small": (GRIDMLP_CONF, 32, 0.02, True),
    "spec_small": (SPEC_SMALL_CONF, 40, 0.02, True, True),
    "spec_full": (SPEC_FULL_CONF, 16, 0.01, False, True),
}


def run_case(name, conf, n_rays, beta, full_grads, if_hdr=False, seed=0):
    net = ref_shim.load_reference()
    torch.manual_seed(seed)
    model = net.MonoSDFNetwork(conf=ref_shim.to_conf(conf), if_hdr=if_hdr)
    with torch.no_grad():
        model.density.beta.fill_(beta)
    rays = port.synthetic_rays(n_rays, seed=1)
    gt = port.synthetic_gt(n_rays, seed=2)
    indices = torch.zeros(n_rays, dtype=torch.long)
    fx = dict(name=name, conf=conf, seed=seed, beta=beta, n_rays=n_rays, if_hdr=if_hdr)
    # ---- eval mode (deterministic sampling)
    model.eval()
    out = model({k: v.clone() for k, v in rays.items()}, indices, if_pixel_input=True)
    fx["eval"] = {k: v.detach().clone() for k, v in out.items()}
    # ---- uv (image) input path, eval
    npx = 16
    uv = torch.stack(torch.meshgrid(torch.arange(4.0), torch.arange(4.0), indexing="xy"), -1).reshape(1, npx, 2) * 90 + 20
    intr = torch.eye(4)[None].clone()
    intr[0, 0, 0] = intr[0, 1, 1] = 300.0
    intr[0, 0, 2] = intr[0, 1, 2] = 192.0
    pose = torch.eye(4)[None].clone()
    c, s = 0.8, 0.6
    pose[0, :3, :3] = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    pose[0, :3, 3] = torch.tensor([0.1, -0.05, 0.2])
    uv_in = dict(uv=uv, pose=pose, intrinsics=intr)
    out = model({k: v.clone() for k, v in uv_in.items()}, torch.zeros(1, dtype=torch.long))
    fx["uv_input"] = uv_in
    fx["uv_eval"] = {k: v.detach().clone() for k, v in out.items()}
    # ---- train mode: one fwd + loss + bwd with the CPU generator seeded
    model.train()
    torch.manual_seed(1234)
    out = model({k: v.clone() for k, v in rays.items()}, indices, if_pixel_input=True)
    loss = port.monosdf_loss(out, gt)
    if "rgb_spec_values" in out:    # give the specular output a gradient of its own (MonoSDFLoss does not read it)
        loss["loss"] = loss["loss"] + 0.25 * (out["rgb_spec_values"] * gt["rgb"].reshape(-1, 3)).mean()
    model.zero_grad()
    loss["loss"].backward()
    fx["train_seed"] = 1234
    fx["train"] = {k: v.detach().clone() for k, v in out.items()}
    fx["train_loss"] = {k: v.detach().clone() for k, v in loss.items()}
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    fx["train_grad_norm"] = {k: g.norm() for k, g in grads.items()}
    fx["train_grad_head"] = {k: g.flatten()[:64].clone() for k, g in grads.items()}
    if full_grads:
        fx["train_grad"] = grads
        fx["state_dict"] = {k: v.detach().clone() for k, v in model.state_dict().items()}
    fx["state_checksum"] = {k: v.double().sum() for k, v in model.state_dict().items()}
    os.makedirs(OUT, exist_ok=True)
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, "loss", float(loss["loss"]), "rgb", out["rgb_values"][0].tolist())


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]            # e.g. `python oracle/make_golden.py spec_small spec_full`: leave the others untouched
    for name, case in CASES.items():
        if not only or name in only:
            run_case(name, *case)
