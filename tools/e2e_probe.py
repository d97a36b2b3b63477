"""Which part of the end-to-end step costs time: the per-step H2D copies, or reading the loss back?"""
import os, sys, time, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import _lib, confs, training
from monosdf_b200.model.loss import MonoSDFLoss
from monosdf_b200.model.network import MonoSDFNetwork
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MonoSDFNetwork(confs.to_conf(confs.SCANNET_MLP)).to(dev).train()
with torch.no_grad():
    model.density.beta.fill_(0.01)
model.set_precision("bf16")
arena, opt = training.build_optimizer(model)
loss_fn = MonoSDFLoss()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator().manual_seed(1)
o = (torch.rand(n, 3, generator=g) - 0.5) * 0.6
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
host = {"ray_dirs": d, "ray_cam_loc": o, "ray_dirs_tmp": d.clone(), "ray_pose": torch.eye(4)[None].repeat(n, 1, 1)}
g2 = torch.Generator().manual_seed(2)
host_gt = {"rgb": torch.rand(1, n, 3, generator=g2), "depth": torch.rand(1, n, 1, generator=g2) * 0.06 + 0.02,
           "normal": torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g2), dim=-1), "mask": torch.ones(1, n, 1)}
host = {k: v.pin_memory() for k, v in host.items()}
host_gt = {k: v.pin_memory() for k, v in host_gt.items()}
res = {k: v.to(dev) for k, v in host.items()}
res_gt = {k: v.to(dev) for k, v in host_gt.items()}
idx = torch.zeros(n, dtype=torch.long, device=dev)


def step(inp, gt):
    arena.zero_grad()
    out = model(inp, idx, if_pixel_input=True)
    loss = loss_fn(out, gt, if_pixel_input=True)["loss"]
    loss.backward()
    opt.step(grad_scale=1.0 / arena.all_reduce())
    return loss


def timed(fn, k=4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def h2d():
    return ({k: v.to(dev, non_blocking=True) for k, v in host.items()}, {k: v.to(dev, non_blocking=True) for k, v in host_gt.items()})


for _ in range(3):
    step(res, res_gt)
for rep in range(2):
    print("resident, no readback      %.1f ms" % timed(lambda: step(res, res_gt)))
    print("resident, loss.item()      %.1f ms" % timed(lambda: step(res, res_gt).item()))
    print("h2d each step, no readback %.1f ms" % timed(lambda: step(*h2d())))
    print("h2d each step, loss.item() %.1f ms" % timed(lambda: step(*h2d()).item()))
