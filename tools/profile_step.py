"""One training step of the bench workload inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).

  python tools/profile_step.py --rays 16384 --precision bf16 [--config grid]
"""
import argparse, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import _lib, confs, training
from monosdf_b200.model.loss import MonoSDFLoss
from monosdf_b200.model.network import MonoSDFNetwork

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=16384)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--config", default="mlp")
ap.add_argument("--beta", type=float, default=0.01)
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
conf = confs.SCANNET_MLP if a.config == "mlp" else confs.KITCHEN_GRIDS
torch.manual_seed(0)
model = MonoSDFNetwork(confs.to_conf(conf)).to(dev).train()
with torch.no_grad():
    model.density.beta.fill_(a.beta)
model.set_precision(a.precision)
arena, opt = training.build_optimizer(model)
loss_fn = MonoSDFLoss()
n = a.rays
g = torch.Generator().manual_seed(1)
o = (torch.rand(n, 3, generator=g) - 0.5) * 0.6
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
inp = {k: v.to(dev) for k, v in {"ray_dirs": d, "ray_cam_loc": o, "ray_dirs_tmp": d.clone(), "ray_pose": torch.eye(4)[None].repeat(n, 1, 1)}.items()}
g2 = torch.Generator().manual_seed(2)
gt = {k: v.to(dev) for k, v in {"rgb": torch.rand(1, n, 3, generator=g2), "depth": torch.rand(1, n, 1, generator=g2) * 0.06 + 0.02,
      "normal": torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g2), dim=-1), "mask": torch.ones(1, n, 1)}.items()}
idx = torch.zeros(n, dtype=torch.long, device=dev)


def step():
    arena.zero_grad()
    out = model(inp, idx, if_pixel_input=True)
    loss = loss_fn(out, gt, if_pixel_input=True)["loss"]
    loss.backward()
    opt.step(grad_scale=1.0 / arena.all_reduce())
    return loss


for _ in range(a.warm):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
l = step()
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("step ms %.3f loss %.6f rounds %d" % (e0.elapsed_time(e1), float(l), model.ray_sampler.last_total_iters))
