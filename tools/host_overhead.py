"""Host-side cost of one training step: wall time of enqueueing a step (no device sync inside except the sampler's
per-round flag read) against the device time of the same step.  python tools/host_overhead.py --rays 32768"""
import argparse, os, sys, time, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import _lib, confs, training
from monosdf_b200.model.loss import MonoSDFLoss
from monosdf_b200.model.network import MonoSDFNetwork
ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=32768)
ap.add_argument("--config", default="mlp")
a = ap.parse_args()
dev = torch.device("cuda", 0)
conf = confs.SCANNET_MLP if a.config == "mlp" else confs.KITCHEN_GRIDS
torch.manual_seed(0)
model = MonoSDFNetwork(confs.to_conf(conf)).to(dev).train()
with torch.no_grad():
    model.density.beta.fill_(0.01)
model.set_precision("bf16")
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
marks = {}
ev = {}
_orig = model.ray_sampler.get_z_vals


def _wrapped(*args, **kw):
    r = _orig(*args, **kw)
    ev["sampler_end"] = torch.cuda.Event(enable_timing=True)
    ev["sampler_end"].record()
    ev["sampler_host"] = time.perf_counter()
    return r


model.ray_sampler.get_z_vals = _wrapped


def step(record=False):
    t0 = time.perf_counter()
    ev["start"] = torch.cuda.Event(enable_timing=True); ev["start"].record()
    arena.zero_grad()
    out = model(inp, idx, if_pixel_input=True)
    ev["fwd_end"] = torch.cuda.Event(enable_timing=True); ev["fwd_end"].record()
    t1 = time.perf_counter()
    loss = loss_fn(out, gt, if_pixel_input=True)["loss"]
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    ev["bwd_end"] = torch.cuda.Event(enable_timing=True); ev["bwd_end"].record()
    opt.step(grad_scale=1.0 / arena.all_reduce())
    t4 = time.perf_counter()
    if record:
        marks.setdefault("sampler_host", []).append((ev["sampler_host"] - t0) * 1e3)
        torch.cuda.synchronize()
        marks.setdefault("gpu_sampler", []).append(ev["start"].elapsed_time(ev["sampler_end"]))
        marks.setdefault("gpu_render_fwd", []).append(ev["sampler_end"].elapsed_time(ev["fwd_end"]))
        marks.setdefault("gpu_loss_bwd", []).append(ev["fwd_end"].elapsed_time(ev["bwd_end"]))
        for k, v in (("forward", t1 - t0), ("loss", t2 - t1), ("backward", t3 - t2), ("optimizer", t4 - t3)):
            marks.setdefault(k, []).append(v * 1e3)


for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(True)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    marks.setdefault("host_total", []).append(t_host * 1e3)
    marks.setdefault("step_total", []).append(t_all * 1e3)
for k, v in marks.items():
    print("%-12s host ms: %s" % (k, " ".join("%.1f" % x for x in v)))
