"""Per-step times of the end-to-end training step with the allocator's cudaMalloc count and the SM clock beside them
(is the step-to-step variation allocator churn or the power cap?):  python tools/step_jitter.py [rays] [steps] [mlp|grid]"""
import os, subprocess, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import confs, training
from monosdf_b200.model.loss import MonoSDFLoss
from monosdf_b200.model.network import MonoSDFNetwork
dev = torch.device("cuda", 0)
torch.manual_seed(0)
CONF = confs.KITCHEN_GRIDS if (len(sys.argv) > 3 and sys.argv[3] == "grid") else confs.SCANNET_MLP
model = MonoSDFNetwork(confs.to_conf(CONF)).to(dev).train()
with torch.no_grad():
    model.density.beta.fill_(0.01)
model.set_precision("bf16")
arena, opt = training.build_optimizer(model)
loss_fn = MonoSDFLoss()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 14
g = torch.Generator().manual_seed(1)
o = (torch.rand(n, 3, generator=g) - 0.5) * 0.6
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
host = {"ray_dirs": d, "ray_cam_loc": o, "ray_dirs_tmp": d.clone(), "ray_pose": torch.eye(4)[None].repeat(n, 1, 1)}
g2 = torch.Generator().manual_seed(2)
host_gt = {"rgb": torch.rand(1, n, 3, generator=g2), "depth": torch.rand(1, n, 1, generator=g2) * 0.06 + 0.02,
           "normal": torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g2), dim=-1), "mask": torch.ones(1, n, 1)}
host = {k: v.pin_memory() for k, v in host.items()}
host_gt = {k: v.pin_memory() for k, v in host_gt.items()}
idx = torch.zeros(n, dtype=torch.long, device=dev)


def clock():
    try:
        return subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits"],
                              capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception:
        return "?"


for it in range(steps):
    s0 = torch.cuda.memory_stats(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    gt = {k: v.to(dev, non_blocking=True) for k, v in host_gt.items()}
    arena.zero_grad()
    out = model(inp, idx, if_pixel_input=True)
    loss = loss_fn(out, gt, if_pixel_input=True)["loss"]
    loss.backward()
    opt.step(grad_scale=1.0 / arena.all_reduce())
    lv = loss.item()
    e1.record()
    torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats(dev)
    print("step %2d  %.1f ms  cudaMalloc +%d  cudaFree +%d  retries +%d  reserved %.1f GB  [sm MHz, W, C: %s]" % (
        it, e0.elapsed_time(e1), s1["num_device_alloc"] - s0["num_device_alloc"], s1["num_device_free"] - s0["num_device_free"],
        s1["num_alloc_retries"] - s0["num_alloc_retries"], s1["reserved_bytes.all.current"] / 1e9, clock()))
