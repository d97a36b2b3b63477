"""cProfile of the host side of training steps (python tools/train_host_profile.py [mlp|grid] [rays])."""
import cProfile, io, os, pstats, sys, time, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import confs, training
from monosdf_b200.model.loss import MonoSDFLoss
from monosdf_b200.model.network import MonoSDFNetwork
cfg = sys.argv[1] if len(sys.argv) > 1 else "grid"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MonoSDFNetwork(confs.to_conf(confs.KITCHEN_GRIDS if cfg == "grid" else confs.SCANNET_MLP)).to(dev).train()
with torch.no_grad():
    model.density.beta.fill_(0.01)
model.set_precision("bf16")
arena, opt = training.build_optimizer(model)
loss_fn = MonoSDFLoss()
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


for _ in range(4):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print("ms per step: %.2f" % ((time.perf_counter() - t0) / 10 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(16); print(st.getvalue()[:3400])
