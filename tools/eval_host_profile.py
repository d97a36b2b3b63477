"""cProfile of the eval path in the reference's 1024-ray chunks (host-bound: where do the ~6 ms per chunk go?)."""
import cProfile, io, os, pstats, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import confs
from monosdf_b200.model.network import MonoSDFNetwork
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MonoSDFNetwork(confs.to_conf(confs.SCANNET_MLP)).to(dev).eval()
with torch.no_grad():
    model.density.beta.fill_(0.01)
model.set_precision("bf16")
R = 384
ys, xs = torch.meshgrid(torch.arange(R), torch.arange(R), indexing="ij")
uv = torch.stack([xs, ys], -1).reshape(1, -1, 2).float().to(dev)
K = torch.eye(4); K[0, 0] = K[1, 1] = 300.0; K[0, 2] = K[1, 2] = R / 2
pose = torch.eye(4); pose[2, 3] = -0.3
inp = {"uv": uv, "intrinsics": K[None].to(dev), "pose": pose[None].to(dev)}
idx = torch.zeros(1, dtype=torch.long, device=dev)


def run(n):
    with torch.no_grad():
        for s in range(0, n * 1024, 1024):
            o = model(dict(inp, uv=uv[:, s:s + 1024]), idx)
    torch.cuda.synchronize()
    return o


run(5)
import time
t0 = time.perf_counter(); run(40); print("ms per 1024-ray chunk: %.2f" % ((time.perf_counter() - t0) / 40 * 1e3))
pr = cProfile.Profile(); pr.enable(); run(40); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(18); print(st.getvalue()[:3800])
