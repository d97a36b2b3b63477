"""BASELINE config 4 on one B200: full-image render (384x384 rays, eval mode, uv input) in the reference's 1024-ray chunks
(utils/rend_util.split_input, evaluation/eval.py:100-125) and in one call, plus SDF-only grid queries for marching
cubes (utils/plots.py:148: implicit_network(points)[:, 0] in 100 000-point chunks) -- the dense 64^3 pass, the 6.03 M
queries of the coarse-to-fine pyramid on the init sphere and a dense-512^3-rate sample.

  python tools/eval_bench.py [--precision bf16|fp32] [--config mlp|grid]
Prints one JSON line.
"""
import argparse, json, os, sys, time, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import _lib, confs
from monosdf_b200.model.network import MonoSDFNetwork

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--config", default="mlp")
ap.add_argument("--res", type=int, default=384)
a = ap.parse_args()
dev = torch.device("cuda", 0)
conf = confs.SCANNET_MLP if a.config == "mlp" else confs.KITCHEN_GRIDS
torch.manual_seed(0)
model = MonoSDFNetwork(confs.to_conf(conf)).to(dev).eval()
with torch.no_grad():
    model.density.beta.fill_(0.01)
model.set_precision(a.precision)

R = a.res
ys, xs = torch.meshgrid(torch.arange(R), torch.arange(R), indexing="ij")
uv = torch.stack([xs, ys], -1).reshape(1, -1, 2).float().to(dev)
K = torch.eye(4)
K[0, 0] = K[1, 1] = 300.0 * R / 384
K[0, 2] = K[1, 2] = R / 2
pose = torch.eye(4)
pose[2, 3] = -0.3
inp = {"uv": uv, "intrinsics": K[None].to(dev), "pose": pose[None].to(dev)}
idx = torch.zeros(1, dtype=torch.long, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def render(chunk):
    outs = []
    with torch.no_grad():
        for s in range(0, R * R, chunk):
            part = dict(inp, uv=uv[:, s:s + chunk])
            o = model(part, idx)
            outs.append((o["rgb_values"], o["depth_values"], o["normal_map"]))
    return [torch.cat(t, 0) for t in zip(*outs)]


ms_1024, img_a = timed(lambda: render(1024), 2)
ms_full, img_b = timed(lambda: render(R * R), 3)
same = max(float((x - y).abs().max()) for x, y in zip(img_a, img_b))   # chunking changes the batch-global round count only


def sdf_query(pts, chunk):
    outs = []
    with torch.no_grad():
        for s in range(0, pts.shape[0], chunk):
            outs.append(model.implicit_network.get_sdf_vals(pts[s:s + chunk]))
    return torch.cat(outs, 0)


g = torch.linspace(-1.0, 1.0, 64, device=dev)
grid64 = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
ms_64, sdf64 = timed(lambda: sdf_query(grid64, 100000))
pyr = (torch.rand(6030000, 3, device=dev) * 2 - 1)
ms_pyr_100k, _ = timed(lambda: sdf_query(pyr, 100000), 2)
ms_pyr_full, _ = timed(lambda: sdf_query(pyr, 1 << 22), 2)
dense = (torch.rand(1 << 24, 3, device=dev) * 2 - 1)
ms_dense, _ = timed(lambda: sdf_query(dense, 1 << 24), 2)
print(json.dumps({
    "config": a.config, "precision": a.precision, "rays": R * R,
    "render_ms_1024_ray_chunks": ms_1024, "render_ms_one_call": ms_full,
    "render_rays_per_s_one_call": R * R / ms_full * 1e3, "max_abs_diff_between_chunkings": same,
    "sdf_grid_64cube_ms_100k_chunks": ms_64, "sdf_inside_fraction_64cube": float((sdf64 < 0).float().mean()),
    "sdf_pyramid_6.03M_ms_100k_chunks": ms_pyr_100k, "sdf_pyramid_6.03M_ms_4M_chunks": ms_pyr_full,
    "sdf_queries_per_s": (1 << 24) / ms_dense * 1e3, "dense_512cube_s_at_that_rate": 512 ** 3 / ((1 << 24) / ms_dense * 1e3),
}))
