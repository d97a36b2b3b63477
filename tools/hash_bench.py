"""Hash-grid operators at production geometry (16 x 2, 2^19, 16 -> 2048): ours against the REFERENCE's own kernels
(hashencoder.cu:104,258,432 compiled untouched into oracle/_ref/) on the same box, same inputs.  CUDA events, best of
`reps` after a warm-up; algorithmic bytes per point from SURVEY 8d.  python tools/hash_bench.py [--points N]"""
import argparse, json, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from monosdf_b200 import _lib
from monosdf_b200.hashencoder import HashEncoder
from oracle import build_ref_hashencoder

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=262144)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--order", default="rays", choices=["rays", "random"])
a = ap.parse_args()
DEV = "cuda"
L, C, D, H = 16, 2, 3, 16
enc = HashEncoder(3, L, C, 2, H, 19, 2048).to(DEV)
g = torch.Generator().manual_seed(3)
emb = (torch.rand(enc.embeddings.shape, generator=g) - 0.5).to(DEV).contiguous()
B = a.points
if a.order == "rays":      # 128 consecutive samples per ray through the unit cube (what the sampler / renderer feed)
    n_rays = B // 128
    o = torch.rand(n_rays, 1, 3, generator=g) * 0.4 + 0.3
    d = torch.nn.functional.normalize(torch.randn(n_rays, 1, 3, generator=g), dim=-1)
    t = torch.linspace(-0.3, 0.3, 128).view(1, 128, 1)
    x = (o + t * d).reshape(-1, 3).clamp(0.0, 1.0)
else:
    x = torch.rand(B, 3, generator=g)
x = x.to(DEV).contiguous()
B = x.shape[0]
S = float(np.log2(enc.per_level_scale))
offsets = enc.offsets.contiguous()
grad = torch.randn(L, B, C, generator=g).to(DEV).contiguous()
gg_in = torch.randn(B, D, generator=g).to(DEV).contiguous()
out = torch.empty(L, B, C, device=DEV)
rows = torch.empty(B, L * C, device=DEV)
dydx = torch.empty(B, L * D * C, device=DEV)
ge, gi, gg = torch.zeros_like(emb), torch.zeros_like(x), torch.zeros_like(grad)
ref = build_ref_hashencoder.load()


def best(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


P = _lib.ptr
st = _lib.stream()
cases = {
    "forward": (1164.0,
                lambda: _lib.call("msdf_hash_encode_forward", P(x), P(emb), P(offsets), P(out), B, D, C, L, S, H, 0, None, st),
                (lambda: ref.hash_encode_forward(x, emb, offsets, out, B, D, C, L, S, H, False, dydx)) if ref else None),
    "forward+dy_dx": (1164.0 + 384.0,
                      lambda: _lib.call("msdf_hash_encode_forward", P(x), P(emb), P(offsets), P(out), B, D, C, L, S, H, 1, P(dydx), st),
                      (lambda: ref.hash_encode_forward(x, emb, offsets, out, B, D, C, L, S, H, True, dydx)) if ref else None),
    "backward (scatter + input grad)": (2188.0 + 524.0,
                 lambda: _lib.call("msdf_hash_encode_backward", P(grad), P(x), P(emb), P(offsets), P(ge), B, D, C, L, S, H, 1, P(dydx), P(gi), st),
                 (lambda: ref.hash_encode_backward(grad, x, emb, offsets, ge, B, D, C, L, S, H, True, dydx, gi)) if ref else None),
    "second backward": (2200.0 + 524.0,
                 lambda: _lib.call("msdf_hash_encode_second_backward", P(grad), P(x), P(emb), P(offsets), B, D, C, L, S, H, 1, P(dydx), P(gg_in), P(gg), P(ge), st),
                 (lambda: ref.hash_encode_second_backward(grad, x, emb, offsets, B, D, C, L, S, H, True, dydx, gg_in, gg, ge)) if ref else None),
}
res = {"points": B, "order": a.order}
for name, (bytes_pp, ours, theirs) in cases.items():
    us = best(ours)
    r = {"ours_us": us, "ours_GBs_algorithmic": bytes_pp * B / us / 1e3}
    if theirs is not None:
        ut = best(theirs)
        r.update({"reference_us": ut, "speedup_vs_reference_kernels": ut / us})
    res[name] = r
# the engine's fused row variant (features straight into [B, 32] rows, what the field kernels call)
from monosdf_b200.model import network  # noqa: F401  (loads nothing new; keeps the import graph honest)
print(json.dumps(res))
