"""Sdf-only query throughput: the fused persistent kernel against the per-layer tensor-core sweep (CUDA events on the
launching stream, inputs far larger than L2).  python tools/fused_bench.py [--points N] [--conf mlp|grid] [--reps R]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monosdf_b200 import _lib, confs, roofline  # noqa: E402
from tests.helpers import build_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=65536 * 128)
    ap.add_argument("--conf", default="mlp")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    conf = confs.SCANNET_MLP if a.conf == "mlp" else confs.KITCHEN_GRIDS
    model = build_model({"conf": conf, "seed": 0, "beta": 0.01}, "cuda")
    model.set_precision("bf16")
    g = torch.Generator().manual_seed(1)
    x = ((torch.rand(a.points, 3, generator=g) * 2 - 1) * 1.2).cuda()
    macs = (roofline.WORK_MLP if a.conf == "mlp" else roofline.WORK_GRID)["A"]
    out = {}
    for name, fused in (("sweep", 0), ("fused", 1)):
        _lib.lib().msdf_set_fused(fused)
        with torch.no_grad():
            for _ in range(2):
                model.implicit_network.get_sdf_vals(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                model.implicit_network.get_sdf_vals(x)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        out[name] = {"ms": ms, "Mpoints_per_s": a.points / ms / 1e3, "algorithmic_TFLOPs": 2.0 * macs * a.points / ms / 1e9}
    _lib.lib().msdf_set_fused(1)
    out["speedup"] = out["sweep"]["ms"] / out["fused"]["ms"]
    out["points"] = a.points
    out["conf"] = a.conf
    print(json.dumps(out))


if __name__ == "__main__":
    main()
