#!/usr/bin/env python
"""Benchmark of the MonoSDF rendering hot path: training rays/s (sampler + field + compositing + loss, fwd+bwd,
gradient all-reduce + fused Adam) on N B200s, with the roofline of the dominant kernel and the CPU baseline.

  python bench.py --gpus 1 --steps 5 --warmup 3                     our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference --steps 2 --warmup 1             the reference's CPU path (oracle/port.py)
  torchrun --nproc-per-node N ... bench.py --gpus N ...             one rank per GPU, rays sharded, weak scaling

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "training rays/s (fwd+bwd)"
UNIT = "rays/s"
# Algorithmic GFLOP per training ray, MLP conf, for k = 1..5 sampler rounds
from monosdf_b200.roofline import GFLOP_PER_RAY_GRID, GFLOP_PER_RAY_MLP  # noqa: E402  (derived from the layer dims)
CPU_SAMPLE_RAYS = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=65536, help="rays per step PER GPU (weak scaling)")
    ap.add_argument("--total-rays", type=int, default=0,
                    help="strong scaling (BASELINE config 5): rays per step of the WHOLE job, split evenly over the GPUs "
                         "(e.g. 262144); overrides --rays")
    ap.add_argument("--config", default="mlp", choices=["mlp", "grid"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="bf16: tcgen05 tensor-core mode (2e-2 parity, headline); fp32: SIMT mode (1e-4 parity)")
    ap.add_argument("--beta", type=float, default=0.01, help="density beta (0.01 -> 2 sampler rounds on the init sphere)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(args):
    if args.config == "mlp":
        net = "scannet_mlp-shaped MonoSDF MLP (8x256 SDF, PE 6, 2x256 colour, PE 4), ErrorBoundSampler 64+32+2"
    else:
        net = "kitchen_HDR_grids-shaped hash grid (16x2, 2^19, 16-2048) + 2x256 SDF MLP + 2x256 colour"
    if args.total_rays:
        return "%s, %d synthetic rays/step over all GPUs (strong scaling), pixel-mode rays, beta=%g, MonoSDFLoss, fwd+bwd+Adam" % (
            net, args.total_rays, args.beta)
    return "%s, %d synthetic rays/step/GPU, pixel-mode rays, beta=%g, MonoSDFLoss, fwd+bwd+Adam" % (net, args.rays, args.beta)


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "500"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """Start of the timed region: samples taken before it (warm-up) are not reported."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if t < getattr(self, "t0", 0.0):
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's PyTorch code on host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_step_time(conf, beta, n_rays, steps, warmup, threads):
    from oracle import port
    from monosdf_b200.model.network import MonoSDFNetwork
    from monosdf_b200.confs import to_conf
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = MonoSDFNetwork(to_conf(conf))          # parameter container only (CPU); the oracle does the arithmetic
    with torch.no_grad():
        model.density.beta.fill_(beta)
    cfg = port.cfg_from_conf(conf)
    if cfg.sdf.grid and cfg.sdf.use_grid_feature:
        raise RuntimeError("the reference has no CPU path for the hash grid (hashencoder.cu is CUDA-only)")
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    rays, gt = port.synthetic_rays(n_rays, seed=1), port.synthetic_gt(n_rays, seed=2)
    idx = torch.zeros(n_rays, dtype=torch.long)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in params.values():
            p.grad = None
        out = port.model_forward(params, cfg, rays, idx, if_pixel_input=True, training=True)
        loss = port.monosdf_loss(out, gt)
        loss["loss"].backward()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from monosdf_b200 import confs
    cores = os.cpu_count() or 1
    conf = confs.SCANNET_MLP          # the hash-grid conf has no CPU reference; the MLP conf figure is quoted for both
    t = cpu_step_time(conf, args.beta, CPU_SAMPLE_RAYS, args.steps, args.warmup, cores)
    v = CPU_SAMPLE_RAYS / t
    sample = "%d-ray fwd+bwd steps of the same workload (MLP conf), torch CPU, %d threads" % (CPU_SAMPLE_RAYS, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from monosdf_b200 import _lib, confs, training
    from monosdf_b200.model.loss import MonoSDFLoss
    from monosdf_b200.model.network import MonoSDFNetwork

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    conf = confs.SCANNET_MLP if args.config == "mlp" else confs.KITCHEN_GRIDS
    torch.manual_seed(0)
    model = MonoSDFNetwork(confs.to_conf(conf)).to(dev).train()
    with torch.no_grad():
        model.density.beta.fill_(args.beta)
        if args.config == "grid":     # non-trivial table content
            g = torch.Generator().manual_seed(3)
            e = model.implicit_network.encoding.embeddings
            e.copy_(((torch.rand(e.shape, generator=g) - 0.5) * 0.02).to(dev))
    model.set_precision(args.precision)
    arena, opt = training.build_optimizer(model)
    loss_fn = MonoSDFLoss()

    if args.total_rays:               # strong scaling: rank r renders rays [r N / W, (r + 1) N / W) of the step
        lo, hi = training.shard_range(args.total_rays, rank, world)
        args.rays = hi - lo
    n = args.rays
    g = torch.Generator().manual_seed(1 + rank)
    o = (torch.rand(n, 3, generator=g) - 0.5) * 0.6
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    host = {"ray_dirs": d, "ray_cam_loc": o, "ray_dirs_tmp": d.clone(), "ray_pose": torch.eye(4)[None].repeat(n, 1, 1)}
    g2 = torch.Generator().manual_seed(2 + rank)
    host_gt = {"rgb": torch.rand(1, n, 3, generator=g2), "depth": torch.rand(1, n, 1, generator=g2) * 0.06 + 0.02,
               "normal": torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g2), dim=-1), "mask": torch.ones(1, n, 1)}
    host = {k: v.pin_memory() for k, v in host.items()}
    host_gt = {k: v.pin_memory() for k, v in host_gt.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in list(host.values()) + list(host_gt.values()))
    indices = torch.zeros(n, dtype=torch.long, device=dev)
    res = {k: v.to(dev) for k, v in host.items()}
    res_gt = {k: v.to(dev) for k, v in host_gt.items()}

    def step(inp, gt):
        arena.zero_grad()
        out = model(inp, indices, if_pixel_input=True)
        loss = loss_fn(out, gt, if_pixel_input=True)["loss"]
        loss.backward()
        w = arena.all_reduce()
        opt.step(grad_scale=1.0 / w)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation, first query) stalls kernel launches for
    # a few hundred ms, which used to land in the first timed steps; only samples taken after mark() are reported
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        step(res, res_gt)
    rounds = model.ray_sampler.last_total_iters
    clocks.mark()
    l0 = _lib.launch_count()
    ms = timed(lambda: step(res, res_gt), args.steps)
    launches = _lib.launch_count() - l0

    def e2e_step():
        inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        gt = {k: v.to(dev, non_blocking=True) for k, v in host_gt.items()}
        return float(step(inp, gt).item())        # device -> host read of the loss

    for _ in range(2):      # untimed: the first end-to-end steps allocate the per-step input tensors (one-time cudaMalloc)
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clk = clocks.stop() if rank == 0 else None

    # one instrumented step: per-launch CUDA-event durations (on the launching stream) of every instrumented kernel
    # class; class 0 fp32 GEMM, 1 tcgen05 GEMM + weight gradient, 2 hash grid, 3 sampler, 4 compositing
    gemm_cls = 1 if args.precision == "bf16" else 0
    _lib.profile_read(0, reset=True)
    _lib.profile_enable(True)
    t_prof = timed(lambda: step(res, res_gt), 1)
    _lib.profile_enable(False)
    prof = {c: _lib.profile_read(c, reset=False) for c in range(5)}
    _lib.profile_read(0, reset=True)
    g_ms, g_flops, g_n, g_bytes = prof[gemm_cls]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_bw = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json: bf16_tflops_sustained, hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)"
    achieved = (g_flops / 1e12) / (g_ms / 1e3) if g_ms > 0 else 0.0
    gbs = (g_bytes / 1e9) / (g_ms / 1e3) if g_ms > 0 else 0.0
    table = GFLOP_PER_RAY_MLP if args.config == "mlp" else GFLOP_PER_RAY_GRID
    step_tf = table[min(max(rounds, 1), 5)] * 1e9 * n * args.steps / (ms / 1e3) / 1e12
    rays_per_step = args.total_rays if args.total_rays else world * n
    value = rays_per_step * args.steps / (ms / 1e3)
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram bytes / launch of the same kernel from ncu --set full
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get("tcgen05_gemm_dram_bytes_per_launch")
    others = {}
    for cls, name in ((2, "hash grid gather/scatter"), (3, "sampler round (warp scans)"), (4, "compositing fwd+bwd")):
        k_ms, k_work, k_n, k_bytes = prof[cls]
        if k_n:
            others[name] = {"bound": "hbm", "achieved": (k_bytes / 1e9) / (k_ms / 1e3), "peak": peak_bw, "unit": "GB/s",
                            "frac": (k_bytes / 1e9) / (k_ms / 1e3) / peak_bw, "launches": int(k_n), "ms_per_step": k_ms}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_rays else "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "rays_per_step_per_gpu": n, "sampler_rounds": rounds,
                   "precision_mode": args.precision, "parallelism": "ray-sharded dp%d, one NCCL all-reduce of the flat gradient arena" % world,
                   "l2_policy": "inputs larger than L2: every field chunk streams %.1f GB of activations through HBM (L2 is 126 MB)"
                                % (min(n * 98, 262144) * 9.9e3 / 1e9)},
        "clocks": clk,
        "e2e": {"value": rays_per_step * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "k_gemm (fp32 SIMT MLP sweeps)" if gemm_cls == 0 else "k_tc_gemm + k_tc_wgrad (tcgen05 MLP sweeps)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                     "traffic": traffic, "peak_source": peak_src, "launches": int(g_n), "kernel_ms_per_step": g_ms,
                     "kernel_share_of_step": g_ms / t_prof if t_prof > 0 else None,
                     "step_algorithmic_tflops": step_tf,
                     # the bound that applies to a per-layer GEMM moving 1-2.5 KB per 131 KFLOP: algorithmic bytes / time
                     "hbm_view": {"achieved": gbs, "peak": peak_bw, "unit": "GB/s", "frac": gbs / peak_bw if peak_bw else None},
                     "other_kernels": others},
    }
    if world == 1 and not args.no_cpu_baseline and args.config == "mlp":
        cores = os.cpu_count() or 1
        t = cpu_step_time(conf, args.beta, CPU_SAMPLE_RAYS, 2, 1, cores)
        line["cpu_baseline"] = {"value": CPU_SAMPLE_RAYS / t, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "2 timed %d-ray fwd+bwd steps of the same workload, oracle/port.py on torch CPU" % CPU_SAMPLE_RAYS}
    elif world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t = cpu_step_time(confs.SCANNET_MLP, args.beta, CPU_SAMPLE_RAYS, 2, 1, cores)
        line["cpu_baseline"] = {"value": CPU_SAMPLE_RAYS / t, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "MLP-conf figure (the reference has no CPU hash grid): 2 timed %d-ray steps" % CPU_SAMPLE_RAYS}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
