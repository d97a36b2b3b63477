#!/usr/bin/env python
"""Benchmark of the MonoSDF rendering hot path: training rays/s (sampler + field + compositing + loss, fwd+bwd,
gradient all-reduce + fused Adam) on N B200s, with the roofline of the dominant kernel class and the CPU baseline.

  python bench.py --gpus 1 --steps 5 --warmup 3                     our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference --steps 2 --warmup 1             the reference's own CPU path on the host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...             one rank per GPU, rays sharded, weak scaling

Prints ONE JSON line (rank 0).  The headline is BASELINE.json configs[1] (scannet-MLP-shaped net, 65 536 rays / step /
GPU, weak scaling); `sub_records` carries short measurements of configs 3 (hash-grid conf, 49.95 MB all-reduce),
5 (2^18 rays / step strong scaling, micro-batched where a shard exceeds 65 536 rays) and 4 (384 x 384 eval render in
1024-ray chunks + the marching-cubes SDF pyramid).  See DESIGN.md "Measurement" for how each field is obtained.

The workload is STATIONARY: parameters and Adam state are restored from a snapshot before every step (a 2.7 MB
device-to-device copy; Adam still runs, the all-reduce still runs), so the sampler's round count cannot drift while the
timed region runs; the rounds of every timed step are recorded and their mean is in `config.sampler_rounds`.
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
from monosdf_b200 import roofline  # noqa: E402  (algorithmic work per ray, derived from the layer dimensions)
CPU_SAMPLE_RAYS = 1024              # the reference's own batch size (num_pixels, mi.conf:18)
MICRO_BATCH = 65536                 # rays per forward/backward when a rank's shard is larger (gradient accumulation)
LOSS_CONF = dict(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.05, smooth_weight=0.005, depth_weight=0.1,
                 normal_l1_weight=0.05, normal_cos_weight=0.05)          # mi.conf:38-43


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
    ap.add_argument("--no-sub-records", action="store_true", help="headline workload only")
    ap.add_argument("--cpu-rays", type=int, default=CPU_SAMPLE_RAYS, help="rays per step of the CPU arm")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (0: all host cores)")
    return ap.parse_args()


def net_name(config):
    if config == "mlp":
        return "scannet_mlp-shaped MonoSDF MLP (8x256 SDF, PE 6, 2x256 colour, PE 4), ErrorBoundSampler 64+32+2"
    return "kitchen_HDR_grids-shaped hash grid (16x2, 2^19, 16-2048) + 2x256 SDF MLP + 2x256 colour"


def workload_name(config, rays, total_rays, beta):
    if total_rays:
        return "%s, %d synthetic rays/step over all GPUs (strong scaling), pixel-mode rays, beta=%g, MonoSDFLoss, fwd+bwd+Adam" % (
            net_name(config), total_rays, beta)
    return "%s, %d synthetic rays/step/GPU, pixel-mode rays, beta=%g, MonoSDFLoss, fwd+bwd+Adam" % (net_name(config), rays, beta)


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0, self.t1 = 0.0, 1e30

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

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if t < self.t0 or t > self.t1:
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
# reference arm / CPU baseline: the reference's own PyTorch code on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_step_time(beta, n_rays, steps, warmup, threads):
    """Seconds per training step (fwd + MonoSDFLoss + bwd + torch.optim.Adam) of the hot path on the CPU.  Runs the
    REFERENCE'S OWN files (model/network.py MonoSDFNetwork, model/loss.py MonoSDFLoss; /root/reference in the build
    container, the copies staged under baseline/_ref/ on the GPU box) when they are there -> kind 'reference';
    otherwise the oracle port (oracle/port.py, pinned to the reference by tests/test_oracle_golden.py) -> kind 'port'."""
    from oracle import port, ref_shim
    from monosdf_b200 import confs
    torch.set_num_threads(threads)
    conf = confs.SCANNET_MLP       # (the hash-grid conf has no CPU path in the reference: hashencoder.cu is CUDA-only)
    rays, gt = port.synthetic_rays(n_rays, seed=1), port.synthetic_gt(n_rays, seed=2)
    idx = torch.zeros(n_rays, dtype=torch.long)
    torch.manual_seed(0)
    if ref_shim.reference_available():
        kind = "reference"
        net = ref_shim.load_reference()
        from model.loss import MonoSDFLoss as RefLoss     # the reference's module (ref_shim put its tree on sys.path)
        model = net.MonoSDFNetwork(conf=ref_shim.to_conf(conf)).train()
        with torch.no_grad():
            model.density.beta.fill_(beta)
        loss_fn = RefLoss(**LOSS_CONF)
        opt = torch.optim.Adam(model.parameters(), lr=5.0e-4)          # monosdf_train.py:221

        def step():
            opt.zero_grad()
            out = model({k: v.clone() for k, v in rays.items()}, idx, if_pixel_input=True)
            loss_fn(out, gt, if_pixel_input=True)["loss"].backward()
            opt.step()
    else:
        kind = "port"
        from monosdf_b200.model.network import MonoSDFNetwork
        model = MonoSDFNetwork(confs.to_conf(conf))          # parameter container only (CPU); the oracle does the arithmetic
        with torch.no_grad():
            model.density.beta.fill_(beta)
        cfg = port.cfg_from_conf(conf)
        params = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
        opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=5.0e-4)

        def step():
            opt.zero_grad()
            out = port.model_forward(params, cfg, rays, idx, if_pixel_input=True, training=True)
            port.monosdf_loss(out, gt)["loss"].backward()
            opt.step()
    import contextlib
    times = []
    with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):      # (the reference's loss prints shapes, loss.py:164)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""          # host cores only: the reference's hard-coded .cuda() calls become no-ops
    cores = args.cpu_threads or os.cpu_count() or 1
    n = args.cpu_rays
    t, kind = cpu_reference_step_time(args.beta, n, args.steps, args.warmup, cores)
    t1 = None
    if not args.cpu_threads and cores > 1:
        # the reference trainer itself runs with torch.set_num_threads(1) (monosdf_train.py:37): one bounded step of that too
        t1, _ = cpu_reference_step_time(args.beta, n, 1, 1, 1)
    v = n / t
    what = "the reference's own model/network.py + model/loss.py + torch.optim.Adam" if kind == "reference" else \
        "oracle/port.py (restatement of the reference, pinned by tests/test_oracle_golden.py) + torch.optim.Adam"
    sample = "%d-ray training steps (fwd+loss+bwd+Adam) of the MLP-conf workload, %s, torch CPU, %d threads" % (n, what, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name("mlp", args.rays, args.total_rays, args.beta), "sample": sample,
                   "note": "each step is a bounded %d-ray sample of the workload (the reference's own batch size is 1024, mi.conf:18)" % n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if t1 is not None:
        line["cpu_baseline_1thread"] = {"value": n / t1, "unit": UNIT, "cores": 1, "kind": kind,
                                        "sample": "one timed %d-ray step after one warm-up, 1 thread (monosdf_train.py:37)" % n}
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
class Workload:
    """One training workload on this rank: model, flat-arena optimizer, synthetic rays (host + device copies) and a
    stationary step (parameters / Adam state restored from a snapshot before every step)."""

    def __init__(self, config, n_rays, precision, beta, rank, world, dev):
        from monosdf_b200 import confs, training
        from monosdf_b200.model.loss import MonoSDFLoss
        from monosdf_b200.model.network import MonoSDFNetwork
        self.dev, self.world, self.n = dev, world, n_rays
        conf = confs.SCANNET_MLP if config == "mlp" else confs.KITCHEN_GRIDS
        torch.manual_seed(0)
        self.model = MonoSDFNetwork(confs.to_conf(conf)).to(dev).train()
        with torch.no_grad():
            self.model.density.beta.fill_(beta)
            if config == "grid":     # non-trivial table content
                g = torch.Generator().manual_seed(3)
                e = self.model.implicit_network.encoding.embeddings
                e.copy_(((torch.rand(e.shape, generator=g) - 0.5) * 0.02).to(dev))
        self.model.set_precision(precision)
        self.arena, self.opt = training.build_optimizer(self.model)
        self.loss_fn = MonoSDFLoss()
        self.micro = [(s, min(s + MICRO_BATCH, n_rays)) for s in range(0, n_rays, MICRO_BATCH)]
        n = n_rays
        g = torch.Generator().manual_seed(1 + rank)
        o = (torch.rand(n, 3, generator=g) - 0.5) * 0.6
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        host = {"ray_dirs": d, "ray_cam_loc": o, "ray_dirs_tmp": d.clone(), "ray_pose": torch.eye(4)[None].repeat(n, 1, 1)}
        g2 = torch.Generator().manual_seed(2 + rank)
        host_gt = {"rgb": torch.rand(1, n, 3, generator=g2), "depth": torch.rand(1, n, 1, generator=g2) * 0.06 + 0.02,
                   "normal": torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g2), dim=-1), "mask": torch.ones(1, n, 1)}
        self.host = {k: v.pin_memory() for k, v in host.items()}
        self.host_gt = {k: v.pin_memory() for k, v in host_gt.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in list(self.host.values()) + list(self.host_gt.values()))
        self.indices = torch.zeros(n, dtype=torch.long, device=dev)
        self.res = {k: v.to(dev) for k, v in self.host.items()}
        self.res_gt = {k: v.to(dev) for k, v in self.host_gt.items()}
        self.snap = None
        self.rounds = []
        self.grad_bytes = self.arena.grad.numel() * 4

    def snapshot(self):
        self.snap = (self.arena.flat.clone(), self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(), self.opt.step_count)

    def restore(self):
        if self.snap is not None:
            self.arena.flat.copy_(self.snap[0])
            self.opt.exp_avg.copy_(self.snap[1])
            self.opt.exp_avg_sq.copy_(self.snap[2])
            self.opt.step_count = self.snap[3]

    def step(self, inp, gt):
        """zero_grad -> forward -> loss -> backward (per micro-batch) -> all-reduce -> Adam; returns the last loss."""
        self.restore()
        self.arena.zero_grad()
        loss = None
        for lo, hi in self.micro:
            if len(self.micro) == 1:
                mi, mg, idx = inp, gt, self.indices
            else:
                mi = {k: v[lo:hi] for k, v in inp.items()}
                mg = {k: v[:, lo:hi] for k, v in gt.items()}
                idx = self.indices[lo:hi]
            out = self.model(mi, idx, if_pixel_input=True)
            loss = self.loss_fn(out, mg, if_pixel_input=True)["loss"]
            loss.backward()
            self.rounds.append(self.model.ray_sampler.last_total_iters)
        w = self.arena.all_reduce()
        self.opt.step(grad_scale=1.0 / (w * len(self.micro)))
        return loss

    def step_resident(self):
        return self.step(self.res, self.res_gt)

    def step_e2e(self):
        inp = {k: v.to(self.dev, non_blocking=True) for k, v in self.host.items()}
        gt = {k: v.to(self.dev, non_blocking=True) for k, v in self.host_gt.items()}
        return float(self.step(inp, gt).item())        # device -> host read of the loss


def make_timer(world, dev):
    import torch.distributed as dist

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
    return timed


def profile_classes(wl, timed, steps):
    """`steps` instrumented steps: per-launch CUDA-event durations (on the launching stream) of every instrumented kernel
    class; class 0 fp32 GEMM, 1 tcgen05 (per-layer GEMMs, weight gradients, the fused sdf network), 2 hash grid,
    3 sampler, 4 compositing.  Returns ({class: (ms, work, launches, bytes) per step}, ms per instrumented step)."""
    from monosdf_b200 import _lib
    _lib.profile_read(0, reset=True)
    _lib.profile_enable(True)
    t = timed(wl.step_resident, steps)
    _lib.profile_enable(False)
    prof = {c: tuple(v / steps for v in _lib.profile_read(c, reset=False)) for c in range(5)}
    for k in range(1, 7):          # sub-classes of the tcgen05 class (include/monosdf_b200.h)
        prof[1 | (k << 8)] = tuple(v / steps for v in _lib.profile_read(1 | (k << 8), reset=False))
    _lib.profile_read(0, reset=True)
    return prof, t / steps


def mean_rounds(rounds):
    return sum(rounds) / max(len(rounds), 1)


def gflop_per_ray(config, k):
    """algorithmic GFLOP per training ray for a (possibly fractional: mean over steps) number of sampler rounds k"""
    work = roofline.WORK_MLP if config == "mlp" else roofline.WORK_GRID
    lo = int(k)
    a, b = roofline.gflop_per_ray(work, lo), roofline.gflop_per_ray(work, lo + 1)
    return a + (b - a) * (k - lo)


def run_sub_grid(args, rank, world, dev, timed, peak_tf):
    """BASELINE config 3: hash-grid conf, 32 768 rays / GPU, weak scaling; the all-reduce carries the 46.5 MB table gradient."""
    n = 32768
    wl = Workload("grid", n, args.precision, args.beta, rank, world, dev)
    for _ in range(3):
        wl.step_resident()
    wl.snapshot()
    wl.rounds = []
    k = 5
    ms = timed(wl.step_resident, k) / k
    rounds = mean_rounds(wl.rounds)
    ms_e2e = timed(wl.step_e2e, k) / k
    prof, t_prof = profile_classes(wl, timed, 2)
    tc_ms, hash_ms = prof[1][0], prof[2][0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        wl.arena.all_reduce()
    e1.record()
    torch.cuda.synchronize()
    ar_ms = e0.elapsed_time(e1) / 5
    rec = {
        "workload": workload_name("grid", n, 0, args.beta), "scaling": "weak", "value": world * n / ms * 1e3, "unit": UNIT,
        "ms_per_step": ms, "e2e_value": world * n / ms_e2e * 1e3, "sampler_rounds": rounds,
        "allreduce_bytes": wl.grad_bytes, "allreduce_ms": ar_ms,
        "tcgen05_kernel_share_of_step": tc_ms / t_prof, "hash_kernel_share_of_step": hash_ms / t_prof,
        "algorithmic_tflops": gflop_per_ray("grid", rounds) * n / ms,            # GFLOP / ms = TFLOP / s
        "frac_of_bf16_peak": gflop_per_ray("grid", rounds) * n / ms / peak_tf,
        "limiter": "kernel time outside the tensor cores (hash gather/scatter, sampler scans, row producers) and host launch "
                   "overhead at 32 768 rays; the all-reduce (%.2f ms of %.1f) is not" % (ar_ms, ms),
    }
    del wl
    torch.cuda.empty_cache()
    return rec


def run_sub_strong(args, rank, world, dev, timed, peak_tf):
    """BASELINE config 5: 2^18 rays per step over all GPUs; a shard above 65 536 rays runs as micro-batches of 65 536 with
    gradient accumulation (saved activations fit, no recompute), one all-reduce + Adam per step."""
    from monosdf_b200 import training
    total = 262144
    lo, hi = training.shard_range(total, rank, world)
    assert total % world == 0, "equal shards: each rank's loss is a mean over its own rays"
    wl = Workload("mlp", hi - lo, args.precision, args.beta, rank, world, dev)
    for _ in range(2 if len(wl.micro) > 1 else 3):
        wl.step_resident()
    wl.snapshot()
    wl.rounds = []
    k = 2 if len(wl.micro) > 1 else 5
    ms = timed(wl.step_resident, k) / k
    rounds = mean_rounds(wl.rounds)
    rec = {
        "workload": workload_name("mlp", 0, total, args.beta), "scaling": "strong", "value": total / ms * 1e3, "unit": UNIT,
        "ms_per_step": ms, "rays_per_gpu": hi - lo, "micro_batches": len(wl.micro), "sampler_rounds": rounds,
        "algorithmic_tflops_per_gpu": gflop_per_ray("mlp", rounds) * (hi - lo) / ms,
        "limiter": "per-GPU kernel efficiency at the shard size (launch overhead grows as the shard shrinks); the 2.68 MB "
                   "all-reduce is launch-latency sized",
    }
    del wl
    torch.cuda.empty_cache()
    return rec


def run_sub_eval(args, dev):
    """BASELINE config 4 (rank 0 only): 384 x 384 eval render (rgb / depth / normal) in the reference's 1024-ray chunks
    (eval.py:105-120, split_n_pixels) and in one call; the marching-cubes SDF pyramid of one 512^3 crop
    (plots.py:131-194: 64^3 dense, finer levels masked) through mesh.sdf_volume_pyramid, and a dense 2^24-point query."""
    from monosdf_b200 import confs, mesh
    from monosdf_b200.model.network import MonoSDFNetwork
    torch.manual_seed(0)
    model = MonoSDFNetwork(confs.to_conf(confs.SCANNET_MLP)).to(dev).eval()
    with torch.no_grad():
        model.density.beta.fill_(args.beta)
    model.set_precision(args.precision)
    R = 384
    ys, xs = torch.meshgrid(torch.arange(R), torch.arange(R), indexing="ij")
    uv = torch.stack([xs, ys], -1).reshape(1, -1, 2).float().to(dev)
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 300.0
    K[0, 2] = K[1, 2] = R / 2
    pose = torch.eye(4)
    pose[2, 3] = -0.3
    inp = {"uv": uv, "intrinsics": K[None].to(dev), "pose": pose[None].to(dev)}
    idx = torch.zeros(1, dtype=torch.long, device=dev)

    def render(chunk):
        outs = []
        with torch.no_grad():
            for s in range(0, R * R, chunk):
                o = model(dict(inp, uv=uv[:, s:s + chunk]), idx)
                outs.append((o["rgb_values"], o["depth_values"], o["normal_map"]))
        return [torch.cat(t, 0) for t in zip(*outs)]

    def t_of(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_chunks = t_of(lambda: render(1024), 2)
    ms_one = t_of(lambda: render(R * R), 3)
    stats = []
    sdf = model.implicit_network.get_sdf_vals
    ms_pyr = t_of(lambda: mesh.sdf_volume_pyramid(sdf, (-1.1,) * 3, (1.1,) * 3, 512, device=dev, stats=stats), 2)
    queries = sum(c for _, c in stats[-4:])
    dense = (torch.rand(1 << 24, 3, device=dev) * 2 - 1)
    with torch.no_grad():
        ms_dense = t_of(lambda: sdf(dense), 3)
    return {
        "workload": "eval.py-shaped: 384x384 render (147 456 rays, rgb/depth/normal, eval-mode sampler) + SDF pyramid of one 512^3 crop, MLP conf",
        "render_ms_1024_ray_chunks": ms_chunks, "render_ms_one_call": ms_one, "render_rays_per_s_one_call": R * R / ms_one * 1e3,
        "sdf_pyramid_ms": ms_pyr, "sdf_pyramid_queries": queries, "sdf_queries_per_s": (1 << 24) / ms_dense * 1e3,
        "sdf_only_algorithmic_tflops": 2.0 * roofline.WORK_MLP["A"] * (1 << 24) / ms_dense / 1e9,
        "dense_512cube_s_at_that_rate": 512 ** 3 / ((1 << 24) / ms_dense * 1e3),
        "limiter": "1024-ray chunks: host launch overhead (one model call per chunk); one call / SDF queries: the tensor pipe "
                   "of the fused sdf kernel and the per-layer render sweeps",
    }


def run_ours(args):
    import torch.distributed as dist
    from monosdf_b200 import _lib, training

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    timed = make_timer(world, dev)

    if args.total_rays:               # strong scaling: rank r renders rays [r N / W, (r + 1) N / W) of the step
        assert args.total_rays % world == 0, "equal shards: each rank's loss is a mean over its own rays"
        lo, hi = training.shard_range(args.total_rays, rank, world)
        args.rays = hi - lo
    n = args.rays
    wl = Workload(args.config, n, args.precision, args.beta, rank, world, dev)

    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation, first query) stalls kernel launches for
    # a few hundred ms, which used to land in the first timed steps; only samples taken inside the timed regions are reported
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):             # Adam runs un-restored here: the weights the timed steps start from are `warm` steps old
        wl.step_resident()
    wl.snapshot()
    wl.rounds = []
    clocks.mark()
    l0 = _lib.launch_count()
    ms = timed(wl.step_resident, args.steps)
    launches = _lib.launch_count() - l0
    rounds_list = list(wl.rounds)
    rounds = mean_rounds(rounds_list)

    for _ in range(2):      # untimed: the first end-to-end steps allocate the per-step input tensors (one-time cudaMalloc)
        wl.step_e2e()
    ms_e2e = timed(wl.step_e2e, args.steps)
    clocks.mark_end()
    clk = clocks.stop() if rank == 0 else None

    # instrumented steps of the SAME stationary workload
    gemm_cls = 1 if args.precision == "bf16" else 0
    prof, t_prof = profile_classes(wl, timed, min(3, args.steps))
    g_ms, g_flops, g_n, g_bytes = prof[gemm_cls]

    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_bw = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json: bf16_tflops_sustained, hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)"

    sub = {}
    if not args.no_sub_records and args.config == "mlp" and not args.total_rays and args.precision == "bf16":
        del wl.snap
        wl.snap = None
        torch.cuda.empty_cache()
        _lib.saved_pool.clear()
        sub["grid_conf_weak_32768"] = run_sub_grid(args, rank, world, dev, timed, peak_tf)
        _lib.saved_pool.clear()
        sub["strong_262144"] = run_sub_strong(args, rank, world, dev, timed, peak_tf)
        _lib.saved_pool.clear()
        if rank == 0:
            sub["eval_384x384_and_sdf_grid"] = run_sub_eval(args, dev)
        if world > 1:
            dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    alg_flops_step = gflop_per_ray(args.config, rounds) * 1e9 * n              # per GPU and step
    achieved = (alg_flops_step / 1e12) / (g_ms / 1e3) if g_ms > 0 else 0.0
    executed = (g_flops / 1e12) / (g_ms / 1e3) if g_ms > 0 else 0.0
    gbs = (g_bytes / 1e9) / (g_ms / 1e3) if g_ms > 0 else 0.0
    rays_per_step = args.total_rays if args.total_rays else world * n
    value = rays_per_step * args.steps / (ms / 1e3)
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram bytes / launch of the same kernels from ncu --set full
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get("dominant_kernel_dram_bytes_per_launch")
    others = {}
    for cls, name, bound in ((2, "hash grid gather/scatter", "hbm"), (3, "sampler round (warp scans)", "issue"),
                             (4, "compositing fwd+bwd", "hbm")):
        k_ms, k_work, k_n, k_bytes = prof[cls]
        if k_n:
            others[name] = {"bound": bound, "achieved": (k_bytes / 1e9) / (k_ms / 1e3), "peak": peak_bw, "unit": "GB/s",
                            "frac": (k_bytes / 1e9) / (k_ms / 1e3) / peak_bw, "launches": k_n, "ms_per_step": k_ms}
            if bound == "issue":
                others[name]["note"] = ("issue/SFU-bound by design (ncu: issue slots 72 % busy, XU 24 %, DRAM < 1 %): the byte "
                                        "figure is reported for completeness, not as its roofline")
    # every tcgen05 kernel against the roofline that bounds IT: the fused sdf-only network against the tensor peak, everything
    # that streams saved activations (training forward, sweeps, weight gradients) against the measured copy bandwidth
    kernels = {}
    for k, name, bound in ((1, "k_fused_sdf (sdf-only passes of the sampler)", "tensor"),
                           (2, "k_fused_sdf (forward sweep of the step, activations saved)", "hbm"),
                           (3, "k_tc_chain (reverse sweep)", "hbm"), (4, "k_tc_stream (tangent / backward layers)", "hbm"),
                           (5, "k_tc_gemm (colour network layers)", "hbm"), (6, "k_tc_wgrad (weight gradients)", "hbm")):
        k_ms, k_flops, k_n, k_bytes = prof.get(1 | (k << 8), (0.0, 0.0, 0, 0.0))
        if not k_n or k_ms <= 0:
            continue
        if bound == "tensor":
            ach, pk, unit = k_flops / 1e12 / (k_ms / 1e3), peak_tf, "TFLOP/s"
        else:
            ach, pk, unit = k_bytes / 1e9 / (k_ms / 1e3), peak_bw, "GB/s"
        kernels[name] = {"bound": bound, "achieved": ach, "peak": pk, "unit": unit, "frac": ach / pk if pk else None,
                         "launches": k_n, "ms_per_step": k_ms, "share_of_step": k_ms / t_prof if t_prof > 0 else None}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_rays else "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args.config, n, args.total_rays, args.beta), "rays_per_step_per_gpu": n,
                   "sampler_rounds": rounds, "sampler_rounds_per_step": rounds_list, "micro_batches": len(wl.micro),
                   "stationary": "parameters and Adam state restored from a snapshot before every step (%d-byte copies)" % (3 * wl.grad_bytes),
                   "precision_mode": args.precision, "parallelism": "ray-sharded dp%d, one NCCL all-reduce of the flat gradient arena" % world,
                   "l2_policy": "inputs larger than L2: every field chunk streams %.1f GB of activations through HBM (L2 is 126 MB)"
                                % (min(n * 98, 1060864) * 9.9e3 / 1e9)},
        "clocks": clk,
        "e2e": {"value": rays_per_step * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor",
                     "kernel": "k_gemm (fp32 SIMT MLP sweeps)" if gemm_cls == 0 else
                               "tcgen05 kernels: k_fused_sdf (sampler passes, training forward) + k_tc_stream / k_tc_gemm (per-layer sweeps) + k_tc_wgrad",
                     # achieved = ALGORITHMIC FLOPs of one step (SURVEY 8d per-ray figure at the measured sampler rounds x rays)
                     # / the time one step spends in those kernels (CUDA events per launch, mean of the instrumented steps)
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                     "traffic": traffic, "traffic_note": "ncu dram bytes per launch of the class's largest kernel (k_tc_wgrad, both products of "
                                                          "a layer, 262144 rows; algorithmic 537 MB); per-kernel table: profiles/ncu_traffic.json",
                     "peak_source": peak_src,
                     "algorithmic_gflop_per_ray": gflop_per_ray(args.config, rounds), "rays_per_step_per_gpu": n,
                     "launches": g_n, "kernel_ms_per_step": g_ms, "instrumented_step_ms": t_prof,
                     "kernel_share_of_step": g_ms / t_prof if t_prof > 0 else None,
                     "executed_tflops": executed, "executed_note": "2MNK of the padded tiles actually issued / same kernel time",
                     "step_algorithmic_tflops": (alg_flops_step / 1e12) / (ms / args.steps / 1e3),
                     "step_frac": (alg_flops_step / 1e12) / (ms / args.steps / 1e3) / peak_tf,
                     # the bound that applies to a per-layer GEMM moving 1-2.5 KB per 131 KFLOP: algorithmic bytes / time
                     "hbm_view": {"achieved": gbs, "peak": peak_bw, "unit": "GB/s", "frac": gbs / peak_bw if peak_bw else None},
                     "tcgen05_kernels": kernels,
                     "tcgen05_kernels_note": "executed 2MNK (fused sdf-only) or algorithmic bytes (the rest) of each kernel's launches / "
                                             "their CUDA-event time in the instrumented steps, against the roofline that bounds that kernel",
                     "other_kernels": others},
    }
    if sub:
        line["sub_records"] = sub
    if world == 1 and not args.no_cpu_baseline:
        # the CPU arm in a process of its own (no GPU visible): a bounded 1024-ray sample, all host cores
        try:
            outp = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                                   "--cpu-rays", str(CPU_SAMPLE_RAYS), "--cpu-threads", str(os.cpu_count() or 1), "--beta", str(args.beta)],
                                  capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0"))
            ref = json.loads(outp.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:      # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The contract's ONE JSON line, on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse()
    # stdout carries the JSON line only: whatever libraries print there (NCCL's version banner under torchrun, the
    # reference's constructors and loss) goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
