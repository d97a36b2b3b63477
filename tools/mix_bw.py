"""DRAM throughput of plain elementwise torch kernels with the read/write mixes of the sweep kernels (how much of the
copy bandwidth is reachable at all with 2-3 input streams?):  python tools/mix_bw.py"""
import torch
dev = "cuda"
n = 262144 * 256 * 4          # 4 chunks' worth, far beyond L2
a, b, c, d = [torch.randn(n, device=dev, dtype=torch.bfloat16) for _ in range(4)]
o1, o2 = torch.empty_like(a), torch.empty_like(a)


def bw(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12


B = n * 2
print("1R 1W  copy           %.2f TB/s" % bw(lambda: o1.copy_(a), 2 * B))
print("2R 1W  add            %.2f TB/s" % bw(lambda: torch.add(a, b, out=o1), 3 * B))
print("3R 1W  addcmul        %.2f TB/s" % bw(lambda: torch.addcmul(a, b, c, out=o1), 4 * B))
print("1R 0W  sum            %.2f TB/s" % bw(lambda: a.sum(), B))
print("0R 1W  fill           %.2f TB/s" % bw(lambda: o1.fill_(1.0), B))
