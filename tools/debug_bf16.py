"""Debug helper (GPU box): bf16 tensor-core mode against fp32 mode of the same model, max-norm and rms errors."""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import build_model

DEV = "cuda"
case = sys.argv[1] if len(sys.argv) > 1 else "mlp_full"
fx = torch.load(os.path.join("tests/golden", case + ".pt"), map_location="cpu", weights_only=False)
model = build_model(fx, DEV)
inet = model.implicit_network


def err(a, b):
    a, b = a.double(), b.double()
    return "max %.2e rms %.2e" % (float((a - b).abs().max() / b.abs().max()), float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()))


n = 4000
g = torch.Generator().manual_seed(3)
x = ((torch.rand(n, 3, generator=g) * 2 - 1) * 0.6).to(DEV)
w = torch.randn(n, 3, generator=g).to(DEV)
ws = torch.randn(n, 1, generator=g).to(DEV)
res = {}
for mode in ("fp32", "bf16"):
    model.set_precision(mode)
    model.zero_grad()
    sdf = inet.get_sdf_vals(x)
    grad = inet.gradient_sdf(x)
    (grad * w).sum().backward()
    g_grad = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad()
    sdf2 = inet.get_sdf_vals(x)
    (sdf2 * ws).sum().backward()
    g_sdf = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    res[mode] = (sdf.detach(), grad.detach(), g_grad, g_sdf)
print("sdf      ", err(res["bf16"][0], res["fp32"][0]))
print("grad_x   ", err(res["bf16"][1], res["fp32"][1]))
bad = ((res["bf16"][1] - res["fp32"][1]).abs().max(-1)[0] > 0.05).float().mean()
print("fraction of points with |d grad| > 0.05:", float(bad))
for k in res["fp32"][2]:
    print("dL(grad)/d %-38s %s" % (k, err(res["bf16"][2][k], res["fp32"][2][k])))
for k in res["fp32"][3]:
    print("dL(sdf)/d  %-38s %s" % (k, err(res["bf16"][3][k], res["fp32"][3][k])))
