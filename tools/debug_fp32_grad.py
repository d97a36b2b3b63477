"""Per-parameter gradient error of the fp32 mode: ours vs the oracle in fp32 and in fp64 on identical sample positions."""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import port
from tests.helpers import build_model, params_of

case = sys.argv[1] if len(sys.argv) > 1 else "mlp_full"
fx = torch.load(os.path.join("tests", "golden", case + ".pt"), map_location="cpu", weights_only=False)
n = fx["n_rays"]
model = build_model(fx, "cuda").train()
model.rng = "reference"
rays, gt = port.synthetic_rays(n, seed=1), port.synthetic_gt(n, seed=2)
torch.manual_seed(fx["train_seed"])
out = model({k: v.cuda() for k, v in rays.items()}, torch.zeros(n, dtype=torch.long, device="cuda"), if_pixel_input=True)
loss = port.monosdf_loss(out, {k: v.cuda() for k, v in gt.items()})
loss["loss"].backward()
cfg = port.cfg_from_conf(fx["conf"], fx.get("if_hdr", False))
res = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    params = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in params_of(model).items()}
    for v in params.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    r = {k: v.to(dt) for k, v in rays.items()}
    g = {k: v.to(dt) for k, v in gt.items()}
    torch.manual_seed(fx["train_seed"])
    o = port.model_forward(params, cfg, r, torch.zeros(n, dtype=torch.long), if_pixel_input=True, training=True,
                           eik_points=model._last_eikonal_points.cpu().to(dt), z_vals=out["z_vals"].detach().cpu().to(dt))
    port.monosdf_loss(o, g)["loss"].backward()
    res[name] = ({k: v.grad for k, v in params.items() if v.is_floating_point() and v.grad is not None}, o)


def err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


print("%-46s %10s %10s %10s   |grad|max" % ("parameter", "ours/f64", "port32/f64", "ours/port32"))
for k, p in model.named_parameters():
    if k not in res["f64"][0]:
        continue
    g64, g32 = res["f64"][0][k], res["f32"][0][k]
    print("%-46s %10.2e %10.2e %10.2e   %.3e" % (k, err(p.grad, g64), err(g32, g64), err(p.grad, g32), float(g64.abs().max())))
for k in ["rgb_values", "depth_values", "normal_map", "sdf", "grad_theta"]:
    print("out %-20s ours/f64 %.2e  port32/f64 %.2e" % (k, err(out[k], res["f64"][1][k]), err(res["f32"][1][k], res["f64"][1][k])))
