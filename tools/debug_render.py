"""Debug helper: prints the error of every intermediate of the render path against the oracle (GPU box)."""
import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import port
from tests.helpers import build_model, params_of, rel_err
from monosdf_b200.model.network import _Composite, _Field
from monosdf_b200 import _lib

DEV = "cuda"
case = sys.argv[1] if len(sys.argv) > 1 else "mlp_small"
fx = torch.load(os.path.join("tests/golden", case + ".pt"), map_location="cpu", weights_only=False)
model = build_model(fx, DEV)
if len(sys.argv) > 2:
    model.set_precision(sys.argv[2])
cfg = port.cfg_from_conf(fx["conf"])
n_rays, S = 96, 40
rays = port.synthetic_rays(n_rays, seed=9)
g = torch.Generator().manual_seed(8)
z = torch.sort(torch.rand(n_rays, S, generator=g) * 2.0, -1)[0]
o, d = rays["ray_cam_loc"], rays["ray_dirs"]
pose = torch.eye(4)[None].repeat(n_rays, 1, 1)
pose[:, :3, :3] = torch.linalg.qr(torch.randn(n_rays, 3, 3, generator=g))[0]
w_rgb, w_dep, w_nrm = torch.randn(n_rays, 3, generator=g), torch.randn(n_rays, 1, generator=g), torch.randn(n_rays, 3, generator=g)
w_w = torch.randn(n_rays, S, generator=g) * 0.1
params = params_of(model, requires_grad=True)
pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
dirs = d.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
sdf_o, feat_o, grad_o = port.sdf_outputs(params, cfg, pts)
rgb_o = port.color_net_forward(params, cfg, pts, grad_o, dirs, feat_o, torch.zeros(n_rays, dtype=torch.long), True)["rgb"].reshape(-1, S, 3)
beta_o = port.get_beta(params, cfg)
w_o = port.render_weights(z, sdf_o, beta_o)
rgbv_o = (w_o.unsqueeze(-1) * rgb_o).sum(1)
dep_o = rays["ray_dirs_tmp"][:, 2:] * ((w_o * z).sum(1, keepdim=True) / (w_o.sum(1, keepdim=True) + 1e-8))
nrm = grad_o / (grad_o.norm(2, -1, keepdim=True) + 1e-6)
nm_o = (w_o.unsqueeze(-1) * nrm.reshape(-1, S, 3)).sum(1)
nm_o = (pose[:, :3, :3].transpose(1, 2) @ nm_o.unsqueeze(-1)).squeeze(-1)

inet = model.implicit_network
zc, oc, dc = z.to(DEV), o.to(DEV), d.to(DEV)
points = torch.empty(n_rays * S, 3, device=DEV)
_lib.call("msdf_ray_points", _lib.ptr(oc), _lib.ptr(dc), _lib.ptr(zc), n_rays, S, _lib.ptr(points), _lib.stream())
print("points", rel_err(points, pts))
sdf, grad, _, rgb = _Field.apply(model._render_spec, "render", inet.sdf_bounding_sphere, inet.sphere_scale, S, points, dc, None,
                                 None, None, *inet._flat_weights(), *model.rendering_network._flat_weights())
print("sdf", rel_err(sdf, sdf_o), "grad", rel_err(grad, grad_o), "rgb", rel_err(rgb, rgb_o.reshape(-1, 3)))
print("rgb head rows", rgb[:3].tolist(), rgb_o.reshape(-1, 3)[:3].tolist())
tmp = rays["ray_dirs_tmp"].to(DEV)
# composite on ORACLE inputs first (isolates the compositing kernels)
sd_in = sdf_o.detach().reshape(n_rays, S).to(DEV).requires_grad_(True)
rgb_in = rgb_o.detach().reshape(-1, 3).to(DEV).requires_grad_(True)
gr_in = grad_o.detach().to(DEV).requires_grad_(True)
beta_in = model.density.get_beta()
wts, rgbv, dep, nm = _Composite.apply(zc, sd_in, rgb_in, gr_in, beta_in, tmp[:, 2:], 3, pose.to(DEV), 1, False, model.bg_color)
print("composite(oracle inputs): weights", rel_err(wts, w_o), "rgbv", rel_err(rgbv, rgbv_o), "depth", rel_err(dep, dep_o), "normal", rel_err(nm, nm_o))
# composite backward vs autograd on oracle
sd_l = sdf_o.detach().clone().requires_grad_(True); rgb_l = rgb_o.detach().clone().requires_grad_(True); gr_l = grad_o.detach().clone().requires_grad_(True)
bp = params["density.beta"].detach().clone().requires_grad_(True)
b_l = bp.abs() + cfg.beta_min
w2 = port.render_weights(z, sd_l, b_l)
rgbv2 = (w2.unsqueeze(-1) * rgb_l).sum(1)
dep2 = rays["ray_dirs_tmp"][:, 2:] * ((w2 * z).sum(1, keepdim=True) / (w2.sum(1, keepdim=True) + 1e-8))
nrm2 = gr_l / (gr_l.norm(2, -1, keepdim=True) + 1e-6)
nm2 = (w2.unsqueeze(-1) * nrm2.reshape(-1, S, 3)).sum(1)
nm2 = (pose[:, :3, :3].transpose(1, 2) @ nm2.unsqueeze(-1)).squeeze(-1)
((rgbv2 * w_rgb).sum() + (dep2 * w_dep).sum() + (nm2 * w_nrm).sum() + (w2 * w_w).sum()).backward()
model.zero_grad()
((rgbv * w_rgb.to(DEV)).sum() + (dep * w_dep.to(DEV)).sum() + (nm * w_nrm.to(DEV)).sum() + (wts * w_w.to(DEV)).sum()).backward()
print("composite bwd: d_sdf", rel_err(sd_in.grad.reshape(-1, 1), sd_l.grad), "d_rgb", rel_err(rgb_in.grad, rgb_l.grad.reshape(-1, 3)),
      "d_grad", rel_err(gr_in.grad, gr_l.grad), "d_beta", float(model.density.beta.grad), float(bp.grad))
# full chain grads
((rgbv_o * w_rgb).sum() + (dep_o * w_dep).sum() + (nm_o * w_nrm).sum() + (w_o * w_w).sum()).backward()
model.zero_grad()
weights, rgbv, dep, nm = _Composite.apply(zc, sdf.reshape(n_rays, S), rgb, grad, model.density.get_beta(), tmp[:, 2:], 3,
                                          pose.to(DEV), 1, False, model.bg_color)
print("full: weights", rel_err(weights, w_o), "rgbv", rel_err(rgbv, rgbv_o), "depth", rel_err(dep, dep_o), "normal", rel_err(nm, nm_o))
((rgbv * w_rgb.to(DEV)).sum() + (dep * w_dep.to(DEV)).sum() + (nm * w_nrm.to(DEV)).sum() + (weights * w_w.to(DEV)).sum()).backward()
for k, p in model.named_parameters():
    if p.grad is None or params[k].grad is None:
        print(k, "grad missing", p.grad is None, params[k].grad is None)
        continue
    print("%-45s %.3e  (|g|max %.3e)" % (k, rel_err(p.grad, params[k].grad), float(params[k].grad.abs().max())))
