"""CPU oracle: a plain-PyTorch restatement of MonoSDF's volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under monosdf_b200/ imports this module; it is
called only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` arm, always as the checker or the timed CPU baseline, never
as the product path.

Parity status: PINNED.  tests/test_oracle_golden.py and the
committed fixtures under tests/golden/ (made by oracle/make_golden.py from the
unmodified reference at /root/reference) check every function here against the
reference's own outputs on identical weights, rays and random draws.

Each function cites the reference code it restates (paths relative to
/root/reference/code).  The restatement is functional (explicit parameter
dicts keyed like the reference state_dict) and, like the reference, obtains
grad_x(sdf) with autograd (create_graph=True), so that timing it on host cores
is a fair stand-in for the reference's CPU path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (shapes only; mirrors the `model{}` block of confs/*.conf)
# --------------------------------------------------------------------------------------
@dataclass
class SdfNetCfg:
    d_in: int = 3
    d_out: int = 1
    dims: List[int] = field(default_factory=lambda: [256] * 8)
    skip_in: Tuple[int, ...] = (4,)
    multires: int = 6
    weight_norm: bool = True
    sphere_scale: float = 1.0
    # ImplicitNetworkGrid only
    grid: bool = False
    use_grid_feature: bool = True
    divide_factor: float = 1.5
    num_levels: int = 16
    level_dim: int = 2
    base_size: int = 16
    end_size: int = 2048
    logmap: int = 19


@dataclass
class ColorNetCfg:
    mode: str = "idr"
    d_in: int = 9
    d_out: int = 3
    dims: List[int] = field(default_factory=lambda: [256, 256])
    weight_norm: bool = True
    multires_view: int = 4
    per_image_code: bool = False
    spec: bool = False


@dataclass
class SamplerCfg:
    near: float = 0.0
    N_samples: int = 64
    N_samples_eval: int = 128
    N_samples_extra: int = 32
    eps: float = 0.1
    beta_iters: int = 10
    max_total_iters: int = 5
    add_tiny: float = 1.0e-6


@dataclass
class ModelCfg:
    feature_vector_size: int = 256
    scene_bounding_sphere: float = 1.1
    white_bkgd: bool = False
    bg_color: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    beta_min: float = 1e-4
    if_hdr: bool = False
    sdf: SdfNetCfg = field(default_factory=SdfNetCfg)
    color: ColorNetCfg = field(default_factory=ColorNetCfg)
    sampler: SamplerCfg = field(default_factory=SamplerCfg)


def cfg_from_conf(conf: dict, if_hdr: bool = False) -> ModelCfg:
    """Build a ModelCfg from a (nested dict) `model{}` conf block (network.py:481-499)."""
    inet = dict(conf["implicit_network"])
    grid = bool(conf.get("Grid_MLP", False))
    sdf = SdfNetCfg(
        d_in=inet.get("d_in", 3), d_out=inet.get("d_out", 1), dims=list(inet["dims"]),
        skip_in=tuple(inet.get("skip_in", ())), multires=inet.get("multires", 0),
        weight_norm=inet.get("weight_norm", True), sphere_scale=inet.get("sphere_scale", 1.0),
        grid=grid, use_grid_feature=inet.get("use_grid_feature", True),
        divide_factor=inet.get("divide_factor", 1.5), num_levels=inet.get("num_levels", 16),
        level_dim=inet.get("level_dim", 2), base_size=inet.get("base_size", 16),
        end_size=inet.get("end_size", 2048), logmap=inet.get("logmap", 19))
    rn = dict(conf["rendering_network"])
    color = ColorNetCfg(mode=rn.get("mode", "idr"), d_in=rn["d_in"], d_out=rn["d_out"], dims=list(rn["dims"]),
                        weight_norm=rn.get("weight_norm", True), multires_view=rn.get("multires_view", 0),
                        per_image_code=rn.get("per_image_code", False), spec=rn.get("spec", False))
    rs = dict(conf["ray_sampler"])
    samp = SamplerCfg(near=rs["near"], N_samples=rs["N_samples"], N_samples_eval=rs["N_samples_eval"],
                      N_samples_extra=rs["N_samples_extra"], eps=rs["eps"], beta_iters=rs["beta_iters"],
                      max_total_iters=rs["max_total_iters"], add_tiny=rs.get("add_tiny", 1.0e-6))
    dens = dict(conf.get("density", {}))
    return ModelCfg(feature_vector_size=conf["feature_vector_size"],
                    scene_bounding_sphere=conf.get("scene_bounding_sphere", 1.0),
                    white_bkgd=conf.get("white_bkgd", False), bg_color=tuple(conf.get("bg_color", (1.0, 1.0, 1.0))),
                    beta_min=dens.get("beta_min", 1e-4), if_hdr=if_hdr, sdf=sdf, color=color, sampler=samp)


# --------------------------------------------------------------------------------------
# embedder.py:5-50  (NeRF positional encoding, include_input, log-sampled bands)
# --------------------------------------------------------------------------------------
def positional_encoding(x: Tensor, multires: int) -> Tensor:
    if multires <= 0:
        return x
    bands = 2.0 ** torch.linspace(0.0, multires - 1, multires)  # embedder.py:22
    parts = [x]
    for f in bands:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


def pe_width(multires: int, d: int = 3) -> int:
    return d + 2 * d * multires if multires > 0 else d


# --------------------------------------------------------------------------------------
# nn.utils.weight_norm (dim=0): W[o,:] = g[o] * v[o,:] / ||v[o,:]||   (network.py:72-73)
# --------------------------------------------------------------------------------------
def linear_weight(params: Dict[str, Tensor], prefix: str) -> Tuple[Tensor, Tensor]:
    if prefix + ".weight_g" in params:
        g, v = params[prefix + ".weight_g"], params[prefix + ".weight_v"]
        w = v * (g / v.norm(2, dim=1, keepdim=True))
    else:
        w = params[prefix + ".weight"]
    return w, params[prefix + ".bias"]


def softplus100(x: Tensor) -> Tensor:
    return F.softplus(x, beta=100)  # network.py:77 (threshold 20)


# --------------------------------------------------------------------------------------
# hash grid: hashgrid.py:108-166 and hashencoder.cu:35-93,104-254 (forward + dy_dx)
# --------------------------------------------------------------------------------------
HASH_PRIMES = (1, 2654435761, 805459861)


def hash_offsets(num_levels=16, base_resolution=16, desired_resolution=2048, log2_hashmap_size=19, input_dim=3):
    """hashgrid.py:112-136 -> (offsets int32[L+1], per_level_scale float64)."""
    per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    max_params = 2 ** log2_hashmap_size
    offs, off = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        offs.append(off)
        off += min(max_params, res ** input_dim)
    offs.append(off)
    return np.array(offs, dtype=np.int32), per_level_scale


def _level_geometry(level: int, S: float, H: int):
    """hashencoder.cu:152-153; exp2f in fp32 with S narrowed to float."""
    scale = np.float32(np.exp2(np.float32(level) * np.float32(S))) * np.float32(H) - np.float32(1.0)
    scale = np.float32(scale)
    res = int(np.ceil(scale)) + 1
    return scale, res


def _grid_index(pg: Tensor, hashmap_size: int, res: int) -> Tensor:
    """hashencoder.cu:54-72 (per-lookup dense-or-hash decision by stride overflow). pg: [...,3] int64 >= 0."""
    M32 = 0xFFFFFFFF
    stride, index, d = 1, torch.zeros_like(pg[..., 0]), 0
    while d < 3 and stride <= hashmap_size:
        index = (index + pg[..., d] * stride) & M32
        stride = (stride * res) & M32
        d += 1
    if stride > hashmap_size:
        index = torch.zeros_like(index)
        for k in range(3):
            index = index ^ ((pg[..., k] * HASH_PRIMES[k]) & M32)
    return index % hashmap_size


def hash_encode(x01: Tensor, emb: Tensor, offsets: np.ndarray, per_level_scale: float, H: int = 16,
                want_dy_dx: bool = False):
    """Forward of kernel_grid (hashencoder.cu:104-254) in differentiable torch ops.

    x01: [B,3] in [0,1]; emb: [sO,C]. Returns feats [B, L*C] (level-major, channel-minor,
    hashgrid.py:44) and optionally dy_dx [B, L, 3, C] computed exactly as :210-253.
    Differentiable wrt emb (and x01, but see the note on dropped terms in hashgrid.py:101).
    """
    B = x01.shape[0]
    L = len(offsets) - 1
    C = emb.shape[1]
    S = np.log2(per_level_scale)
    oob = ((x01 < 0) | (x01 > 1)).any(-1, keepdim=True)  # :125-149
    feats, dydx = [], []
    for l in range(L):
        hs = int(offsets[l + 1] - offsets[l])
        scale, res = _level_geometry(l, S, H)
        table = emb[int(offsets[l]):int(offsets[l + 1])]
        pos = x01 * float(scale)
        pg = torch.floor(pos)
        fr = pos - pg
        pgi = pg.to(torch.int64).clamp_min(0)
        w1 = fr * fr * (3.0 - 2.0 * fr)          # smoothstep :87-89
        dw = 6.0 * fr * (1.0 - fr)               # :91-93
        w0 = 1.0 - w1
        out = torch.zeros(B, C, dtype=x01.dtype)
        corner_vals = {}
        for idx in range(8):
            offs = torch.tensor([(idx >> d) & 1 for d in range(3)], dtype=torch.int64)
            w = torch.ones(B, dtype=x01.dtype)
            for d in range(3):
                w = w * (w1[:, d] if (idx >> d) & 1 else w0[:, d])
            gi = _grid_index(pgi + offs, hs, res)
            v = table[gi]
            corner_vals[idx] = v
            out = out + w[:, None] * v
        feats.append(torch.where(oob, torch.zeros_like(out), out))
        if want_dy_dx:
            per_d = []
            for gd in range(3):
                acc = torch.zeros(B, C, dtype=x01.dtype)
                others = [d for d in range(3) if d != gd]
                for sub in range(4):
                    w = torch.full((B,), float(scale), dtype=x01.dtype)
                    base = 0
                    for nd, d in enumerate(others):
                        bit = (sub >> nd) & 1
                        w = w * (w1[:, d] if bit else w0[:, d])
                        base |= bit << d
                    left, right = corner_vals[base], corner_vals[base | (1 << gd)]
                    acc = acc + w[:, None] * (right - left) * dw[:, gd:gd + 1]
                per_d.append(torch.where(oob, torch.zeros_like(acc), acc))
            dydx.append(torch.stack(per_d, 1))  # [B,3,C]
    feats = torch.cat(feats, -1)
    if want_dy_dx:
        return feats, torch.stack(dydx, 1)  # [B,L,3,C]
    return feats


# --------------------------------------------------------------------------------------
# ImplicitNetwork / ImplicitNetworkGrid forward  (network.py:79-96, 247-275)
# --------------------------------------------------------------------------------------
def sdf_net_dims(cfg: ModelCfg) -> List[int]:
    s = cfg.sdf
    d0 = pe_width(s.multires, s.d_in)
    if s.grid:
        d0 += s.num_levels * s.level_dim
    return [d0] + list(s.dims) + [s.d_out + cfg.feature_vector_size]


def sdf_net_forward(params: Dict[str, Tensor], cfg: ModelCfg, x: Tensor, prefix="implicit_network") -> Tensor:
    s = cfg.sdf
    if s.grid:
        if s.use_grid_feature:
            offsets = params[prefix + ".encoding.offsets"].cpu().numpy()
            _, pls = hash_offsets(s.num_levels, s.base_size, s.end_size, s.logmap)
            xin = (x / s.divide_factor + 1.0) / 2.0   # network.py:250, hashgrid.py:158
            feat = hash_encode(xin, params[prefix + ".encoding.embeddings"], offsets, pls, s.base_size)
        else:
            feat = torch.zeros(x.shape[0], s.num_levels * s.level_dim, dtype=x.dtype)  # network.py:252
        inp = torch.cat([positional_encoding(x, s.multires), feat], -1)
    else:
        inp = positional_encoding(x, s.multires)
    dims = sdf_net_dims(cfg)
    n_lin = len(dims) - 1
    h = inp
    for l in range(n_lin):
        w, b = linear_weight(params, f"{prefix}.lin{l}")
        if l in s.skip_in:
            h = torch.cat([h, inp], 1) / np.sqrt(2)   # network.py:88-89
        h = F.linear(h, w, b)
        if l < n_lin - 1:
            h = softplus100(h)
    return h


def _clamp_sdf(cfg: ModelCfg, x: Tensor, sdf: Tensor) -> Tensor:
    """network.py:116-118,134-136; plain net only, disabled for white_bkgd (network.py:492)."""
    R = 0.0 if cfg.white_bkgd else cfg.scene_bounding_sphere
    if (not cfg.sdf.grid) and R > 0.0:
        sphere = cfg.sdf.sphere_scale * (R - x.norm(2, 1, keepdim=True))
        sdf = torch.minimum(sdf, sphere)
    return sdf


def sdf_vals(params, cfg: ModelCfg, x: Tensor) -> Tensor:
    """get_sdf_vals (network.py:131-137, 307-309)."""
    return _clamp_sdf(cfg, x, sdf_net_forward(params, cfg, x)[:, :1])


def sdf_outputs(params, cfg: ModelCfg, x: Tensor, create_graph: bool = True):
    """get_outputs (network.py:111-129, 290-305): sdf, features, grad_x sdf via autograd."""
    x = x.detach().requires_grad_(True)
    out = sdf_net_forward(params, cfg, x)
    sdf = _clamp_sdf(cfg, x, out[:, :1])
    feat = out[:, 1:]
    (g,) = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=create_graph, retain_graph=True)
    return sdf, feat, g


def sdf_gradient(params, cfg: ModelCfg, x: Tensor, create_graph: bool = True) -> Tensor:
    """gradient_sdf (network.py:98-109, 277-288): never clamps."""
    x = x.detach().requires_grad_(True)
    y = sdf_net_forward(params, cfg, x)[:, :1]
    (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph, retain_graph=True)
    return g


# --------------------------------------------------------------------------------------
# RenderingNetwork.forward (network.py:389-470)
# --------------------------------------------------------------------------------------
def color_net_forward(params, cfg: ModelCfg, points, normals, view_dirs, feats, indices=None,
                      if_pixel_input=False, prefix="rendering_network"):
    c = cfg.color
    vd = positional_encoding(view_dirs, c.multires_view)
    if c.mode == "idr":
        x = torch.cat([points, vd, normals, feats], -1)
    elif c.mode == "nerf":
        x = torch.cat([vd, feats], -1)
    else:
        raise NotImplementedError
    if c.per_image_code:
        table = params[prefix + ".embeddings"]
        if not if_pixel_input:
            code = table[indices].expand(x.shape[0], -1)                       # :409
        else:
            ns = x.shape[0] // indices.shape[0]
            code = table[indices].unsqueeze(1).expand(-1, ns, -1).flatten(0, 1)  # :412
        x = torch.cat([x, code], -1)
    n_lin = len(c.dims) + 1
    if c.spec:  # :427-454 (diffuse/specular split, HDR only)
        for l in range(n_lin - 2):
            w, b = linear_weight(params, f"{prefix}.lin{l}")
            x = torch.relu(F.linear(x, w, b))
        diff, x = x[:, :3], x[:, 3:]
        for l in range(n_lin - 2, n_lin):
            w, b = linear_weight(params, f"{prefix}.lin{l}")
            x = torch.relu(F.linear(x, w, b))
        return {"rgb": diff + x, "rgb_diff": diff, "rgb_spec": x}
    for l in range(n_lin):
        w, b = linear_weight(params, f"{prefix}.lin{l}")
        x = F.linear(x, w, b)
        if l < n_lin - 1:
            x = torch.relu(x)
    x = torch.relu(x) if cfg.if_hdr else torch.sigmoid(x)  # :465-468
    return {"rgb": x}


# --------------------------------------------------------------------------------------
# LaplaceDensity (density.py:16-30) and volume_rendering (network.py:626-640)
# --------------------------------------------------------------------------------------
def get_beta(params, cfg: ModelCfg) -> Tensor:
    return params["density.beta"].abs() + torch.tensor(cfg.beta_min)


def laplace_density(sdf: Tensor, beta) -> Tensor:
    alpha = 1 / beta
    return alpha * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


def render_weights(z_vals: Tensor, sdf: Tensor, beta) -> Tensor:
    dens = laplace_density(sdf, beta).reshape(-1, z_vals.shape[1])
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    dists = torch.cat([dists, torch.full((dists.shape[0], 1), 1e10)], -1)
    fe = dists * dens
    sfe = torch.cat([torch.zeros(dists.shape[0], 1), fe[:, :-1]], -1)
    alpha = 1 - torch.exp(-fe)
    trans = torch.exp(-torch.cumsum(sfe, -1))
    return alpha * trans


# --------------------------------------------------------------------------------------
# ErrorBoundSampler (ray_sampler.py:48-83, 110-272)
# --------------------------------------------------------------------------------------
def near_far_from_cube(o: Tensor, d: Tensor, bound: float, near_min: float, far_max: float):
    tmin = (-bound - o) / (d + 1e-15)
    tmax = (bound - o) / (d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).max(-1, keepdim=True)[0]
    far = torch.where(tmin > tmax, tmin, tmax).min(-1, keepdim=True)[0]
    miss = far < near
    near = torch.where(miss, torch.full_like(near, 1e9), near)
    far = torch.where(miss, torch.full_like(far, 1e9), far)
    return near.clamp(min=near_min), far.clamp(max=far_max)


def error_bound(beta, sdf2d: Tensor, dists: Tensor, d_star: Tensor) -> Tensor:
    """get_error_bound (ray_sampler.py:264-272). beta: 0-d or [N,1]."""
    dens = laplace_density(sdf2d, beta)
    sfe = torch.cat([torch.zeros(dists.shape[0], 1), dists * dens[:, :-1]], -1)
    integral = torch.cumsum(sfe, -1)
    eps_sec = torch.exp(-d_star / beta) * (dists ** 2.0) / (4 * beta ** 2)
    eint = torch.cumsum(eps_sec, -1)
    bound = (torch.clamp(torch.exp(eint), max=1.0e6) - 1.0) * torch.exp(-integral[:, :-1])
    return bound.max(-1)[0]


def d_star_bound(z: Tensor, d: Tensor):
    """Theorem-1 bound per interval (ray_sampler.py:141-153). Returns (dists, d_star)."""
    dists = z[:, 1:] - z[:, :-1]
    a, b, c = dists, d[:, :-1].abs(), d[:, 1:].abs()
    first = a.pow(2) + b.pow(2) <= c.pow(2)
    second = a.pow(2) + c.pow(2) <= b.pow(2)
    ds = torch.zeros_like(dists)
    ds = torch.where(first, b, ds)
    ds = torch.where(second, c, ds)
    s = (a + b + c) / 2.0
    area = s * (s - a) * (s - b) * (s - c)
    mask = ~first & ~second & (b + c - a > 0)
    heron = (2.0 * torch.sqrt(area)) / a
    ds = torch.where(mask, heron, ds)
    ds = (d[:, 1:].sign() * d[:, :-1].sign() == 1) * ds
    return dists, ds


def inverse_cdf(cdf: Tensor, bins: Tensor, u: Tensor) -> Tensor:
    """ray_sampler.py:216-228."""
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = (inds - 1).clamp_min(0)
    above = inds.clamp_max(cdf.shape[-1] - 1)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = ca - cb
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cb) / denom
    return bb + t * (ba - bb)


def sampler_get_z_vals(cfg: ModelCfg, ray_dirs: Tensor, cam_loc: Tensor, sdf_fn: Callable[[Tensor], Tensor],
                       beta0: Tensor, training: bool, trace: Optional[dict] = None):
    """ErrorBoundSampler.get_z_vals (ray_sampler.py:110-262).

    Random draws (training only) come from torch's global CPU generator in the
    reference's order: rand[N,128] (:79), rand[N,64] (:213), randperm (:244),
    randint (:254).  `trace`, if given, records per-round tensors for tests.
    """
    sc = cfg.sampler
    R = cfg.scene_bounding_sphere
    far_const = 2.0 * R * 1.75                                      # :91
    N = ray_dirs.shape[0]
    # --- UniformSampler.get_z_vals (:63-83), take_sphere_intersection=True -> cube (:95)
    _, far = near_far_from_cube(cam_loc, ray_dirs, R, sc.near, far_const)
    near = sc.near * torch.ones(N, 1)
    t_vals = torch.linspace(0.0, 1.0, steps=sc.N_samples_eval)
    z_vals = near * (1.0 - t_vals) + far * t_vals
    if training:
        mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        z_vals = lower + (upper - lower) * torch.rand(z_vals.shape)
    samples, samples_idx = z_vals, None
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    bound = (1.0 / (4.0 * torch.log(torch.tensor(sc.eps + 1.0)))) * (dists ** 2.0).sum(-1)
    beta = torch.sqrt(bound)
    total_iters, not_converge = 0, True
    sdf = None
    while not_converge and total_iters < sc.max_total_iters:
        pts = (cam_loc.unsqueeze(1) + samples.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
        with torch.no_grad():
            new_sdf = sdf_fn(pts)
        if samples_idx is not None:
            merged = torch.cat([sdf.reshape(-1, z_vals.shape[1] - samples.shape[1]),
                                new_sdf.reshape(-1, samples.shape[1])], -1)
            sdf = torch.gather(merged, 1, samples_idx).reshape(-1, 1)
        else:
            sdf = new_sdf
        d = sdf.reshape(z_vals.shape)
        dists, d_star = d_star_bound(z_vals, d)
        # beta line search (:157-165); NB beta_max aliases beta in the reference
        err = error_bound(beta0, d, dists, d_star)
        beta = torch.where(err <= sc.eps, beta0.expand_as(beta), beta)
        beta_min, beta_max = beta0.unsqueeze(0).repeat(N), beta
        for _ in range(sc.beta_iters):
            mid = (beta_min + beta_max) / 2.0
            err = error_bound(mid.unsqueeze(-1), d, dists, d_star)
            beta_max = torch.where(err <= sc.eps, mid, beta_max)
            beta_min = torch.where(err > sc.eps, mid, beta_min)
        beta = beta_max
        dens = laplace_density(d, beta.unsqueeze(-1))
        dists_inf = torch.cat([dists, torch.full((N, 1), 1e10)], -1)
        fe = dists_inf * dens
        sfe = torch.cat([torch.zeros(N, 1), fe[:, :-1]], -1)
        alpha = 1 - torch.exp(-fe)
        trans = torch.exp(-torch.cumsum(sfe, -1))
        weights = alpha * trans
        total_iters += 1
        not_converge = bool(beta.max() > beta0)
        if trace is not None:
            trace.setdefault("rounds", []).append(
                dict(z=z_vals.clone(), sdf=d.clone(), d_star=d_star.clone(), beta=beta.clone(), weights=weights.clone()))
        upsample = not_converge and total_iters < sc.max_total_iters
        if upsample:
            n_new = sc.N_samples_eval
            eps_sec = torch.exp(-d_star / beta.unsqueeze(-1)) * (dists ** 2.0) / (4 * beta.unsqueeze(-1) ** 2)
            eint = torch.cumsum(eps_sec, -1)
            bound_op = (torch.clamp(torch.exp(eint), max=1.0e6) - 1.0) * trans[:, :-1]
            pdf = bound_op + sc.add_tiny
        else:
            n_new = sc.N_samples
            pdf = weights[..., :-1] + 1e-5
        pdf = pdf / torch.sum(pdf, -1, keepdim=True)
        cdf = torch.cumsum(pdf, -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        if upsample or not training:
            u = torch.linspace(0.0, 1.0, steps=n_new).unsqueeze(0).repeat(N, 1)
        else:
            u = torch.rand(N, n_new)
        samples = inverse_cdf(cdf, z_vals, u)
        if trace is not None:
            trace["rounds"][-1].update(cdf=cdf.clone(), samples=samples.clone(), upsample=upsample)
        if upsample:
            z_vals, samples_idx = torch.sort(torch.cat([z_vals, samples], -1), -1)
    z_samples = samples
    near = sc.near * torch.ones(N, 1)
    farc = far_const * torch.ones(N, 1)
    if sc.N_samples_extra > 0:
        if training:
            pick = torch.randperm(z_vals.shape[1])[:sc.N_samples_extra]
        else:
            pick = torch.linspace(0, z_vals.shape[1] - 1, sc.N_samples_extra).long()
        extra = torch.cat([near, farc, z_vals[:, pick]], -1)
    else:
        extra = torch.cat([near, farc], -1)
    z_final, _ = torch.sort(torch.cat([z_samples, extra], -1), -1)
    idx = torch.randint(z_final.shape[-1], (N,))
    z_eik = torch.gather(z_final, 1, idx.unsqueeze(-1))
    if trace is not None:
        trace.update(total_iters=total_iters, z_dense=z_vals.clone(), beta=beta.clone())
    return z_final, z_eik


# --------------------------------------------------------------------------------------
# rend_util.get_camera_params / lift (rend_util.py:63-91,105-118), matrix poses only
# --------------------------------------------------------------------------------------
def camera_rays(uv: Tensor, pose: Tensor, intrinsics: Tensor):
    cam_loc = pose[:, :3, 3]
    fx, fy = intrinsics[:, 0, 0:1], intrinsics[:, 1, 1:2]
    cx, cy, sk = intrinsics[:, 0, 2:3], intrinsics[:, 1, 2:3], intrinsics[:, 0, 1:2]
    x, y = uv[:, :, 0], uv[:, :, 1]
    z = torch.ones_like(x)
    xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    yl = (y - cy) / fy * z
    pc = torch.stack((xl, yl, z, torch.ones_like(z)), -1).permute(0, 2, 1)
    world = torch.bmm(pose, pc).permute(0, 2, 1)[:, :, :3]
    dirs = F.normalize(world - cam_loc[:, None, :], dim=2)
    return dirs, cam_loc


# --------------------------------------------------------------------------------------
# MonoSDFNetwork.forward (network.py:502-624)
# --------------------------------------------------------------------------------------
def model_forward(params: Dict[str, Tensor], cfg: ModelCfg, inp: Dict[str, Tensor], indices: Tensor,
                  if_pixel_input: bool = False, training: bool = False, trace: Optional[dict] = None,
                  eik_points: Optional[Tensor] = None, z_vals: Optional[Tensor] = None):
    """z_vals / eik_points: tests may inject the sample depths / eikonal points of the run under test, so that the
    field, compositing and loss are compared on identical sample positions (the sampler has its own bit-exact test)."""
    if not if_pixel_input:
        ray_dirs, cam_loc = camera_rays(inp["uv"], inp["pose"], inp["intrinsics"])
        ray_dirs_tmp, _ = camera_rays(inp["uv"], torch.eye(4)[None], inp["intrinsics"])
        cam_loc = cam_loc.unsqueeze(1).repeat(1, ray_dirs.shape[1], 1).reshape(-1, 3)
    else:
        ray_dirs = inp["ray_dirs"].unsqueeze(0)
        cam_loc = inp["ray_cam_loc"]
        ray_dirs_tmp = inp["ray_dirs_tmp"].unsqueeze(0)
    depth_scale = ray_dirs_tmp[0, :, 2:]
    bsz, npix, _ = ray_dirs.shape
    ray_dirs = ray_dirs.reshape(-1, 3)
    beta0 = get_beta(params, cfg).detach()
    if z_vals is None:
        z_vals, z_eik = sampler_get_z_vals(cfg, ray_dirs, cam_loc, lambda p: sdf_vals(params, cfg, p), beta0,
                                           training, trace)
    else:
        z_eik = z_vals[:, :1]
    S = z_vals.shape[1]
    pts = (cam_loc.unsqueeze(1) + z_vals.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
    dirs = ray_dirs.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
    sdf, feats, grads = sdf_outputs(params, cfg, pts)
    cout = color_net_forward(params, cfg, pts, grads, dirs, feats, indices, if_pixel_input)
    rgb = cout["rgb"].reshape(-1, S, 3)
    weights = render_weights(z_vals, sdf, get_beta(params, cfg))
    rgb_values = torch.sum(weights.unsqueeze(-1) * rgb, 1)
    depth_values = torch.sum(weights * z_vals, 1, keepdims=True) / (weights.sum(dim=1, keepdims=True) + 1e-8)
    depth_values = depth_scale * depth_values
    if cfg.white_bkgd:
        acc = torch.sum(weights, -1)
        rgb_values = rgb_values + (1.0 - acc[..., None]) * torch.tensor(cfg.bg_color).unsqueeze(0)
    out = {"rgb": rgb, "rgb_values": rgb_values, "depth_values": depth_values, "z_vals": z_vals,
           "depth_vals": z_vals * depth_scale, "sdf": sdf.reshape(z_vals.shape), "weights": weights}
    if cfg.color.spec:
        rs = cout["rgb_spec"].reshape(-1, S, 3)
        out.update(rgb_spec=rs, rgb_spec_values=torch.sum(weights.unsqueeze(-1) * rs, 1))
    if training:
        n = bsz * npix
        R = cfg.scene_bounding_sphere
        eik = torch.empty(n, 3).uniform_(-R, R)
        eik_near = (cam_loc.unsqueeze(1) + z_eik.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
        eik = torch.cat([eik, eik_near], 0)
        nei = eik + (torch.rand_like(eik) - 0.5) * 0.01
        eik = torch.cat([eik, nei], 0)
        if eik_points is not None:   # tests inject the device-generated points of the CUDA run
            eik = eik_points
        if trace is not None:
            trace["eik_points"] = eik.clone()
        gt = sdf_gradient(params, cfg, eik)
        out["grad_theta"] = gt[: gt.shape[0] // 2]
        out["grad_theta_nei"] = gt[gt.shape[0] // 2:]
    normals = grads / (grads.norm(2, -1, keepdim=True) + 1e-6)
    normal_map = torch.sum(weights.unsqueeze(-1) * normals.reshape(-1, S, 3), 1)
    if if_pixel_input:
        rot = inp["ray_pose"][:, :3, :3].transpose(1, 2)
        normal_map = (rot @ normal_map.unsqueeze(-1)).squeeze(-1)
    else:
        rot = inp["pose"][0, :3, :3].permute(1, 0).contiguous()
        normal_map = (rot @ normal_map.permute(1, 0)).permute(1, 0).contiguous()
    out["normal_map"] = normal_map
    return out


# --------------------------------------------------------------------------------------
# utils/plots.py:131-194  (get_surface_sliding: the SDF volume of one crop, coarse to fine with masks)
# --------------------------------------------------------------------------------------
def sdf_volume_pyramid(sdf: Callable[[Tensor], Tensor], lo, hi, crop_n: int, trace: Optional[list] = None) -> Tensor:
    """The volume `z` of plots.py:194 for the crop [lo, hi]^3 with crop_n samples per axis; sdf: [M,3] -> [M]."""
    avg_pool_3d = torch.nn.AvgPool3d(2, stride=2)                      # :105
    upsample = torch.nn.Upsample(scale_factor=2, mode="nearest")       # :106
    x = np.linspace(lo[0], hi[0], crop_n)                              # :135-137
    y = np.linspace(lo[1], hi[1], crop_n)
    z = np.linspace(lo[2], hi[2], crop_n)
    xx, yy, zz = torch.meshgrid(torch.tensor(x), torch.tensor(y), torch.tensor(z), indexing="ij")   # :142
    points = torch.vstack([xx.flatten(), yy.flatten(), zz.flatten()]).T.float()                     # :143
    points = points.reshape(crop_n, crop_n, crop_n, 3).permute(3, 0, 1, 2)                          # :154
    pyramid = [points]
    for _ in range(3):                                                 # :156-158
        points = avg_pool_3d(points[None])[0]
        pyramid.append(points)
    pyramid = pyramid[::-1]
    mask = None
    threshold = 2 * (hi[0] - lo[0]) / crop_n * 8                       # :162
    pts_sdf = None
    for pid, pts in enumerate(pyramid):                                # :164-190
        coarse_n = pts.shape[-1]
        pts = pts.reshape(3, -1).permute(1, 0).contiguous()
        if mask is None:
            pts_sdf = sdf(pts).reshape(-1)
            if trace is not None:
                trace.append((coarse_n, pts.shape[0]))
        else:
            mask = mask.reshape(-1)
            pts_to_eval = pts[mask]
            if pts_to_eval.shape[0] > 0:
                pts_sdf[mask] = sdf(pts_to_eval.contiguous()).reshape(-1)
            if trace is not None:
                trace.append((coarse_n, pts_to_eval.shape[0]))
        if pid < 3:
            mask = torch.abs(pts_sdf) < threshold
            mask = mask.reshape(coarse_n, coarse_n, coarse_n)[None, None]
            mask = upsample(mask.float()).bool()
            pts_sdf = pts_sdf.reshape(coarse_n, coarse_n, coarse_n)[None, None]
            pts_sdf = upsample(pts_sdf).reshape(-1)
        threshold /= 2.0
    return pts_sdf.reshape(crop_n, crop_n, crop_n)


# --------------------------------------------------------------------------------------
# pixel-mode data path: SceneDatasetDN.convert_to_pixels + __getitem__ + collate_fn
# (datasets/scene_dataset.py:258-260, 269-307, 374-401, 438-464)
# --------------------------------------------------------------------------------------
def pixel_bank(poses: Tensor, intrinsics: Tensor, H: int, W: int) -> Dict[str, Tensor]:
    """The per-ray arrays convert_to_pixels materialises for F frames of H x W pixels (all frames selected)."""
    uv = np.mgrid[0:H, 0:W].astype(np.int32)                                   # :258
    uv = torch.from_numpy(np.flip(uv, axis=0).copy()).float()                  # :259  (x = column, y = row)
    uv = uv.reshape(2, -1).transpose(1, 0)                                     # :260  (HW, 2)
    F_ = poses.shape[0]
    uv_all = uv.unsqueeze(0).expand(F_, -1, -1)
    ray_dirs, cam_loc = camera_rays(uv_all, poses, intrinsics)                 # :283
    ray_dirs_tmp, _ = camera_rays(uv_all, torch.eye(4)[None].expand(F_, -1, -1), intrinsics)   # :288
    hw = H * W
    return {"ray_dirs": ray_dirs.reshape(-1, 3), "ray_dirs_tmp": ray_dirs_tmp.reshape(-1, 3),
            "ray_cam_loc": cam_loc.unsqueeze(1).expand(-1, hw, -1).reshape(-1, 3),
            "ray_pose": poses.unsqueeze(1).expand(-1, hw, -1, -1).reshape(-1, 4, 4),
            "ray_frame_idx": torch.arange(F_).reshape(-1, 1).expand(-1, hw).flatten()}     # :305


def pixel_batch(bank: Dict[str, Tensor], images: Dict[str, Tensor], ray_ids: Tensor):
    """__getitem__ for every id + collate_fn (torch.stack / LongTensor): (indices, model_input, ground_truth)."""
    inp = {k: bank[k][ray_ids] for k in ("ray_dirs", "ray_dirs_tmp", "ray_cam_loc", "ray_pose")}
    gt = {k: v.reshape(-1, v.shape[-1])[ray_ids] for k, v in images.items()}
    return bank["ray_frame_idx"][ray_ids].long(), inp, gt


# --------------------------------------------------------------------------------------
# MonoSDFLoss (loss.py:29-49, 75-86, 180-311) in pixel mode -- the consumer right after the path
# --------------------------------------------------------------------------------------
def monosdf_loss(out: Dict[str, Tensor], gt: Dict[str, Tensor], w=None) -> Dict[str, Tensor]:
    w = dict(eik=0.05, smooth=0.005, depth=0.1, nl1=0.05, ncos=0.05) if w is None else w
    rgb_loss = F.l1_loss(out["rgb_values"], gt["rgb"].reshape(-1, 3))
    if "grad_theta" in out:
        eik = ((out["grad_theta"].norm(2, dim=1) - 1) ** 2).mean()
        g1, g2 = out["grad_theta"], out["grad_theta_nei"]
        n1 = g1 / (g1.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        n2 = g2 / (g2.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        smooth = torch.norm(n1 - n2, dim=-1).mean()
    else:
        eik = torch.tensor(0.0)
        smooth = torch.tensor(0.0)
    mask = ((out["sdf"] > 0.0).any(-1) & (out["sdf"] < 0.0).any(-1))[None, :, None]
    mask = (gt["mask"] > 0.5) & mask
    pred = out["depth_values"].reshape(1, -1)
    tgt = (gt["depth"] * 50 + 0.5).reshape(1, -1)
    m = mask.reshape(1, -1).to(pred.dtype)
    a00, a01, a11 = (m * pred * pred).sum(1), (m * pred).sum(1), m.sum(1)
    b0, b1 = (m * pred * tgt).sum(1), (m * tgt).sum(1)
    det = a00 * a11 - a01 * a01
    ok = det != 0
    safe = torch.where(ok, det, torch.ones_like(det))
    x0 = torch.where(ok, (a11 * b0 - a01 * b1) / safe, torch.zeros_like(det))
    x1 = torch.where(ok, (-a01 * b0 + a00 * b1) / safe, torch.zeros_like(det))
    ssi = x0.view(1, -1) * pred + x1.view(1, -1)
    res = ssi - tgt
    M = m.sum(1)
    depth = (m * res * res).sum(1).sum() / (2 * M).sum() if float(M.sum()) != 0 else torch.tensor(0.0)
    npred = F.normalize(out["normal_map"][None] * mask, p=2, dim=-1)
    ngt = F.normalize(gt["normal"], p=2, dim=-1)
    nl1 = torch.abs(npred - ngt).sum(-1).mean()
    ncos = (1.0 - torch.sum(npred * ngt, -1)).mean()
    loss = rgb_loss + w["eik"] * eik + w["smooth"] * smooth + w["depth"] * depth + w["nl1"] * nl1 + w["ncos"] * ncos
    return dict(loss=loss, rgb_loss=rgb_loss, eikonal_loss=eik, smooth_loss=smooth, depth_loss=depth,
                normal_l1=nl1, normal_cos=ncos)


# --------------------------------------------------------------------------------------
# synthetic workload -- shared by tests and bench so both sides see identical inputs
# --------------------------------------------------------------------------------------
def synthetic_rays(n: int, seed: int = 1) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(n, 3, generator=g) - 0.5) * 0.6
    d = F.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return {"ray_dirs": d, "ray_cam_loc": o, "ray_dirs_tmp": d.clone(),
            "ray_pose": torch.eye(4)[None].repeat(n, 1, 1)}


def synthetic_gt(n: int, seed: int = 2) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(1, n, 3, generator=g)
    depth = torch.rand(1, n, 1, generator=g) * 0.06 + 0.02
    normal = F.normalize(torch.randn(1, n, 3, generator=g), dim=-1)
    return {"rgb": rgb, "depth": depth, "normal": normal, "mask": torch.ones(1, n, 1)}
