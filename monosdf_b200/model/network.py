"""MonoSDF's differentiable SDF volume renderer as a drop-in nn.Module on hand-written sm_100a kernels.

Mirrors the interface of reference code/model/network.py: `ImplicitNetwork` (:12-137), `ImplicitNetworkGrid`
(:141-322), `RenderingNetwork` (:325-470) and `MonoSDFNetwork` (:472-640) -- same constructors, parameter names
(state_dict keys), attributes and output dictionary -- so that `train.model_class =
monosdf_b200.model.network.MonoSDFNetwork` in a conf file makes the unmodified trainer / eval scripts use it.

All arithmetic of the hot path runs in libmonosdf_b200.so (see include/monosdf_b200.h); the modules here hold
parameters and wire the kernels into autograd.  grad_x(sdf) is analytic (a reverse sweep through the MLP) and its
double backward is an explicit tangent/backward sweep pair, replacing torch.autograd.grad(create_graph=True).
There is no PyTorch/CPU fallback: without the CUDA library every forward raises.
"""
import contextlib
import os

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib
from ..hashencoder.hashgrid import HashEncoder
from .density import LaplaceDensity
from .embedder import get_embedder
from .ray_sampler import ErrorBoundSampler

_GB = 1 << 30
WORKSPACE_CAP_BYTES = int(os.environ.get("MSDF_WORKSPACE_CAP_GB", "24")) * _GB     # per-call scratch (7.7 GB for a full chunk of the MLP conf)
CHUNK_POINTS_BF16 = 1060864        # points per chunk of the tensor-core mode: csrc/mlp.cu kChunkPointsBf16 (148 SMs x 256 rows x 28)
# Training keeps every layer's activations of a field evaluation in HBM between forward and backward (~10 KB per point
# in bf16 mode, 64 GB for a 65536-ray step) instead of recomputing them, when they fit in this fraction of the FREE
# device memory; set to 0 to always recompute.
SAVED_ACTIVATION_FRACTION = 0.7


def _round4(n):
    return (n + 3) // 4 * 4


# ------------------------------------------------------------------------------------------------------------
# autograd glue
# ------------------------------------------------------------------------------------------------------------
class _WeightNorm(Function):
    """W[o,:] = g[o] v[o,:]/||v[o,:]|| into a row-padded [out, ldw] buffer (nn.utils.weight_norm, dim=0)."""

    @staticmethod
    def forward(ctx, g, v):
        out_dim, in_dim = v.shape
        ldw = _round4(in_dim)
        g, v = g.contiguous(), v.contiguous()
        W = torch.empty(out_dim, ldw, device=v.device, dtype=torch.float32)
        _lib.call("msdf_weightnorm_forward", _lib.ptr(g), _lib.ptr(v), out_dim, in_dim, _lib.ptr(W), ldw, _lib.stream())
        ctx.save_for_backward(g, v)
        return W

    @staticmethod
    def backward(ctx, dW):
        g, v = ctx.saved_tensors
        out_dim, in_dim = v.shape
        dW = dW.contiguous()
        dg, dv = torch.empty_like(g), torch.empty_like(v)
        _lib.call("msdf_weightnorm_backward", _lib.ptr(g), _lib.ptr(v), _lib.ptr(dW), dW.shape[1], out_dim, in_dim,
                  _lib.ptr(dg), _lib.ptr(dv), _lib.stream())
        return dg, dv


def _effective_weights(module, n_lin):
    """[(W_l [out, ldw], b_l)] for lin0..lin{n-1}; weight-normed layers go through the kernel, plain ones are padded.
    Without grad mode (eval renders, SDF grid queries: one model call per 1024-ray chunk / 100 000-point chunk) the result
    is kept while the parameters' version counters stand still: 12 launches + autograd nodes less per call."""
    use_cache = not torch.is_grad_enabled()
    if use_cache:
        key = (_lib.param_epoch[0],) + tuple((p.data_ptr(), p._version) for l in range(n_lin)
                                             for p in getattr(module, "lin" + str(l)).parameters())
        hit = getattr(module, "_weff_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
    packs = []
    for l in range(n_lin):
        lin = getattr(module, "lin" + str(l))
        if hasattr(lin, "weight_g"):
            W = _WeightNorm.apply(lin.weight_g, lin.weight_v)
        else:
            W = lin.weight
            pad = _round4(W.shape[1]) - W.shape[1]
            if pad:
                W = torch.nn.functional.pad(W, (0, pad))
        packs.append((W, lin.bias))
    if use_cache:
        module._weff_cache = (key, [(W.detach(), b.detach()) for W, b in packs])
        return module._weff_cache[1]
    return packs


def _pose_from_quaternion(pose7):
    """[B,7] (unit-normalised quaternion w,x,y,z + camera centre) -> [B,4,4] camera-to-world matrices."""
    q = torch.nn.functional.normalize(pose7[:, :4].float(), dim=1)
    w, x, y, z = q.unbind(dim=1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]
    p = torch.eye(4, device=pose7.device, dtype=torch.float32).repeat(pose7.shape[0], 1, 1)
    p[:, :3, :3] = torch.stack(rows, dim=1).reshape(-1, 3, 3)
    p[:, :3, 3] = pose7[:, 4:].float()
    return p


class _NetSpec:
    """Static geometry of one MLP as the C ABI wants it."""

    def __init__(self, in_dims, out_dims, d0, skip_layer):
        self.in_dims, self.out_dims, self.d0, self.skip_layer = list(in_dims), list(out_dims), d0, skip_layer
        self.n = len(in_dims)

    def desc(self, tensors):
        d = _lib.MlpDesc()
        d.n_layers, d.d0, d.skip_layer = self.n, self.d0, self.skip_layer
        for l in range(self.n):
            W, b = tensors[2 * l], tensors[2 * l + 1]
            d.in_dim[l], d.out_dim[l], d.ldw[l] = self.in_dims[l], self.out_dims[l], W.shape[1]
            d.W[l], d.b[l] = _lib.ptr(W), _lib.ptr(b)
        return d

    def grads(self, tensors):
        g = _lib.MlpGrads()
        outs = []
        for l in range(self.n):
            dW, db = torch.zeros_like(tensors[2 * l]), torch.zeros_like(tensors[2 * l + 1])
            g.dW[l], g.db[l] = _lib.ptr(dW), _lib.ptr(db)
            outs += [dW, db]
        return g, outs


class _FieldSpec:
    """Everything non-tensor that one field evaluation needs."""

    def __init__(self, sdf_spec, multires, grid, color_spec=None, color=None):
        self.sdf_spec, self.multires, self.grid, self.color_spec, self.color = sdf_spec, multires, grid, color_spec, color
        self.flags = 0

    def enc_desc(self, table, offsets):
        e = _lib.EncodingDesc()
        e.multires = self.multires
        if self.grid is not None:
            e.grid_feat_dim = self.grid["L"] * self.grid["C"]
            e.n_levels, e.level_dim, e.base_res = self.grid["L"], self.grid["C"], self.grid["H"]
            e.log2_per_level_scale, e.divide_factor = self.grid["S"], self.grid["divide_factor"]
            e.table = _lib.ptr(table) if table is not None else None
            e.offsets = _lib.ptr(offsets) if offsets is not None else None
        return e

    def color_desc(self, code_per_ray):
        c = _lib.ColorDesc()
        c.mode_idr = 1 if self.color["mode"] == "idr" else 0
        c.multires_view, c.feat_dim = self.color["multires_view"], self.color["feat_dim"]
        c.code_dim, c.code_per_ray = self.color["code_dim"], int(code_per_ray)
        c.final_act = 1 if self.color["hdr"] else 0
        c.spec = 1 if self.color.get("spec") else 0
        return c


def _workspace_for(sdf_d, enc_d, col_d, cd_d, M, mode, flags, device):
    base = int(os.environ.get("MSDF_CHUNK_POINTS", str(CHUNK_POINTS_BF16)))
    cap = 2 * base if mode == _lib.MODE_SDF_ONLY else (base if flags & _lib.FLAG_TENSOR_BF16 else 65536)
    chunk = min(max(int(M), 128), cap)
    need = _lib.lib().msdf_field_workspace_bytes(sdf_d, enc_d, col_d, cd_d, chunk, mode, flags)
    if need == 0:
        raise RuntimeError("monosdf_b200: msdf_field_workspace_bytes failed: " + _lib.lib().msdf_last_error().decode())
    return _lib.workspace(min(need, WORKSPACE_CAP_BYTES), device)


_CALLER_GRAD = [True]      # torch.is_grad_enabled() of the code that called _apply_field() (see _Field.forward)


def _apply_field(*args):
    """_Field.apply with the caller's grad mode on record."""
    _CALLER_GRAD[0] = torch.is_grad_enabled()
    try:
        return _Field.apply(*args)
    finally:
        _CALLER_GRAD[0] = True


class _Field(Function):
    """sdf / grad_x sdf / features / colours at M points, with the analytic (double) backward.

    kind: 'sdf' (get_sdf_vals), 'forward' (raw output), 'outputs' (get_outputs), 'gradient' (gradient_sdf),
          'render' (get_outputs + RenderingNetwork).
    args: x [M,3], view_dirs [n_rays,3]|None, code [n_rays|1, 32]|None, table|None, offsets|None, then
          W0, b0, ... of the SDF net followed by those of the colour net.
    Returns (sdf [M,1]|None, grad [M,3]|None, feat [M,F]|None, rgb [M,3]|None).
    """

    @staticmethod
    def forward(ctx, spec, kind, clamp_radius, sphere_scale, n_samples, x, view_dirs, code, table, offsets, *params):
        x = x.detach().contiguous().float()
        M, dev = x.shape[0], x.device
        ns = 2 * spec.sdf_spec.n
        sdf_t, col_t = params[:ns], params[ns:]
        use_color = kind == "render"
        sdf_d = spec.sdf_spec.desc(sdf_t)
        enc_d = spec.enc_desc(table, offsets)
        col_d = spec.color_spec.desc(col_t) if use_color else None
        code_per_ray = code is not None and code.shape[0] > 1
        cd_d = spec.color_desc(code_per_ray) if use_color else None
        mode = _lib.MODE_SDF_ONLY if kind == "sdf" else _lib.MODE_FORWARD
        # the engine is picked ONCE, here: kinds that return the raw feature vector ('forward', 'outputs') run the fp32
        # engine in every precision mode (msdf_field_forward only has fp32 feature outputs), and so must their saved
        # activations, workspace and backward
        flags = spec.flags & ~_lib.FLAG_TENSOR_BF16 if kind in ("forward", "outputs") else spec.flags
        F_dim = spec.sdf_spec.out_dims[-1] - 1
        sdf = torch.empty(M, 1, device=dev) if kind != "gradient" else None
        grad = torch.empty(M, 3, device=dev) if kind in ("outputs", "gradient", "render") else None
        feat = torch.empty(M, F_dim, device=dev) if kind in ("forward", "outputs") else None
        # spec variant: [rgb | rgb_spec] per point (include/monosdf_b200.h, msdf_color_desc.spec)
        n_rgb = 6 if (use_color and spec.color.get("spec")) else (spec.color_spec.out_dims[-1] if use_color else 0)
        rgb = torch.empty(M, n_rgb, device=dev) if use_color else None
        if use_color:
            view_dirs = view_dirs.contiguous().float()
            code = code.contiguous().float() if code is not None else None
        n_rays = view_dirs.shape[0] if use_color else 0
        saved = None
        # (ctx.needs_input_grad reflects requires_grad even under torch.no_grad(), and grad mode is always off INSIDE a
        # Function's forward: the caller's grad mode is recorded by _apply_field() -- eval must not save activations)
        if _CALLER_GRAD[0] and any(ctx.needs_input_grad) and grad is not None and M > 0 and SAVED_ACTIVATION_FRACTION > 0:
            nbytes = _lib.lib().msdf_field_saved_bytes(sdf_d, enc_d, col_d, cd_d, M, int(n_samples), flags)
            # does it fit?  A pooled buffer of the right size is free by construction; otherwise ask the driver, and
            # only if that says no also count what torch's caching allocator holds but has not handed out (slow query)
            if _lib.saved_pool.best_fit_bytes(nbytes, dev) > 0:
                saved = _lib.saved_pool.acquire(nbytes, dev)
            elif nbytes > 0:
                free = torch.cuda.mem_get_info(dev)[0] + sum(t.numel() for t in _lib.saved_pool.free if t.device == dev)
                if nbytes > SAVED_ACTIVATION_FRACTION * free:
                    free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
                if nbytes <= SAVED_ACTIVATION_FRACTION * free:
                    saved = _lib.saved_pool.acquire(nbytes, dev)
        bwd_mode = _lib.MODE_BACKWARD if saved is not None else mode      # the saved layout needs the backward's workspace
        ws = _workspace_for(sdf_d, enc_d, col_d, cd_d, M, bwd_mode, flags, dev)
        _lib.call("msdf_field_forward", sdf_d, enc_d, col_d, cd_d, _lib.ptr(x), M, _lib.ptr(view_dirs) if use_color else None,
                  n_rays, int(n_samples), _lib.ptr(code) if use_color else None, mode, float(clamp_radius), float(sphere_scale),
                  flags, _lib.ptr(ws), ws.numel(), _lib.ptr(sdf), _lib.ptr(grad), _lib.ptr(feat), F_dim, _lib.ptr(rgb),
                  _lib.ptr(saved), saved.numel() if saved is not None else 0, _lib.stream())
        ctx.spec, ctx.kind, ctx.clamp, ctx.sphere_scale, ctx.n_samples, ctx.ns = spec, kind, clamp_radius, sphere_scale, n_samples, ns
        ctx.code_per_ray = code_per_ray
        ctx.flags = flags
        ctx.saved_acts = saved
        ctx.save_for_backward(x, view_dirs if use_color else None, code if use_color else None, table, offsets, rgb, *params)
        return sdf, grad, feat, rgb

    @staticmethod
    def backward(ctx, d_sdf, d_grad, d_feat, d_rgb):
        x, view_dirs, code, table, offsets, rgb, *params = ctx.saved_tensors
        spec, kind, ns = ctx.spec, ctx.kind, ctx.ns
        M, dev = x.shape[0], x.device
        sdf_t, col_t = params[:ns], params[ns:]
        use_color = kind == "render" and d_rgb is not None
        sdf_d = spec.sdf_spec.desc(sdf_t)
        enc_d = spec.enc_desc(table, offsets)
        col_d = spec.color_spec.desc(col_t) if use_color else None
        cd_d = spec.color_desc(ctx.code_per_ray) if use_color else None
        sdf_g, sdf_outs = spec.sdf_spec.grads(sdf_t)
        col_g, col_outs = (spec.color_spec.grads(col_t) if use_color else (None, [None] * len(col_t)))
        if kind == "render" and not use_color:
            col_outs = [None] * len(col_t)
        d_table = torch.zeros_like(table) if (table is not None and ctx.needs_input_grad[8]) else None
        d_code = torch.zeros_like(code) if (use_color and code is not None) else None

        def c(t):
            return t.contiguous().float() if t is not None else None
        d_sdf, d_grad, d_feat, d_rgb = c(d_sdf), c(d_grad), c(d_feat), c(d_rgb)
        F_dim = spec.sdf_spec.out_dims[-1] - 1
        n_rays = view_dirs.shape[0] if use_color else 0
        ws = _workspace_for(sdf_d, enc_d, col_d, cd_d, M, _lib.MODE_BACKWARD, ctx.flags, dev)
        # the backward consumes (overwrites) the saved activations: a second backward through the same node recomputes.
        # They were laid out for the forward's network (with the colour net for 'render'), so they are only usable
        # when this backward sees the same one.
        ctx_saved = ctx.saved_acts
        saved = ctx_saved if (use_color or kind != "render") else None
        ctx.saved_acts = None
        _lib.call("msdf_field_backward", sdf_d, enc_d, col_d, cd_d, _lib.ptr(x), M, _lib.ptr(view_dirs) if use_color else None,
                  n_rays, int(ctx.n_samples), _lib.ptr(code) if use_color else None, float(ctx.clamp), float(ctx.sphere_scale),
                  ctx.flags, _lib.ptr(ws), ws.numel(), _lib.ptr(d_sdf), _lib.ptr(d_grad), _lib.ptr(d_feat), F_dim,
                  _lib.ptr(rgb) if use_color else None, _lib.ptr(d_rgb) if use_color else None, sdf_g, col_g,
                  _lib.ptr(d_table), _lib.ptr(d_code), _lib.ptr(saved), saved.numel() if saved is not None else 0, _lib.stream())
        _lib.saved_pool.release(ctx_saved)
        del saved
        return (None, None, None, None, None, None, None, d_code, d_table, None, *sdf_outs, *col_outs)


class _CodeLookup(Function):
    """embeddings[indices] (network.py:400-413) with a scatter backward that does not serialise on equal indices."""

    @staticmethod
    def forward(ctx, table, indices):
        indices = indices.reshape(-1).long().contiguous()
        ctx.save_for_backward(indices)
        ctx.shape = table.shape
        return table.detach()[indices]

    @staticmethod
    def backward(ctx, d_code):
        (indices,) = ctx.saved_tensors
        d_code = d_code.contiguous().float()
        d_table = torch.zeros(ctx.shape, device=d_code.device, dtype=torch.float32)
        _lib.call("msdf_code_scatter", _lib.ptr(d_code), _lib.ptr(indices), indices.shape[0], ctx.shape[1], ctx.shape[0],
                  _lib.ptr(d_table), _lib.stream())
        return d_table, None


class _Composite(Function):
    """Laplace density + alpha compositing (density.py:21-30, network.py:626-640,552-562,603-616)."""

    @staticmethod
    def forward(ctx, z_vals, sdf, rgb, grad, beta, depth_scale, ds_stride, pose, pose_per_ray, white_bkgd, bg_color):
        N, S = z_vals.shape
        dev = z_vals.device
        z_vals, sdf, rgb, grad = z_vals.contiguous(), sdf.contiguous(), rgb.contiguous(), grad.contiguous()
        beta = beta.detach().reshape(1).float().contiguous()
        pose = pose.contiguous().float()
        weights = torch.empty(N, S, device=dev)
        rgb_values = torch.empty(N, 3, device=dev)
        depth_values = torch.empty(N, 1, device=dev)
        normal_map = torch.empty(N, 3, device=dev)
        _lib.call("msdf_render_forward", _lib.ptr(z_vals), _lib.ptr(sdf), _lib.ptr(rgb), _lib.ptr(grad), N, S, _lib.ptr(beta),
                  depth_scale.data_ptr(), int(ds_stride), _lib.ptr(pose), int(pose_per_ray), int(white_bkgd),
                  _lib.ptr(bg_color) if white_bkgd else None, _lib.ptr(weights), _lib.ptr(rgb_values), _lib.ptr(depth_values),
                  _lib.ptr(normal_map), _lib.stream())
        ctx.save_for_backward(z_vals, sdf, rgb, grad, beta, depth_scale, pose, bg_color)
        ctx.cfg = (int(ds_stride), int(pose_per_ray), int(white_bkgd))
        return weights, rgb_values, depth_values, normal_map

    @staticmethod
    def backward(ctx, d_weights, d_rgbv, d_depth, d_nmap):
        z_vals, sdf, rgb, grad, beta, depth_scale, pose, bg_color = ctx.saved_tensors
        ds_stride, pose_per_ray, white_bkgd = ctx.cfg
        N, S = z_vals.shape

        def c(t):
            return t.contiguous().float() if t is not None else None
        d_weights, d_rgbv, d_depth, d_nmap = c(d_weights), c(d_rgbv), c(d_depth), c(d_nmap)
        d_sdf, d_rgb, d_grad = torch.empty_like(sdf), torch.empty_like(rgb), torch.empty_like(grad)
        d_beta = torch.zeros(1, device=z_vals.device)
        _lib.call("msdf_render_backward", _lib.ptr(z_vals), _lib.ptr(sdf), _lib.ptr(rgb), _lib.ptr(grad), N, S, _lib.ptr(beta),
                  depth_scale.data_ptr(), ds_stride, _lib.ptr(pose), pose_per_ray, white_bkgd,
                  _lib.ptr(bg_color) if white_bkgd else None, _lib.ptr(d_weights), _lib.ptr(d_rgbv), _lib.ptr(d_depth),
                  _lib.ptr(d_nmap), _lib.ptr(d_sdf), _lib.ptr(d_rgb), _lib.ptr(d_grad), _lib.ptr(d_beta), _lib.stream())
        return None, d_sdf, d_rgb, d_grad, d_beta.reshape(()), None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------------------
# SDF / feature networks
# ------------------------------------------------------------------------------------------------------------
def _geometric_init(lin, l, n_lin, dims, out_dim, multires, skip_in, bias, inside_outside):
    """Geometric (sphere) initialisation of SAL/IDR as used by the reference (network.py:51-70)."""
    if l == n_lin - 1:
        mean = np.sqrt(np.pi) / np.sqrt(dims[l])
        torch.nn.init.normal_(lin.weight, mean=-mean if inside_outside else mean, std=0.0001)
        torch.nn.init.constant_(lin.bias, bias if inside_outside else -bias)
        return
    torch.nn.init.constant_(lin.bias, 0.0)
    std = np.sqrt(2) / np.sqrt(out_dim)
    if multires > 0 and l == 0:
        torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
        torch.nn.init.normal_(lin.weight[:, :3], 0.0, std)
    elif multires > 0 and l in skip_in:
        torch.nn.init.normal_(lin.weight, 0.0, std)
        torch.nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
    else:
        torch.nn.init.normal_(lin.weight, 0.0, std)


class _ImplicitBase(nn.Module):
    """Shared machinery of ImplicitNetwork / ImplicitNetworkGrid."""

    def _build_layers(self, dims, geometric_init, bias, skip_in, weight_norm, multires, inside_outside):
        self.num_layers = len(dims)
        self.skip_in = skip_in
        n_lin = self.num_layers - 1
        in_dims, out_dims = [], []
        for l in range(n_lin):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:
                _geometric_init(lin, l, n_lin, dims, out_dim, multires, skip_in, bias, inside_outside)
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
            in_dims.append(dims[l])
            out_dims.append(out_dim)
        skips = [l for l in skip_in if 0 < l < n_lin]
        if len(skips) > 1:
            raise NotImplementedError("monosdf_b200: at most one skip connection is supported")
        self._net_spec = _NetSpec(in_dims, out_dims, dims[0], skips[0] if skips else -1)
        self.softplus = nn.Softplus(beta=100)
        self._cached = None

    # effective weights are computed once per model forward and shared by the sampler / render / eikonal passes
    @contextlib.contextmanager
    def cached_weights(self):
        self._cached = self._flat_weights()
        try:
            yield
        finally:
            self._cached = None

    def _flat_weights(self):
        if self._cached is not None:
            return self._cached
        flat = []
        for W, b in _effective_weights(self, self.num_layers - 1):
            flat += [W, b]
        return flat

    def _table(self):
        return None, None

    def _field(self, kind, x, clamp):
        table, offsets = self._table()
        return _apply_field(self._field_spec, kind, clamp, self.sphere_scale, 1, x, None, None, table, offsets,
                            *self._flat_weights())

    def gradient_sdf(self, x):
        """grad_x sdf without the sphere clamp (network.py:98-109, 277-288)."""
        return self._field("gradient", x, 0.0)[1]

    def mlp_parameters(self):
        params = []
        for l in range(self.num_layers - 1):
            params += list(getattr(self, "lin" + str(l)).parameters())
        return params


class ImplicitNetwork(_ImplicitBase):
    def __init__(self, feature_vector_size, sdf_bounding_sphere, d_in, d_out, dims, geometric_init=True, bias=1.0,
                 skip_in=(), weight_norm=True, multires=0, sphere_scale=1.0, inside_outside=False):
        super().__init__()
        self.sdf_bounding_sphere = sdf_bounding_sphere
        self.sphere_scale = sphere_scale
        if d_in != 3 or d_out != 1:
            raise NotImplementedError("monosdf_b200: the SDF network takes 3-d points and returns one SDF value")
        dims = [d_in] + list(dims) + [d_out + feature_vector_size]
        self.embed_fn = None
        self.multires = multires
        if multires > 0:
            self.embed_fn, dims[0] = get_embedder(multires, input_dims=d_in)
        self._build_layers(dims, geometric_init, bias, skip_in, weight_norm, multires, inside_outside)
        self._field_spec = _FieldSpec(self._net_spec, multires, None)

    def forward(self, input):
        """Raw network output [M, 1 + F] (network.py:79-96)."""
        sdf, _, feat, _ = self._field("forward", input, 0.0)
        return torch.cat([sdf, feat], dim=1)

    def get_outputs(self, x):
        sdf, grad, feat, _ = self._field("outputs", x, self.sdf_bounding_sphere)
        return sdf, feat, grad

    def get_sdf_vals(self, x):
        return self._field("sdf", x, self.sdf_bounding_sphere)[0]


class ImplicitNetworkGrid(_ImplicitBase):
    def __init__(self, feature_vector_size, sdf_bounding_sphere, d_in, d_out, dims, geometric_init=True, bias=1.0,
                 skip_in=(), weight_norm=True, multires=0, sphere_scale=1.0, inside_outside=False, base_size=16,
                 end_size=2048, logmap=19, num_levels=16, level_dim=2, divide_factor=1.5, use_grid_feature=True,
                 debug=False):
        super().__init__()
        self.sdf_bounding_sphere = sdf_bounding_sphere
        self.sphere_scale = sphere_scale
        if d_in != 3 or d_out != 1:
            raise NotImplementedError("monosdf_b200: the SDF network takes 3-d points and returns one SDF value")
        dims = [d_in] + list(dims) + [d_out + feature_vector_size]
        self.embed_fn = None
        self.divide_factor = divide_factor
        self.grid_feature_dim = num_levels * level_dim
        self.use_grid_feature = use_grid_feature
        self.debug = debug
        self.multires = multires
        dims[0] += self.grid_feature_dim
        self.encoding = HashEncoder(input_dim=3, num_levels=num_levels, level_dim=level_dim, per_level_scale=2,
                                    base_resolution=base_size, log2_hashmap_size=logmap, desired_resolution=end_size)
        if multires > 0:
            self.embed_fn, input_ch = get_embedder(multires, input_dims=d_in)
            dims[0] += input_ch - 3
        self._build_layers(dims, geometric_init, bias, skip_in, weight_norm, multires, inside_outside)
        grid = dict(L=num_levels, C=level_dim, H=base_size, S=float(np.log2(self.encoding.per_level_scale)),
                    divide_factor=float(divide_factor))
        self._field_spec = _FieldSpec(self._net_spec, multires, grid)
        self.cache_sdf = None

    def _table(self):
        if not self.use_grid_feature:   # zero features (network.py:251-252): the table does not enter the graph
            return None, None
        return self.encoding.embeddings, self.encoding.offsets

    def forward(self, input):
        sdf, _, feat, _ = self._field("forward", input, 0.0)
        return {"sdf": sdf, "feature": feat}

    def get_outputs(self, x):
        sdf, grad, feat, _ = self._field("outputs", x, 0.0)   # no sphere clamp in the grid variant (:290-305)
        return sdf, feat, grad

    def get_sdf_vals(self, x):
        return self._field("sdf", x, 0.0)[0]

    def grid_parameters(self):
        return self.encoding.parameters()


# ------------------------------------------------------------------------------------------------------------
# colour network
# ------------------------------------------------------------------------------------------------------------
class RenderingNetwork(nn.Module):
    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True, multires_view=0,
                 per_image_code=False, if_hdr=False, spec=False, debug=False):
        super().__init__()
        if mode not in ("idr", "nerf"):
            raise NotImplementedError(mode)
        self.mode, self.debug, self.spec = mode, debug, spec
        dims = [d_in + feature_vector_size] + list(dims) + [d_out]
        self.embedview_fn = None
        self.multires_view = multires_view
        if multires_view > 0:
            self.embedview_fn, input_ch = get_embedder(multires_view)
            dims[0] += input_ch - 3
        self.per_image_code = per_image_code
        if per_image_code:
            self.embeddings = nn.Parameter(torch.empty(1024, 32))
            self.embeddings.data.uniform_(-1e-4, 1e-4)
            dims[0] += 32
        self.num_layers = len(dims)
        self.if_hdr = if_hdr
        in_dims = list(dims[:-1])
        if spec:
            # diffuse/specular split (network.py:376-380, 427-454): the first 3 outputs of layer num_layers-4 are the
            # diffuse colour, the rest feeds layer num_layers-3; every layer is followed by ReLU (HDR only)
            assert if_hdr, "spec=True is an HDR-only variant (network.py:428)"
            assert self.num_layers >= 5 and d_out == 3, "spec=True needs at least 4 linear layers and d_out = 3"
            in_dims[self.num_layers - 3] -= 3
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            if in_dims[l] != dims[l]:      # the reference builds the full-width layer first and then replaces it
                lin = nn.Linear(in_dims[l], dims[l + 1])   # (network.py:375-380): same draws from torch's RNG
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.relu, self.sigmoid = nn.ReLU(), nn.Sigmoid()
        self._net_spec = _NetSpec(in_dims, dims[1:], dims[0], -1)
        self._color_cfg = dict(mode=mode, multires_view=multires_view, feat_dim=feature_vector_size,
                               code_dim=32 if per_image_code else 0, hdr=bool(if_hdr), spec=bool(spec))
        self._cached = None

    def _flat_weights(self):
        if self._cached is not None:
            return self._cached
        flat = []
        for W, b in _effective_weights(self, self.num_layers - 1):
            flat += [W, b]
        return flat

    def image_code(self, indices, if_pixel_input):
        """[1,32] (one image, network.py:409) or [n_rays,32] (pixel mode, :411-412); None without per_image_code."""
        if not self.per_image_code:
            return None
        if self.embeddings.is_cuda:
            return _CodeLookup.apply(self.embeddings, indices).reshape(-1, 32)
        return self.embeddings[indices].reshape(-1, 32)     # CPU: parameter container only (no rendering path there)

    def forward(self, points, normals, view_dirs, feature_vectors, indices, if_pixel_input=False):
        raise NotImplementedError(
            "monosdf_b200.RenderingNetwork is evaluated fused with the SDF network inside MonoSDFNetwork.forward "
            "(msdf_field_forward); a stand-alone call on materialised feature vectors is not part of the hot path")


# ------------------------------------------------------------------------------------------------------------
# the renderer
# ------------------------------------------------------------------------------------------------------------
class MonoSDFNetwork(nn.Module):
    def __init__(self, conf, if_hdr=False):
        super().__init__()
        self.feature_vector_size = conf.get_int("feature_vector_size")
        self.scene_bounding_sphere = conf.get_float("scene_bounding_sphere", default=1.0)
        self.white_bkgd = conf.get_bool("white_bkgd", default=False)
        self.register_buffer("bg_color", torch.tensor(conf.get_list("bg_color", default=[1.0, 1.0, 1.0])).float(),
                             persistent=False)
        self.if_hdr = if_hdr
        self.Grid_MLP = conf.get_bool("Grid_MLP", default=False)
        radius = 0.0 if self.white_bkgd else self.scene_bounding_sphere
        net_cls = ImplicitNetworkGrid if self.Grid_MLP else ImplicitNetwork
        self.implicit_network = net_cls(self.feature_vector_size, radius, **conf.get_config("implicit_network"))
        self.rendering_network = RenderingNetwork(self.feature_vector_size, **conf.get_config("rendering_network"),
                                                  if_hdr=self.if_hdr)
        self.spec = conf.get_config("rendering_network").get_bool("spec", False)
        self.density = LaplaceDensity(**conf.get_config("density"))
        self.ray_sampler = ErrorBoundSampler(self.scene_bounding_sphere, **conf.get_config("ray_sampler"))
        inet = self.implicit_network
        self._render_spec = _FieldSpec(inet._net_spec, inet.multires, inet._field_spec.grid,
                                       self.rendering_network._net_spec, self.rendering_network._color_cfg)
        # 'device': random draws on the GPU; 'reference': CPU generator in the reference's order (parity tests)
        self.rng = "device"

    def set_precision(self, mode):
        """'fp32' (1e-4 parity with the reference) or 'bf16' (tcgen05 tensor cores, 2e-2)."""
        flags = {"fp32": 0, "bf16": _lib.FLAG_TENSOR_BF16}[mode]
        self._render_spec.flags = flags
        self.implicit_network._field_spec.flags = flags

    def _rays(self, input, if_pixel_input):
        if if_pixel_input:
            ray_dirs = input["ray_dirs"].reshape(-1, 3).contiguous().float()
            cam_loc = input["ray_cam_loc"].reshape(-1, 3).contiguous().float()
            ray_dirs_tmp = input["ray_dirs_tmp"].reshape(-1, 3).contiguous().float()
            return ray_dirs, cam_loc, ray_dirs_tmp, 1, ray_dirs.shape[0]
        uv, pose, intrinsics = input["uv"], input["pose"], input["intrinsics"]
        if pose.dim() == 2 and pose.shape[1] == 7:      # [qr, qi, qj, qk, tx, ty, tz] (rend_util.py:63-70, 121-138)
            pose = _pose_from_quaternion(pose)
        self._pose_matrix = pose
        B, N = uv.shape[0], uv.shape[1]
        if B != 1:
            # the reference breaks here too (depth_scale = ray_dirs_tmp[0] has N rows against B*N rays, network.py:522,555);
            # fail loudly instead of reading depth_scale out of bounds
            raise ValueError("monosdf_b200: image (uv) input supports batch size 1, got %d" % B)
        dev = uv.device
        uv, pose, intrinsics = uv.contiguous().float(), pose.contiguous().float(), intrinsics.contiguous().float()
        eye = torch.eye(4, device=dev).repeat(B, 1, 1)
        outs = []
        for p in (pose, eye):   # second pass: un-rotated directions for the depth scale (network.py:513)
            dirs = torch.empty(B, N, 3, device=dev)
            loc = torch.empty(B, 3, device=dev)
            _lib.call("msdf_camera_rays", _lib.ptr(uv), _lib.ptr(p), _lib.ptr(intrinsics), B, N, _lib.ptr(dirs), _lib.ptr(loc),
                      _lib.stream())
            outs.append((dirs, loc))
        (ray_dirs, cam_loc), (ray_dirs_tmp, _) = outs
        cam_loc = cam_loc.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)
        return ray_dirs.reshape(-1, 3), cam_loc, ray_dirs_tmp[0].contiguous(), B, N

    def forward(self, input, indices, if_pixel_input=False):
        ray_dirs, cam_loc, ray_dirs_tmp, batch_size, num_pixels = self._rays(input, if_pixel_input)
        dev = ray_dirs.device
        depth_scale = ray_dirs_tmp[:, 2:]                      # [N,1] view, stride 3 (network.py:522)
        inet = self.implicit_network
        self.ray_sampler.rng = self.rng
        with inet.cached_weights():
            z_vals, z_samples_eik = self.ray_sampler.get_z_vals(ray_dirs, cam_loc, self)
            N, S = z_vals.shape
            points = torch.empty(N * S, 3, device=dev)
            _lib.call("msdf_ray_points", _lib.ptr(cam_loc), _lib.ptr(ray_dirs), _lib.ptr(z_vals), N, S, _lib.ptr(points),
                      _lib.stream())
            code = self.rendering_network.image_code(indices, if_pixel_input)
            table, offsets = inet._table()
            clamp = inet.sdf_bounding_sphere if not self.Grid_MLP else 0.0
            sdf, grad, _, rgb_flat = _apply_field(self._render_spec, "render", clamp, inet.sphere_scale, S, points, ray_dirs,
                                                  code, table, offsets, *inet._flat_weights(),
                                                  *self.rendering_network._flat_weights())
            rgb_spec = None
            if self.spec:      # [rgb | rgb_spec] per point (network.py:441-453)
                rgb_flat, rgb_spec = rgb_flat[:, :3].contiguous(), rgb_flat[:, 3:].reshape(-1, S, 3)
            if if_pixel_input:
                pose, pose_per_ray = input["ray_pose"], 1
            else:
                pose, pose_per_ray = self._pose_matrix[:1], 0
            weights, rgb_values, depth_values, normal_map = _Composite.apply(
                z_vals, sdf.reshape(N, S), rgb_flat, grad, self.density.get_beta(), depth_scale, 3, pose, pose_per_ray,
                self.white_bkgd, self.bg_color)
            output = {
                "rgb": rgb_flat.reshape(-1, S, 3),
                "rgb_values": rgb_values,
                "depth_values": depth_values,
                "z_vals": z_vals,
                "depth_vals": z_vals * depth_scale,
                "sdf": sdf.reshape(z_vals.shape),
                "weights": weights,
            }
            if self.spec:      # network.py:576-582 (auxiliary output of the archived variant: plain torch on our weights)
                output["rgb_spec"] = rgb_spec
                output["rgb_spec_values"] = torch.sum(weights.unsqueeze(-1) * rgb_spec, 1)
            if self.training:
                n_eik = batch_size * num_pixels
                R = self.scene_bounding_sphere
                if self.rng == "reference":
                    eik = torch.empty(n_eik, 3).uniform_(-R, R).to(dev)
                else:
                    eik = torch.empty(n_eik, 3, device=dev).uniform_(-R, R)
                eik_near = torch.empty(N, 3, device=dev)
                _lib.call("msdf_ray_points", _lib.ptr(cam_loc), _lib.ptr(ray_dirs), _lib.ptr(z_samples_eik.contiguous()), N, 1,
                          _lib.ptr(eik_near), _lib.stream())
                eik = torch.cat([eik, eik_near], 0)
                eik = torch.cat([eik, eik + (torch.rand_like(eik) - 0.5) * 0.01], 0)
                self._last_eikonal_points = eik
                grad_theta = inet.gradient_sdf(eik)
                half = grad_theta.shape[0] // 2
                output["grad_theta"] = grad_theta[:half]
                output["grad_theta_nei"] = grad_theta[half:]
            output["normal_map"] = normal_map
        return output

    def volume_rendering(self, z_vals, sdf):
        """Compositing weights only (network.py:626-640); kept for API compatibility."""
        N, S = z_vals.shape
        dev = z_vals.device
        zeros3 = torch.zeros(N * S, 3, device=dev)
        ones = torch.ones(N, 3, device=dev)
        eye = torch.eye(4, device=dev)[None]
        weights, _, _, _ = _Composite.apply(z_vals, sdf.reshape(N, S), zeros3, zeros3, self.density.get_beta(), ones[:, 2:], 3,
                                            eye, 0, False, self.bg_color)
        return weights
