"""Multi-resolution hash-grid encoder: the op boundary of the reference's one native extension.

Mirrors reference code/hashencoder/hashgrid.py (`HashEncoder` :107-166, `hash_encode` :14-104) on top of
libmonosdf_b200's msdf_hash_encode_{forward,backward,second_backward}, which keep the tensor layouts of the
reference's pybind functions (hashencoder.h:13-15).  Inside MonoSDFNetwork the encoder is not called through
this module: the field kernels read the table directly (csrc/mlp.cu); this module is the stand-alone operator.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib


def _geometry(embeddings, offsets, per_level_scale, inputs):
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = float(np.log2(per_level_scale))   # hashgrid.py:30
    return B, D, C, L, S


class _HashEncode(Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False):
        inputs = inputs.contiguous().float()
        embeddings = embeddings.contiguous()
        offsets = offsets.contiguous()
        B, D, C, L, S = _geometry(embeddings, offsets, per_level_scale, inputs)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=torch.float32)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=torch.float32) if calc_grad_inputs else None
        _lib.call("msdf_hash_encode_forward", _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets), _lib.ptr(outputs),
                  B, D, C, L, S, int(base_resolution), int(calc_grad_inputs), _lib.ptr(dy_dx), _lib.stream())
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.geom = (B, D, C, L, S, int(base_resolution), bool(calc_grad_inputs))
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, calc = ctx.geom
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_inputs, grad_embeddings = _HashEncodeBackward.apply(grad, inputs, embeddings, offsets, B, D, C, L, S, H, calc, dy_dx)
        return (grad_inputs if calc else None), grad_embeddings, None, None, None, None


class _HashEncodeBackward(Function):
    """Backward as a Function of its own so that it is differentiable (the eikonal term, hashgrid.py:71-101)."""

    @staticmethod
    def forward(ctx, grad, inputs, embeddings, offsets, B, D, C, L, S, H, calc, dy_dx):
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs) if calc else None
        _lib.call("msdf_hash_encode_backward", _lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets),
                  _lib.ptr(grad_embeddings), B, D, C, L, S, H, int(calc), _lib.ptr(dy_dx), _lib.ptr(grad_inputs), _lib.stream())
        ctx.save_for_backward(grad, inputs, embeddings, offsets, dy_dx)
        ctx.geom = (B, D, C, L, S, H, calc)
        return grad_inputs, grad_embeddings

    @staticmethod
    def backward(ctx, gg_inputs, gg_embeddings):
        grad, inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, calc = ctx.geom
        if gg_inputs is None or not calc:
            return (None,) * 12
        gg_inputs = gg_inputs.contiguous()
        grad_grad = torch.zeros_like(grad)
        grad2_embeddings = torch.zeros_like(embeddings)
        _lib.call("msdf_hash_encode_second_backward", _lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets),
                  B, D, C, L, S, H, int(calc), _lib.ptr(dy_dx), _lib.ptr(gg_inputs), _lib.ptr(grad_grad),
                  _lib.ptr(grad2_embeddings), _lib.stream())
        # like the reference, no second-order term flows to `inputs` (hashgrid.py:101)
        return grad_grad, None, grad2_embeddings, None, None, None, None, None, None, None, None, None


hash_encode = _HashEncode.apply


def level_offsets(num_levels, base_resolution, per_level_scale, log2_hashmap_size, input_dim=3):
    """Entries per level: min(2^log2_hashmap_size, ceil(base * scale^l)^D)  (hashgrid.py:128-137)."""
    offsets, offset = [], 0
    cap = 2 ** log2_hashmap_size
    for l in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** l))
        offsets.append(offset)
        offset += min(cap, res ** input_dim)
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


class HashEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        if input_dim != 3:
            raise NotImplementedError("monosdf_b200 HashEncoder: only input_dim=3 is built")
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.log2_hashmap_size, self.base_resolution = per_level_scale, log2_hashmap_size, base_resolution
        self.output_dim = num_levels * level_dim
        self.max_params = 2 ** log2_hashmap_size
        offsets = level_offsets(num_levels, base_resolution, per_level_scale, log2_hashmap_size, input_dim)
        self.register_buffer("offsets", torch.from_numpy(offsets))
        self.n_params = int(offsets[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return (f"HashEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"base_resolution={self.base_resolution} per_level_scale={self.per_level_scale} params={tuple(self.embeddings.shape)}")

    def forward(self, inputs, size=1):
        inputs = (inputs + size) / (2 * size)
        prefix = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        out = hash_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, inputs.requires_grad)
        return out.view(prefix + [self.output_dim])
