"""ctypes binding of libmonosdf_b200.so (C ABI declared in include/monosdf_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint, c_uint32, c_ulonglong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmonosdf_b200.so")
MAX_LAYERS = 12

MODE_SDF_ONLY, MODE_FORWARD, MODE_BACKWARD = 0, 1, 2
FLAG_TENSOR_BF16 = 1


class MlpDesc(Structure):
    _fields_ = [("n_layers", c_int32), ("d0", c_int32), ("skip_layer", c_int32),
                ("in_dim", c_int32 * MAX_LAYERS), ("out_dim", c_int32 * MAX_LAYERS), ("ldw", c_int32 * MAX_LAYERS),
                ("W", c_void_p * MAX_LAYERS), ("b", c_void_p * MAX_LAYERS)]


class MlpGrads(Structure):
    _fields_ = [("dW", c_void_p * MAX_LAYERS), ("db", c_void_p * MAX_LAYERS)]


class EncodingDesc(Structure):
    _fields_ = [("multires", c_int32), ("grid_feat_dim", c_int32), ("n_levels", c_int32), ("level_dim", c_int32),
                ("base_res", c_int32), ("log2_per_level_scale", c_float), ("divide_factor", c_float),
                ("table", c_void_p), ("offsets", c_void_p)]


class LossDesc(Structure):
    _fields_ = [("eikonal_weight", c_float), ("smooth_weight", c_float), ("depth_weight", c_float),
                ("normal_l1_weight", c_float), ("normal_cos_weight", c_float), ("decay", c_float),
                ("rgb_mse", c_int32), ("gamma", c_int32), ("scale_invariant_depth", c_int32)]


class ColorDesc(Structure):
    _fields_ = [("mode_idr", c_int32), ("multires_view", c_int32), ("feat_dim", c_int32), ("code_dim", c_int32),
                ("code_per_ray", c_int32), ("final_act", c_int32), ("spec", c_int32)]


_P = c_void_p
_SIGNATURES = {
    "msdf_last_error": (c_char_p, []),
    "msdf_abi_version": (c_int, []),
    "msdf_launch_count": (c_ulonglong, []),
    "msdf_sampler_init": (c_int, [_P, _P, c_int64, _P, _P, c_int, c_float, c_float, c_float, c_float, _P, c_int, _P, _P, _P]),
    "msdf_sampler_round": (c_int, [c_int64, c_int, c_int, _P, _P, _P, _P, c_int, _P, c_float, c_int, _P, _P, _P]),
    "msdf_sampler_upsample": (c_int, [c_int64, c_int, _P, _P, c_int, _P, c_float, _P, c_int, _P, _P, _P, _P, _P]),
    "msdf_sampler_finalize": (c_int, [c_int64, c_int, _P, _P, c_int, _P, _P, c_int, c_int, _P, c_int, c_float, c_float, _P, _P, _P, _P]),
    "msdf_hash_encode_forward": (c_int, [_P, _P, _P, _P, c_uint32, c_uint32, c_uint32, c_uint32, c_float, c_uint32, c_int, _P, _P]),
    "msdf_hash_encode_backward": (c_int, [_P, _P, _P, _P, _P, c_uint32, c_uint32, c_uint32, c_uint32, c_float, c_uint32, c_int, _P, _P, _P]),
    "msdf_hash_encode_second_backward": (c_int, [_P, _P, _P, _P, c_uint32, c_uint32, c_uint32, c_uint32, c_float, c_uint32, c_int, _P, _P, _P, _P, _P]),
    "msdf_field_workspace_bytes": (c_size_t, [POINTER(MlpDesc), POINTER(EncodingDesc), POINTER(MlpDesc), POINTER(ColorDesc), c_int64, c_int, c_uint]),
    "msdf_field_forward": (c_int, [POINTER(MlpDesc), POINTER(EncodingDesc), POINTER(MlpDesc), POINTER(ColorDesc), _P, c_int64, _P, c_int64,
                                   c_int, _P, c_int, c_float, c_float, c_uint, _P, c_size_t, _P, _P, _P, c_int64, _P, _P, c_size_t, _P]),
    "msdf_field_saved_bytes": (c_size_t, [POINTER(MlpDesc), POINTER(EncodingDesc), POINTER(MlpDesc), POINTER(ColorDesc), c_int64, c_int, c_uint]),
    "msdf_field_backward": (c_int, [POINTER(MlpDesc), POINTER(EncodingDesc), POINTER(MlpDesc), POINTER(ColorDesc), _P, c_int64, _P, c_int64,
                                    c_int, _P, c_float, c_float, c_uint, _P, c_size_t, _P, _P, _P, c_int64, _P, _P,
                                    POINTER(MlpGrads), POINTER(MlpGrads), _P, _P, _P, c_size_t, _P]),
    "msdf_ray_points": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "msdf_camera_rays": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P]),
    "msdf_sdfgrid_level_points": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    "msdf_sdfgrid_level_assemble": (c_int, [c_int, _P, _P, _P, c_float, _P, _P, _P]),
    "msdf_pixel_batch": (c_int, [_P, c_int64, _P, _P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "msdf_render_forward": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, _P, c_int64, _P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "msdf_render_backward": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, _P, c_int64, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "msdf_weightnorm_forward": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P]),
    "msdf_weightnorm_backward": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "msdf_loss_forward_backward": (c_int, [POINTER(LossDesc), c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P,
                                           _P, _P, _P, _P, _P, _P]),
    "msdf_code_scatter": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P]),
    "msdf_fused_adam": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int64, c_float, _P]),
    "msdf_tc_selftest": (c_int, [c_int, POINTER(c_float), _P]),
    "msdf_tc_selftest_count": (c_int, []),
    "msdf_set_fused": (None, [c_int]),
    "msdf_set_sweeps": (None, [c_int, c_int]),
    "msdf_profile_enable": (c_int, [c_int]),
    "msdf_profile_read": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(ctypes.c_longlong), c_int]),
    "msdf_profile_read_bytes": (c_int, [c_int, POINTER(ctypes.c_double)]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """The loaded shared library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "monosdf_b200: %s is missing -- build it with `python -m monosdf_b200.build` "
                "(there is no CPU / PyTorch fallback for the rendering hot path)" % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if h.msdf_abi_version() != 3:
            raise RuntimeError("monosdf_b200: ABI version mismatch")
        _lib = h
    return _lib


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError("monosdf_b200.%s failed (%d): %s" % (name, rc, lib().msdf_last_error().decode()))


def launch_count():
    return int(lib().msdf_launch_count())


def ptr(t):
    """Device pointer of a CUDA fp32/int tensor (None -> NULL). Refuses CPU tensors: there is no CPU path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("monosdf_b200: expected a CUDA tensor (the hot path has no CPU implementation)")
    if not t.is_contiguous():
        raise RuntimeError("monosdf_b200: expected a contiguous tensor")
    return t.data_ptr()


def stream():
    """Raw handle of torch's CURRENT stream (torch.cuda.current_stream() builds a Python Stream object: 15 us per call,
    a third of a millisecond per 1024-ray eval chunk; the raw getter is the one Triton's launcher uses)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:      # pragma: no cover  (older / newer torch without the private getter)
        return torch.cuda.current_stream().cuda_stream


# bumped by everything that rewrites parameter memory behind autograd's back (training.FusedAdam.step: a raw kernel on the
# flat arena does not touch the tensors' version counters); part of the key of the effective-weight cache
param_epoch = [0]

_consts = {}


def device_constant(key, device, make):
    """Small constant tensors the host loops re-create every call (linspace grids ...), made once per device."""
    k = (key, device.type, device.index)
    t = _consts.get(k)
    if t is None:
        t = make().to(device)
        _consts[k] = t
    return t


_workspaces = {}


def workspace(nbytes, device):
    """A cached scratch buffer of at least nbytes on `device` (grown on demand, reused across calls)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspaces.pop(key, None)
        buf = None
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


class _SavedPool:
    """Recycles the multi-GB saved-activation buffers across training steps.  Handing them back to torch's caching
    allocator lets it split the block for small tensors, and the next step's 30-60 GB request then falls through to a
    synchronous cudaMalloc (measured: +70-90 ms on every other step)."""

    def __init__(self):
        self.free = []

    def acquire(self, nbytes, device):
        best = None
        for i, t in enumerate(self.free):
            if t.device == device and t.numel() >= nbytes and (best is None or t.numel() < self.free[best].numel()):
                best = i
        if best is not None and self.free[best].numel() <= 1.5 * nbytes + (64 << 20):
            return self.free.pop(best)
        # nothing fits: buffers of other sizes (ragged last batch, eval chunks, another shard size) would only pile up
        # next to the new one -- hand them back before asking for more
        self.free = [t for t in self.free if t.device != device]
        return torch.empty(int(nbytes), dtype=torch.uint8, device=device)

    def best_fit_bytes(self, nbytes, device):
        """size of the pooled buffer acquire() would hand out for this request (0: none) -- the only pooled memory a
        caller may count as free"""
        fits = [t.numel() for t in self.free if t.device == device and nbytes <= t.numel() <= 1.5 * nbytes + (64 << 20)]
        return min(fits) if fits else 0

    def release(self, t):
        if t is not None:
            self.free.append(t)
            if len(self.free) > 3:           # a step needs two (render + eikonal); keep the largest
                self.free.sort(key=lambda x: -x.numel())
                del self.free[3:]

    def clear(self):
        self.free.clear()


saved_pool = _SavedPool()


def profile_enable(on):
    call("msdf_profile_enable", int(bool(on)))


def profile_read(cls, reset=False):
    """(total_ms, total_flops, count, total_algorithmic_bytes) of the recorded launches of one kernel class."""
    ms, work, n, nbytes = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    torch.cuda.synchronize()
    call("msdf_profile_read_bytes", int(cls), ctypes.byref(nbytes))
    call("msdf_profile_read", int(cls), ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n), int(bool(reset)))
    return ms.value, work.value, n.value, nbytes.value
