"""Import shim for the REAL reference (`/root/reference/code`) -- TEST INFRASTRUCTURE ONLY.

Usable in the build container, where /root/reference exists, and on the GPU box through the files
oracle/stage_reference.py staged under baseline/_ref/.  Used by oracle/make_golden.py to produce the committed
fixtures under tests/golden/ and by tests that pin oracle/port.py against the
unmodified reference.  Nothing in monosdf_b200/ may import this file.

What the shim does:
  * stubs side-imports the hot path never calls: matplotlib(.pyplot),
    tkinter.messagebox (ray_sampler.py:2 imports `NO`), imageio, skimage,
    and hashencoder.backend (hashgrid.py:12 would JIT-compile CUDA at import);
  * replaces pyhocon's ConfigTree with a dict-backed stand-in;
  * on CPU makes Tensor.cuda / Module.cuda identity (the reference hard-codes
    .cuda(), e.g. network.py:484, ray_sampler.py:65-79, density.py:19).
"""
import os
import sys
import types
import contextlib

def _find_root():
    """/root/reference in the build container; on the GPU box the copy oracle/stage_reference.py left under
    baseline/_ref/ (git-ignored, travels with the gpurun snapshot)."""
    cands = [os.environ.get("MSDF_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "code", "model", "network.py")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()
REF_CODE = os.path.join(REF_ROOT, "code")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_CODE, "model", "network.py"))


class Conf(dict):
    """dict-backed stand-in for pyhocon.ConfigTree (get_int/get_float/...)."""

    def _get(self, key, default=None, required=True):
        if key in self:
            return self[key]
        if default is not None or not required:
            return default
        raise KeyError(key)

    def get_int(self, k, default=None):
        return int(self._get(k, default))

    def get_float(self, k, default=None):
        return float(self._get(k, default))

    def get_bool(self, k, default=None):
        v = self._get(k, default, required=default is None)
        return bool(v)

    def get_string(self, k, default=None):
        return str(self._get(k, default))

    def get_list(self, k, default=None):
        return list(self._get(k, default))

    def get_config(self, k, default=None):
        v = self._get(k, default)
        return v if isinstance(v, Conf) else Conf(v)


def to_conf(d):
    out = Conf()
    for k, v in d.items():
        out[k] = to_conf(v) if isinstance(v, dict) else v
    return out


_loaded = {}


def load_reference(cpu_identity_cuda: bool = True):
    """Return the reference's `model.network` module (imported once)."""
    if "network" in _loaded:
        return _loaded["network"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    mpl = stub("matplotlib")
    plt = stub("matplotlib.pyplot")
    mpl.pyplot = plt
    tk = stub("tkinter")
    mb = stub("tkinter.messagebox", NO="no")
    tk.messagebox = mb
    stub("imageio")
    stub("skimage")
    if not torch.cuda.is_available() and cpu_identity_cuda:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    # hashencoder.backend JIT-builds CUDA on import; the MLP confs never call it.
    # (the stub must be registered before the package's __init__ runs.)
    import importlib
    be = types.ModuleType("hashencoder.backend")
    be._backend = None
    sys.modules["hashencoder.backend"] = be
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        net = importlib.import_module("model.network")
    _loaded["network"] = net
    return net


MLP_CONF = {
    "feature_vector_size": 256,
    "scene_bounding_sphere": 1.1,
    "Grid_MLP": False,
    "implicit_network": {
        "d_in": 3, "d_out": 1, "dims": [256] * 8, "geometric_init": True, "bias": 0.9,
        "skip_in": [4], "weight_norm": True, "multires": 6, "inside_outside": True,
    },
    "rendering_network": {
        "mode": "idr", "d_in": 9, "d_out": 3, "dims": [256, 256], "weight_norm": True,
        "multires_view": 4, "per_image_code": False,
    },
    "density": {"params_init": {"beta": 0.1}, "beta_min": 0.0001},
    "ray_sampler": {
        "near": 0.0, "N_samples": 64, "N_samples_eval": 128, "N_samples_extra": 32,
        "eps": 0.1, "beta_iters": 10, "max_total_iters": 5,
    },
}
