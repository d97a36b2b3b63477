"""ctypes front end of oracle/sampler_oracle.c -- TEST INFRASTRUCTURE ONLY.

Drives the three oracle phases exactly like monosdf_b200.ray_sampler.ErrorBoundSampler drives
the CUDA kernels (same host-side loop, same random draws in the reference's order,
ray_sampler.py:79,213,244,254), with the SDF supplied by `sdf_fn` on host tensors.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmsdf_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f(x):
    return ctypes.c_float(float(x))


def beta_coef(eps: float) -> float:
    """(1/(4 log(1+eps))) evaluated like ray_sampler.py:119 (fp32 torch log on a 0-d tensor)."""
    return float(1.0 / (4.0 * torch.log(torch.tensor(eps + 1.0))))


class OracleSampler:
    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra, eps, beta_iters,
                 max_total_iters, add_tiny=1.0e-6):
        self.R = scene_bounding_sphere
        self.near, self.far = near, 2.0 * scene_bounding_sphere * 1.75
        self.N_samples, self.N_eval, self.N_extra = N_samples, N_samples_eval, N_samples_extra
        self.eps, self.beta_iters, self.max_iters, self.add_tiny = eps, beta_iters, max_total_iters, add_tiny

    def get_z_vals(self, ray_dirs, cam_loc, sdf_fn, beta0: float, training: bool, trace=None):
        L = lib()
        N, n0 = ray_dirs.shape[0], self.N_eval
        cap = n0 * self.max_iters
        o, d = cam_loc.contiguous().float(), ray_dirs.contiguous().float()
        t_vals = torch.linspace(0.0, 1.0, steps=n0)
        t_rand = torch.rand(N, n0) if training else None
        z = torch.zeros(N, cap); sdf = torch.zeros(N, cap); beta = torch.zeros(N)
        pts = torch.empty(N * n0, 3)
        L.msdf_oracle_sampler_init(_p(o), _p(d), ctypes.c_int64(N), _p(t_vals), _p(t_rand), n0, _f(self.R),
                                   _f(self.near), _f(self.far), _f(beta_coef(self.eps)), _p(z), cap, _p(beta), _p(pts))
        n_old, n_new, z_new = 0, n0, None
        sdf_new = sdf_fn(pts).reshape(N, n0).contiguous().float()
        iters = 0
        flag = torch.zeros(1, dtype=torch.int32)
        while True:
            flag.zero_()
            L.msdf_oracle_sampler_round(ctypes.c_int64(N), n_old, n_new, _p(z), _p(sdf), _p(z_new), _p(sdf_new), cap,
                                        _f(beta0), _f(self.eps), self.beta_iters, _p(beta), _p(flag))
            n = n_old + n_new
            iters += 1
            if trace is not None:
                trace.setdefault("rounds", []).append(dict(z=z[:, :n].clone(), sdf=sdf[:, :n].clone(), beta=beta.clone()))
            if bool(flag.item()) and iters < self.max_iters:
                u = torch.linspace(0.0, 1.0, steps=self.N_eval)
                z_new = torch.empty(N, self.N_eval); pts = torch.empty(N * self.N_eval, 3)
                L.msdf_oracle_sampler_upsample(ctypes.c_int64(N), n, _p(z), _p(sdf), cap, _p(beta), _f(self.add_tiny),
                                               _p(u), self.N_eval, _p(o), _p(d), _p(z_new), _p(pts))
                sdf_new = sdf_fn(pts).reshape(N, self.N_eval).contiguous().float()
                n_old, n_new = n, self.N_eval
                continue
            if training:
                u = torch.rand(N, self.N_samples); per_ray = 1
                pick = torch.randperm(n)[: self.N_extra]
            else:
                u = torch.linspace(0.0, 1.0, steps=self.N_samples); per_ray = 0
                pick = torch.linspace(0, n - 1, self.N_extra).long()
            pick = pick.to(torch.int32).contiguous()
            n_out = self.N_samples + 2 + self.N_extra
            eik_idx = torch.randint(n_out, (N,))
            z_out = torch.empty(N, n_out); z_eik = torch.empty(N, 1)
            L.msdf_oracle_sampler_finalize(ctypes.c_int64(N), n, _p(z), _p(sdf), cap, _p(beta), _p(u), per_ray,
                                           self.N_samples, _p(pick), self.N_extra, _f(self.near), _f(self.far),
                                           _p(eik_idx), _p(z_out), _p(z_eik))
            if trace is not None:
                trace.update(total_iters=iters, beta=beta.clone(), z_dense=z[:, :n].clone())
            return z_out, z_eik
