"""Ray samplers (reference code/model/ray_sampler.py) driving the warp-per-ray CUDA kernels in csrc/sampler.cu.

The host loop is the reference's Algorithm-1 loop (ray_sampler.py:125-233): one SDF evaluation of the new samples
per round (the MLP), one kernel for merge + d* + beta bisection, one for the inverse-CDF up-sampling, and the
batch-global convergence test `beta.max() > beta0` (:179) read back once per round like the reference does.
"""
import abc

import torch

from .. import _lib


class RaySampler(metaclass=abc.ABCMeta):
    def __init__(self, near, far):
        self.near = near
        self.far = far

    @abc.abstractmethod
    def get_z_vals(self, ray_dirs, cam_loc, model):
        pass


def _rand(shape, device, rng):
    """Random draws: 'reference' = CPU generator then copy (what the reference does, ray_sampler.py:79,213),
    'device' = generated on the GPU (same distribution, no host round trip)."""
    if rng == "reference":
        return torch.rand(shape).to(device)
    return torch.rand(shape, device=device)


class UniformSampler(RaySampler):
    """ray_sampler.py:16-83; with take_sphere_intersection the far bound is the [-R,R]^3 cube exit (:48-60)."""

    def __init__(self, scene_bounding_sphere, near, N_samples, take_sphere_intersection=False, far=-1):
        super().__init__(near, 2.0 * scene_bounding_sphere * 1.75 if far == -1 else far)
        self.N_samples = N_samples
        self.scene_bounding_sphere = scene_bounding_sphere
        self.take_sphere_intersection = take_sphere_intersection
        self.rng = "device"

    def _init(self, ray_dirs, cam_loc, training, cap, beta_coef=0.0):
        N, n0, dev = ray_dirs.shape[0], self.N_samples, ray_dirs.device
        t_vals = _lib.device_constant(("linspace01", n0), dev, lambda: torch.linspace(0.0, 1.0, steps=n0))
        t_rand = _rand((N, n0), dev, self.rng).contiguous() if training else None
        z = torch.empty(N, cap, device=dev)
        beta = torch.empty(N, device=dev)
        pts = torch.empty(N * n0, 3, device=dev)
        # a cube half-width of 0 disables the slab test: far stays at self.far (take_sphere_intersection=False, :69)
        bound = self.scene_bounding_sphere if self.take_sphere_intersection else 1e30
        _lib.call("msdf_sampler_init", _lib.ptr(cam_loc), _lib.ptr(ray_dirs), N, _lib.ptr(t_vals), _lib.ptr(t_rand), n0,
                  float(bound), float(self.near), float(self.far), float(beta_coef), _lib.ptr(z), cap, _lib.ptr(beta),
                  _lib.ptr(pts), _lib.stream())
        return z, beta, pts

    def get_z_vals(self, ray_dirs, cam_loc, model):
        z, _, _ = self._init(ray_dirs.contiguous().float(), cam_loc.contiguous().float(), model.training, self.N_samples)
        return z


class ErrorBoundSampler(RaySampler):
    """VolSDF error-bounded sampling, ray_sampler.py:86-272."""

    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra, eps, beta_iters,
                 max_total_iters, inverse_sphere_bg=False, N_samples_inverse_sphere=0, add_tiny=1.0e-6, rng="device"):
        super().__init__(near, 2.0 * scene_bounding_sphere * 1.75)
        if inverse_sphere_bg:
            raise NotImplementedError("monosdf_b200: inverse_sphere_bg is not part of the MonoSDF rendering path")
        self.N_samples, self.N_samples_eval, self.N_samples_extra = N_samples, N_samples_eval, N_samples_extra
        self.uniform_sampler = UniformSampler(scene_bounding_sphere, near, N_samples_eval, take_sphere_intersection=True)
        self.eps, self.beta_iters, self.max_total_iters = eps, beta_iters, max_total_iters
        self.scene_bounding_sphere, self.add_tiny = scene_bounding_sphere, add_tiny
        self.inverse_sphere_bg = inverse_sphere_bg
        self.rng = rng
        self.last_total_iters = 0

    def get_z_vals(self, ray_dirs, cam_loc, model):
        dev = ray_dirs.device
        ray_dirs, cam_loc = ray_dirs.contiguous().float(), cam_loc.contiguous().float()
        N, n0 = ray_dirs.shape[0], self.N_samples_eval
        cap = n0 * max(1, self.max_total_iters)
        training = model.training
        with torch.no_grad():
            beta0 = model.density.get_beta().detach().reshape(1).float().contiguous()
            # Lemma-2 coefficient evaluated like the reference (fp32 log of a 0-d tensor, :119)
            beta_coef = float(1.0 / (4.0 * torch.log(torch.tensor(self.eps + 1.0))))
            self.uniform_sampler.rng = self.rng
            z, beta, pts = self.uniform_sampler._init(ray_dirs, cam_loc, training, cap, beta_coef)
            sdf = torch.empty(N, cap, device=dev)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            n_old, n_new, z_new = 0, n0, None
            sdf_new = model.implicit_network.get_sdf_vals(pts).reshape(N, n0).contiguous()
            iters = 0
            while True:
                flag.zero_()
                _lib.call("msdf_sampler_round", N, n_old, n_new, _lib.ptr(z), _lib.ptr(sdf), _lib.ptr(z_new), _lib.ptr(sdf_new),
                          cap, _lib.ptr(beta0), float(self.eps), int(self.beta_iters), _lib.ptr(beta), _lib.ptr(flag),
                          _lib.stream())
                n = n_old + n_new
                iters += 1
                not_converge = iters < self.max_total_iters and bool(flag.item())
                if not not_converge:
                    break
                u = _lib.device_constant(("linspace01", n0), dev, lambda: torch.linspace(0.0, 1.0, steps=n0))
                z_new = torch.empty(N, n0, device=dev)
                pts = torch.empty(N * n0, 3, device=dev)
                _lib.call("msdf_sampler_upsample", N, n, _lib.ptr(z), _lib.ptr(sdf), cap, _lib.ptr(beta), float(self.add_tiny),
                          _lib.ptr(u), n0, _lib.ptr(cam_loc), _lib.ptr(ray_dirs), _lib.ptr(z_new), _lib.ptr(pts), _lib.stream())
                sdf_new = model.implicit_network.get_sdf_vals(pts).reshape(N, n0).contiguous()
                n_old, n_new = n, n0
            self.last_total_iters = iters
            ns = self.N_samples
            if training:
                u = _rand((N, ns), dev, self.rng).contiguous()
                per_ray = 1
            else:
                u = _lib.device_constant(("linspace01", ns), dev, lambda: torch.linspace(0.0, 1.0, steps=ns))
                per_ray = 0
            if self.N_samples_extra > 0:
                if training:
                    pick = torch.randperm(n)[: self.N_samples_extra] if self.rng == "reference" else \
                        torch.randperm(n, device=dev)[: self.N_samples_extra]
                else:
                    ne = self.N_samples_extra
                    pick = _lib.device_constant(("pick", n, ne), dev, lambda: torch.linspace(0, n - 1, ne).long().to(torch.int32))
                pick = pick.to(device=dev, dtype=torch.int32).contiguous()
            else:
                pick = None
            n_out = ns + 2 + self.N_samples_extra
            eik_idx = (torch.randint(n_out, (N,)).to(dev) if self.rng == "reference" else
                       torch.randint(n_out, (N,), device=dev)).contiguous()
            z_out = torch.empty(N, n_out, device=dev)
            z_eik = torch.empty(N, 1, device=dev)
            _lib.call("msdf_sampler_finalize", N, n, _lib.ptr(z), _lib.ptr(sdf), cap, _lib.ptr(beta), _lib.ptr(u), per_ray, ns,
                      _lib.ptr(pick), self.N_samples_extra, float(self.near), float(self.far), _lib.ptr(eik_idx),
                      _lib.ptr(z_out), _lib.ptr(z_eik), _lib.stream())
        return z_out, z_eik
