"""Density modules (reference code/model/density.py).  Inside MonoSDFNetwork the Laplace density is evaluated by
the compositing kernels (csrc/render.cu) and the sampler kernels (csrc/sampler.cu); the module holds the
learnable beta and offers the same stand-alone call for API compatibility."""
import torch
import torch.nn as nn


class Density(nn.Module):
    def __init__(self, params_init={}):
        super().__init__()
        for p in params_init:
            setattr(self, p, nn.Parameter(torch.tensor(params_init[p])))

    def forward(self, sdf, beta=None):
        return self.density_func(sdf, beta=beta)


class LaplaceDensity(Density):
    """alpha * Laplace(loc=0, scale=beta).cdf(-sdf), density.py:16-30"""

    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__(params_init=params_init)
        self.register_buffer("_beta_min", torch.tensor(float(beta_min)), persistent=False)

    @property
    def beta_min(self):
        return self._beta_min

    def density_func(self, sdf, beta=None):
        if beta is None:
            beta = self.get_beta()
        return (1 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))

    def get_beta(self):
        return self.beta.abs() + self._beta_min
