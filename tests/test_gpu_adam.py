"""msdf_fused_adam + training.ExponentialLR against torch.optim.Adam + torch.optim.lr_scheduler.ExponentialLR with the
reference trainer's parameter groups (code/training/monosdf_train.py:210-226: hash table lr x20, betas (0.9, 0.99),
eps 1e-15 for Grid_MLP models; plain Adam(lr) otherwise), including the 1/world gradient scale the all-reduce path
folds into the step."""
import pytest
import torch

from monosdf_b200 import training
from tests.helpers import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference_optimizer(model, lr, factor):
    if model.Grid_MLP:       # monosdf_train.py:210-219
        return torch.optim.Adam([
            {"name": "encoding", "params": list(model.implicit_network.grid_parameters()), "lr": lr * factor},
            {"name": "net", "params": list(model.implicit_network.mlp_parameters()) + list(model.rendering_network.parameters()), "lr": lr},
            {"name": "density", "params": list(model.density.parameters()), "lr": lr},
        ], betas=(0.9, 0.99), eps=1e-15)
    return torch.optim.Adam(model.parameters(), lr=lr)      # :221


@pytest.mark.parametrize("case", ["mlp_small", "gridmlp_small"])
@pytest.mark.parametrize("world", [1, 4])
def test_fused_adam_matches_torch_adam(golden, case, world):
    fx = golden(case)
    ours = build_model(fx, DEV)
    ref = build_model(fx, DEV)        # same seed -> identical parameters (weight_norm modules do not deepcopy)
    for p, q in zip(ours.parameters(), ref.parameters()):
        assert torch.equal(p, q)
    lr, factor, gamma = 5.0e-4, 20.0, 0.1 ** (1.0 / 50.0)
    arena, opt = training.build_optimizer(ours, lr=lr, grid_lr_factor=factor)
    sched = training.ExponentialLR(opt, gamma)
    ropt = _reference_optimizer(ref, lr, factor)
    rsched = torch.optim.lr_scheduler.ExponentialLR(ropt, gamma)
    names = [n for n, _ in ours.named_parameters()]
    g = torch.Generator().manual_seed(11)
    for step in range(10):
        arena.zero_grad()
        ropt.zero_grad()
        for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
            # gradients over many decades (eps = 1e-15 matters for the small ones)
            scale = 10.0 ** float(torch.randint(-9, 1, (1,), generator=g))
            gr = (torch.randn(p.shape, generator=g) * scale).to(DEV)
            p.grad.copy_(gr * world)          # what the sum all-reduce over `world` ranks would leave in the arena
            q.grad = gr.clone()
        opt.step(grad_scale=1.0 / world)
        sched.step()
        ropt.step()
        rsched.step()
        assert sched.get_last_lr() == pytest.approx(rsched.get_last_lr(), rel=1e-12)
    worst = 0.0
    for n, (p, q) in zip(names, zip(ours.parameters(), ref.parameters())):
        err = float((p - q).abs().max() / q.abs().max().clamp_min(1e-12))
        worst = max(worst, err)
        assert err < 2e-6, (n, err)
    print("REPORT fused adam vs torch.optim.Adam (%s, world %d): max relative parameter difference %.2e" % (case, world, worst))


def test_fused_adam_refuses_a_moved_model(golden):
    fx = golden("mlp_small")
    model = build_model(fx, DEV)
    arena, opt = training.build_optimizer(model)
    model.double().float()        # re-allocates every parameter
    arena.zero_grad()
    with pytest.raises(RuntimeError):
        opt.step()
