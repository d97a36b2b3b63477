"""Data-parallel training step plumbing: flat parameter / gradient arenas, one gradient all-reduce, fused Adam.

Replaces, for the rendering hot path, the reference's DistributedDataParallel wrap (code/training/monosdf_train.py:
228-229: bucketed gradient all-reduce, mean over ranks) and its torch.optim.Adam with three parameter groups
(:210-221: hash table lr x20, MLPs, density beta).  Rays shard across ranks (each rank renders its own rays, like
the reference where every rank draws its own pixels); the only exchange is ONE all-reduce of the flat gradient
arena per step, followed by a fused Adam kernel per group that folds the 1/world_size in.
"""
import torch
import torch.distributed as dist

from . import _lib


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of n_total rays for `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatArena:
    """Re-homes the parameters of `groups` (lists of nn.Parameter) into one flat fp32 buffer, with a matching flat
    gradient buffer whose slices are the parameters' .grad.  Parameter names / shapes are untouched, so state_dict
    keys stay the reference's."""

    def __init__(self, groups):
        params = [p for g in groups for p in g]
        assert len({id(p) for p in params}) == len(params), "a parameter appears in two groups"
        dev = params[0].device
        self.group_ranges = []
        n = 0
        for g in groups:
            start = n
            for p in g:
                n += (p.numel() + 3) // 4 * 4      # keep every slice 16-byte aligned
            self.group_ranges.append((start, n))
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        self.slices = []
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            self.slices.append((off, k))
            off += (k + 3) // 4 * 4
        self.params = params

    def zero_grad(self):
        self.grad.zero_()
        for p, (off, k) in zip(self.params, self.slices):   # autograd may have replaced .grad; re-attach the views
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + k].view(p.shape)

    def all_reduce(self, group=None):
        """Sum the gradient arena over ranks (one collective); the mean's 1/world is folded into the Adam step."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1


class FusedAdam:
    """torch.optim.Adam semantics (no amsgrad) over a FlatArena, one kernel launch per parameter group."""

    def __init__(self, arena, lrs, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        assert len(lrs) == len(arena.group_ranges)
        self.arena, self.lrs, self.betas, self.eps, self.weight_decay = arena, list(lrs), betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(arena.flat)
        self.exp_avg_sq = torch.zeros_like(arena.flat)
        self.step_count = 0

    def step(self, grad_scale=1.0):
        self.step_count += 1
        _lib.param_epoch[0] += 1          # parameter memory changes without the tensors' version counters noticing
        a = self.arena
        # model.to() / .cuda() / .float() after build_optimizer() re-allocates parameter storage: Adam would then update
        # an arena the model no longer reads
        for p, (off, k) in zip(a.params, a.slices):
            if p.data_ptr() != a.flat.data_ptr() + 4 * off:
                raise RuntimeError("monosdf_b200: a parameter no longer lives in the optimizer's flat arena "
                                   "(module moved / cast after build_optimizer?)")
        for (lo, hi), lr in zip(a.group_ranges, self.lrs):
            if hi == lo:
                continue
            _lib.call("msdf_fused_adam", a.flat.data_ptr() + 4 * lo, a.grad.data_ptr() + 4 * lo, self.exp_avg.data_ptr() + 4 * lo,
                      self.exp_avg_sq.data_ptr() + 4 * lo, hi - lo, float(lr), float(self.betas[0]), float(self.betas[1]),
                      float(self.eps), float(self.weight_decay), self.step_count, float(grad_scale), _lib.stream())

    def scale_lr(self, factor):
        self.lrs = [lr * factor for lr in self.lrs]


class ExponentialLR:
    """torch.optim.lr_scheduler.ExponentialLR for a FusedAdam (monosdf_train.py:223-226, stepped once per iteration
    :480): every group's learning rate is multiplied by gamma per step()."""

    def __init__(self, optimizer, gamma):
        self.optimizer, self.gamma = optimizer, float(gamma)
        self.last_epoch = 0
        self.base_lrs = list(optimizer.lrs)

    def step(self):
        self.last_epoch += 1
        self.optimizer.scale_lr(self.gamma)

    def get_last_lr(self):
        return list(self.optimizer.lrs)

    def state_dict(self):
        return {"gamma": self.gamma, "last_epoch": self.last_epoch, "base_lrs": self.base_lrs, "_last_lr": self.get_last_lr()}

    def load_state_dict(self, sd):
        self.gamma, self.last_epoch, self.base_lrs = float(sd["gamma"]), int(sd["last_epoch"]), list(sd["base_lrs"])
        self.optimizer.lrs = list(sd["_last_lr"])


def build_optimizer(model, lr=5.0e-4, grid_lr_factor=20.0):
    """Parameter groups of the reference trainer (monosdf_train.py:210-221) on a flat arena."""
    inet = model.implicit_network
    if getattr(model, "Grid_MLP", False):
        groups = [list(inet.grid_parameters()), list(inet.mlp_parameters()) + list(model.rendering_network.parameters()),
                  list(model.density.parameters())]
        lrs = [lr * grid_lr_factor, lr, lr]
        betas, eps = (0.9, 0.99), 1e-15
    else:
        groups = [list(model.parameters())]
        lrs = [lr]
        betas, eps = (0.9, 0.999), 1e-8
    arena = FlatArena(groups)
    return arena, FusedAdam(arena, lrs, betas=betas, eps=eps)
