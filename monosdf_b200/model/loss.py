"""MonoSDFLoss, the consumer directly after the rendering path (reference code/model/loss.py:180-311).

SURVEY.md section 8 row f1: on CUDA tensors the loss and its gradient with respect to the renderer's outputs are ONE
library call (msdf_loss_forward_backward, csrc/loss.cu: three launches, no host synchronisation) instead of the
reference's ~40 small kernels, boolean-mask indexing and host syncs (`det.nonzero()` :44, `if divisor == 0` :57, the
stray print :164).  `forward_torch` is the same loss in plain torch ops; it is what runs for CPU tensors (tests of the
host logic) and what tests/ compare the fused kernel with.  Same loss terms and weights; pixel-batch mode only (the
reference asserts that too, :167-168).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function

from .. import _lib


class _FusedLoss(Function):
    """(loss, terms[8]) = MonoSDFLoss(outputs, ground truth); gradients flow through `loss` only."""

    @staticmethod
    def forward(ctx, desc, rgb, depth, normal, sdf, g1, g2, rgb_gt, depth_gt, normal_gt, gt_mask):
        dev = rgb.device
        f = lambda t: t.detach().contiguous().float()
        rgb, depth, normal, sdf = f(rgb), f(depth).reshape(-1), f(normal), f(sdf)
        n, S = sdf.shape
        n_eik = 0 if g1 is None else g1.shape[0]
        g1c, g2c = (f(g1), f(g2)) if n_eik else (None, None)
        ws = torch.empty(32 + n, device=dev)
        out = torch.empty(8, device=dev)
        d_rgb, d_depth, d_normal = torch.empty_like(rgb), torch.empty(n, device=dev), torch.empty_like(normal)
        d_g1 = torch.empty_like(g1c) if n_eik else None
        d_g2 = torch.empty_like(g2c) if n_eik else None
        _lib.call("msdf_loss_forward_backward", desc, n, S, _lib.ptr(rgb), _lib.ptr(f(rgb_gt).reshape(-1, 3)), _lib.ptr(depth),
                  _lib.ptr(f(depth_gt).reshape(-1)), _lib.ptr(f(gt_mask).reshape(-1)), _lib.ptr(normal),
                  _lib.ptr(f(normal_gt).reshape(-1, 3)), _lib.ptr(sdf), n_eik, _lib.ptr(g1c), _lib.ptr(g2c), _lib.ptr(ws),
                  _lib.ptr(out), _lib.ptr(d_rgb), _lib.ptr(d_depth), _lib.ptr(d_normal), _lib.ptr(d_g1), _lib.ptr(d_g2),
                  _lib.stream())
        ctx.save_for_backward(d_rgb, d_depth, d_normal, d_g1, d_g2)
        terms = out.clone()
        ctx.mark_non_differentiable(terms)
        return out[0].clone(), terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        d_rgb, d_depth, d_normal, d_g1, d_g2 = ctx.saved_tensors
        s = lambda t: None if t is None else t * g_loss
        return (None, s(d_rgb), s(d_depth).reshape(-1, 1), s(d_normal), None, s(d_g1), s(d_g2), None, None, None, None)


def compute_scale_and_shift_1D(prediction, target, mask):
    """Closed-form least-squares scale/shift (loss.py:29-49); zero where the system is singular."""
    a_00 = torch.sum(mask * prediction * prediction, 1)
    a_01 = torch.sum(mask * prediction, 1)
    a_11 = torch.sum(mask, 1)
    b_0 = torch.sum(mask * prediction * target, 1)
    b_1 = torch.sum(mask * target, 1)
    det = a_00 * a_11 - a_01 * a_01
    ok = det != 0
    safe = torch.where(ok, det, torch.ones_like(det))
    x_0 = torch.where(ok, (a_11 * b_0 - a_01 * b_1) / safe, torch.zeros_like(det))
    x_1 = torch.where(ok, (-a_01 * b_0 + a_00 * b_1) / safe, torch.zeros_like(det))
    return x_0, x_1


class MonoSDFLoss(nn.Module):
    def __init__(self, rgb_loss="torch.nn.L1Loss", eikonal_weight=0.05, smooth_weight=0.005, depth_weight=0.1,
                 depth_alpha=0.5, normal_l1_weight=0.05, normal_cos_weight=0.05, if_gamma_loss=False,
                 if_scale_invariant_depth=True, end_step=-1):
        super().__init__()
        self.eikonal_weight, self.smooth_weight, self.depth_weight = eikonal_weight, smooth_weight, depth_weight
        self.normal_l1_weight, self.normal_cos_weight = normal_l1_weight, normal_cos_weight
        self.rgb_loss = nn.MSELoss(reduction="mean") if "MSE" in rgb_loss else nn.L1Loss(reduction="mean")
        self.if_gamma_loss, self.if_scale_invariant_depth = if_gamma_loss, if_scale_invariant_depth
        self.step, self.end_step = 0, end_step

    @staticmethod
    def gamma2(x):
        return torch.where(x <= 0.0031308, 12.92 * x, 1.055 * x.clamp_min(0.0031308).pow(1 / 2.4) - 0.055)

    def get_depth_loss(self, depth_pred, depth_gt, mask):
        pred = depth_pred.reshape(1, -1)
        tgt = (depth_gt * 50 + 0.5).reshape(1, -1) if self.if_scale_invariant_depth else depth_gt.reshape(1, -1)
        m = mask.reshape(1, -1).to(pred.dtype)
        if self.if_scale_invariant_depth:
            scale, shift = compute_scale_and_shift_1D(pred, tgt, m)
            pred = scale.view(1, -1) * pred + shift.view(1, -1)
        res = pred - tgt
        num = torch.sum(m * res * res)
        div = torch.sum(2 * m)
        return torch.where(div > 0, num / div.clamp_min(1e-30), torch.zeros_like(num))

    def forward(self, model_outputs, ground_truth, if_pixel_input=False):
        if not model_outputs["rgb_values"].is_cuda:
            return self.forward_torch(model_outputs, ground_truth, if_pixel_input)
        dev = model_outputs["rgb_values"].device
        decay = math.exp(-self.step / self.end_step * 10.0) if self.end_step > 0 else 1.0
        self.step += 1
        d = _lib.LossDesc(self.eikonal_weight, self.smooth_weight, self.depth_weight, self.normal_l1_weight,
                          self.normal_cos_weight, decay, int(isinstance(self.rgb_loss, nn.MSELoss)), int(self.if_gamma_loss),
                          int(self.if_scale_invariant_depth))
        has_eik = "grad_theta" in model_outputs
        loss, t = _FusedLoss.apply(d, model_outputs["rgb_values"], model_outputs["depth_values"], model_outputs["normal_map"],
                                   model_outputs["sdf"], model_outputs["grad_theta"] if has_eik else None,
                                   model_outputs["grad_theta_nei"] if has_eik else None, ground_truth["rgb"].to(dev),
                                   ground_truth["depth"].to(dev), ground_truth["normal"].to(dev), ground_truth["mask"].to(dev))
        return {"loss": loss, "rgb_loss": t[1], "eikonal_loss": t[2], "smooth_loss": t[3], "depth_loss": t[4],
                "normal_l1": t[5], "normal_cos": t[6]}

    def forward_torch(self, model_outputs, ground_truth, if_pixel_input=False):
        dev = model_outputs["rgb_values"].device
        rgb_gt = ground_truth["rgb"].to(dev).reshape(-1, 3)
        depth_gt, normal_gt = ground_truth["depth"].to(dev), ground_truth["normal"].to(dev)
        rgb = model_outputs["rgb_values"]
        rgb_loss = self.rgb_loss(self.gamma2(rgb), self.gamma2(rgb_gt)) if self.if_gamma_loss else self.rgb_loss(rgb, rgb_gt)
        if "grad_theta" in model_outputs:
            g1, g2 = model_outputs["grad_theta"], model_outputs["grad_theta_nei"]
            eikonal_loss = ((g1.norm(2, dim=1) - 1) ** 2).mean()
            n1 = g1 / (g1.norm(2, dim=1).unsqueeze(-1) + 1e-5)
            n2 = g2 / (g2.norm(2, dim=1).unsqueeze(-1) + 1e-5)
            smooth_loss = torch.norm(n1 - n2, dim=-1).mean()
        else:
            eikonal_loss = torch.zeros((), device=dev)
            smooth_loss = torch.zeros((), device=dev)
        sdf = model_outputs["sdf"]
        mask = ((sdf > 0.0).any(dim=-1) & (sdf < 0.0).any(dim=-1))[None, :, None]
        mask = (ground_truth["mask"].to(dev) > 0.5) & mask
        depth_loss = self.get_depth_loss(model_outputs["depth_values"], depth_gt, mask)
        n_pred = F.normalize(model_outputs["normal_map"][None] * mask, p=2, dim=-1)
        n_gt = F.normalize(normal_gt, p=2, dim=-1)
        normal_l1 = torch.abs(n_pred - n_gt).sum(dim=-1).mean()
        normal_cos = (1.0 - torch.sum(n_pred * n_gt, dim=-1)).mean()
        decay = math.exp(-self.step / self.end_step * 10.0) if self.end_step > 0 else 1.0
        self.step += 1
        loss = rgb_loss + self.eikonal_weight * eikonal_loss + self.smooth_weight * smooth_loss + \
            decay * (self.depth_weight * depth_loss + self.normal_l1_weight * normal_l1 + self.normal_cos_weight * normal_cos)
        return {"loss": loss, "rgb_loss": rgb_loss, "eikonal_loss": eikonal_loss, "smooth_loss": smooth_loss,
                "depth_loss": depth_loss, "normal_l1": normal_l1, "normal_cos": normal_cos}
