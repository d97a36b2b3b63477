"""MonoSDFLoss in plain torch ops -- TEST INFRASTRUCTURE ONLY (the checker of csrc/loss.cu; pinned to the reference's
model/loss.py:MonoSDFLoss by tests/test_loss_reference_pin.py).  `forward_torch(loss_module, outputs, ground_truth)`
evaluates the loss configured on a monosdf_b200.model.loss.MonoSDFLoss instance; works on CPU and CUDA tensors, any
float dtype.  Nothing under monosdf_b200/ imports this file."""
import math

import torch
import torch.nn.functional as F


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


def forward_torch(self, model_outputs, ground_truth, if_pixel_input=False):
    dev = model_outputs["rgb_values"].device
    rgb_gt = ground_truth["rgb"].to(dev).reshape(-1, 3)
    depth_gt, normal_gt = ground_truth["depth"].to(dev), ground_truth["normal"].to(dev)
    rgb = model_outputs["rgb_values"]
    rgb_loss = self.rgb_loss(gamma2(rgb), gamma2(rgb_gt)) if self.if_gamma_loss else self.rgb_loss(rgb, rgb_gt)
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
    depth_loss = get_depth_loss(self, model_outputs["depth_values"], depth_gt, mask)
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
