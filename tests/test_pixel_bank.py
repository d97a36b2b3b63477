"""Pixel-mode batch assembly: oracle/port.pixel_bank + pixel_batch restate
SceneDatasetDN.convert_to_pixels / __getitem__ / collate_fn (datasets/scene_dataset.py:258-307, 374-401, 438-464);
monosdf_b200.data.DevicePixelBank (msdf_pixel_batch) must reproduce them: indices and gathered ground truth bit for bit,
directions to fp32 rounding (1e-6)."""
import os
import sys

import pytest
import torch

from oracle import port
from tests.helpers import rel_err

REF = "/root/reference/code"


def _bank_inputs(F=3, H=6, W=9, seed=5):
    g = torch.Generator().manual_seed(seed)
    poses = torch.eye(4)[None].repeat(F, 1, 1)
    poses[:, :3, :3] = torch.linalg.qr(torch.randn(F, 3, 3, generator=g))[0]
    poses[:, :3, 3] = torch.randn(F, 3, generator=g) * 0.3
    intr = torch.eye(4)[None].repeat(F, 1, 1)
    intr[:, 0, 0] = 11.0 + torch.rand(F, generator=g)
    intr[:, 1, 1] = 10.0 + torch.rand(F, generator=g)
    intr[:, 0, 1] = 0.05                                    # a little skew: exercises every term of lift()
    intr[:, 0, 2], intr[:, 1, 2] = float(W // 2), float(H // 2)
    images = {"rgb": torch.rand(F, H * W, 3, generator=g), "depth": torch.rand(F, H * W, 1, generator=g),
              "mask": (torch.rand(F, H * W, 1, generator=g) > 0.3).float(), "normal": torch.randn(F, H * W, 3, generator=g)}
    return poses, intr, H, W, images


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_oracle_pixel_bank_matches_reference_camera_model():
    """The oracle's per-ray arrays against the reference's own rend_util.get_camera_params on the reference's uv grid."""
    from oracle import ref_shim
    ref_shim.load_reference()
    from utils import rend_util
    import numpy as np
    poses, intr, H, W, _ = _bank_inputs()
    bank = port.pixel_bank(poses, intr, H, W)
    uv = np.mgrid[0:H, 0:W].astype(np.int32)
    uv = torch.from_numpy(np.flip(uv, axis=0).copy()).float().reshape(2, -1).transpose(1, 0)
    uv_all = uv.unsqueeze(0).expand(poses.shape[0], -1, -1)
    dirs, loc = rend_util.get_camera_params(uv_all, poses, intr)
    dirs_tmp, _ = rend_util.get_camera_params(uv_all, torch.eye(4)[None].expand(poses.shape[0], -1, -1), intr)
    assert torch.equal(bank["ray_dirs"], dirs.reshape(-1, 3))
    assert torch.equal(bank["ray_dirs_tmp"], dirs_tmp.reshape(-1, 3))
    assert torch.equal(bank["ray_cam_loc"].reshape(poses.shape[0], H * W, 3)[:, 0], loc)


def test_oracle_pixel_batch_layout():
    poses, intr, H, W, images = _bank_inputs()
    bank = port.pixel_bank(poses, intr, H, W)
    ids = torch.tensor([0, H * W - 1, H * W, 2 * H * W + 7, 5])
    idx, inp, gt = port.pixel_batch(bank, images, ids)
    assert idx.tolist() == [0, 0, 1, 2, 0] and idx.dtype == torch.int64
    assert inp["ray_pose"].shape == (5, 4, 4) and torch.equal(inp["ray_pose"][3], poses[2])
    assert torch.equal(gt["rgb"][3], images["rgb"][2, 7]) and gt["depth"].shape == (5, 1)
    # pixel p = row p // W, column p % W; the un-rotated direction of the principal point looks down +z
    cx, cy = int(intr[0, 0, 2]), int(intr[0, 1, 2])
    centre = port.pixel_batch(bank, images, torch.tensor([cy * W + cx]))[1]["ray_dirs_tmp"][0]
    assert torch.allclose(centre, torch.tensor([0.0, 0.0, 1.0]), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("F,H,W,n", [(3, 6, 9, 100), (1, 1, 1, 4), (5, 48, 64, 20000)])
def test_device_pixel_bank_matches_oracle(F, H, W, n):
    from monosdf_b200.data import DevicePixelBank
    poses, intr, H, W, images = _bank_inputs(F, H, W)
    bank_o = port.pixel_bank(poses, intr, H, W)
    g = torch.Generator().manual_seed(17)
    ids = torch.randint(0, F * H * W, (n,), generator=g)
    idx_o, inp_o, gt_o = port.pixel_batch(bank_o, images, ids)
    bank = DevicePixelBank(poses, intr, (H, W), **images)
    idx, inp, gt = bank.batch(ids)
    bank.check()
    assert idx.dtype == torch.int64 and torch.equal(idx.cpu(), idx_o)                 # index work: bit-exact
    for k in ("ray_cam_loc", "ray_pose"):
        assert torch.equal(inp[k].cpu(), inp_o[k]), k                                  # copies: bit-exact
    for k in gt_o:
        assert gt[k].shape == gt_o[k].shape and torch.equal(gt[k].cpu(), gt_o[k]), k   # gathers: bit-exact
    for k in ("ray_dirs", "ray_dirs_tmp"):
        assert rel_err(inp[k], inp_o[k]) < 1e-6, k                                     # fp32 lift + normalise
    # ragged / empty / out-of-range ids
    assert bank.batch(ids[:0])[0].numel() == 0
    bank.batch(torch.tensor([F * H * W]))
    with pytest.raises(IndexError):
        bank.check()


@pytest.mark.gpu
def test_device_pixel_bank_epoch_covers_the_sampled_pixels():
    """change_sampling_idx + batches(): an epoch is a permutation prefix, cut into num_pixels-sized batches, that feeds
    MonoSDFNetwork.forward directly (same dict keys as the reference's collate_fn)."""
    from monosdf_b200.data import DevicePixelBank
    poses, intr, H, W, images = _bank_inputs(4, 16, 16)
    bank = DevicePixelBank(poses, intr, (H, W), **images)
    gen = torch.Generator().manual_seed(3)
    bank.change_sampling_idx(128, generator=gen)                      # 128 of 256 pixels per image -> half of all rays
    assert len(bank) == 4 * 128
    expect = torch.randperm(4 * H * W, generator=torch.Generator().manual_seed(3))[:len(bank)]
    assert torch.equal(bank.sampling_idx.cpu(), expect)
    seen, sizes = [], []
    for idx, inp, gt in bank.batches(200):
        sizes.append(idx.numel())
        assert set(inp) == {"ray_dirs", "ray_dirs_tmp", "ray_cam_loc", "ray_pose"} and set(gt) == {"rgb", "depth", "mask", "normal"}
        seen.append(gt["rgb"].cpu())
    assert sizes == [200, 200, 112]
    assert torch.equal(torch.cat(seen), images["rgb"].reshape(-1, 3)[expect])
    bank.change_sampling_idx(-1)
    assert len(bank) == 4 * H * W
