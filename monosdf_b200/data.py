"""Device-resident pixel-mode batches ("next" row f3 of SURVEY.md section 8).

Replaces, for `if_pixel_train` runs, SceneDatasetDN.convert_to_pixels / __getitem__ / collate_fn and the DataLoader
around them (reference code/datasets/scene_dataset.py:269-307, 374-401, 438-464, 468-478; training/monosdf_train.py:
180-184, 420): the reference materialises ray_dirs / ray_dirs_tmp / ray_cam_loc / ray_pose (64 B) / ground truth for
EVERY pixel of every frame on the host and lets 8 workers build each batch ray by ray through `__getitem__` and
`torch.stack`.  Here the per-frame data (poses, intrinsics, images) stays on the GPU and one kernel
(`msdf_pixel_batch`, csrc/rays.cu) produces a whole batch -- the same dictionaries, already on the device.
"""
import torch

from . import _lib


class DevicePixelBank:
    """poses, intrinsics [F,4,4]; img_res (H, W); rgb / normal [F, H*W, 3], depth / mask [F, H*W, 1] (or None).

    Ray id r = f * H*W + p addresses pixel p (row p // W, column p % W) of frame f, the order of the reference's
    flattened per-ray arrays (scene_dataset.py:283-302)."""

    def __init__(self, poses, intrinsics, img_res, rgb=None, depth=None, mask=None, normal=None, device="cuda"):
        dev = torch.device(device)
        f32 = lambda t: None if t is None else t.to(dev).float().contiguous()   # noqa: E731
        self.poses, self.intrinsics = f32(poses), f32(intrinsics)
        self.H, self.W = int(img_res[0]), int(img_res[1])
        self.n_frames = self.poses.shape[0]
        self.total_pixels_im = self.H * self.W
        self.total_pixels = self.n_frames * self.total_pixels_im
        self.rgb, self.depth, self.mask, self.normal = f32(rgb), f32(depth), f32(mask), f32(normal)
        for name, t, c in (("rgb", self.rgb, 3), ("depth", self.depth, 1), ("mask", self.mask, 1), ("normal", self.normal, 3)):
            if t is not None and t.numel() != self.total_pixels * c:
                raise ValueError("DevicePixelBank: %s has %d elements, expected %d" % (name, t.numel(), self.total_pixels * c))
        self.device = dev
        self.sampling_idx = None
        self._bad = torch.zeros(1, dtype=torch.int32, device=dev)

    def __len__(self):
        return self.total_pixels if self.sampling_idx is None else int(self.sampling_idx.numel())

    def change_sampling_idx(self, sampling_size, generator=None):
        """scene_dataset.py:468-478 (pixel mode): a fresh random subset of sampling_size / (H*W) of all pixels; -1 = all,
        in order.  `generator`: a CPU torch.Generator reproduces the reference's CPU randperm; None draws on the GPU."""
        if sampling_size == -1:
            self.sampling_idx = None
            return
        total = int(float(sampling_size) / float(self.total_pixels_im) * self.total_pixels)
        if generator is not None:
            perm = torch.randperm(self.total_pixels, generator=generator).to(self.device)
        else:
            perm = torch.randperm(self.total_pixels, device=self.device)
        self.sampling_idx = perm[:total].contiguous()

    def batch(self, ray_ids):
        """(indices, model_input, ground_truth) for the given ray ids, shaped like the reference's collate_fn output:
        indices int64 [n] (frame of every ray); model_input ray_dirs / ray_dirs_tmp / ray_cam_loc [n,3], ray_pose
        [n,4,4]; ground_truth rgb [n,3], depth [n,1], mask [n,1], normal [n,3]."""
        ids = ray_ids.to(self.device, torch.int64).contiguous()
        n, dev = ids.numel(), self.device
        e = lambda *s: torch.empty(*s, device=dev)        # noqa: E731
        inp = {"ray_dirs": e(n, 3), "ray_dirs_tmp": e(n, 3), "ray_cam_loc": e(n, 3), "ray_pose": e(n, 4, 4)}
        gt = {"rgb": e(n, 3) if self.rgb is not None else None, "depth": e(n, 1) if self.depth is not None else None,
              "mask": e(n, 1) if self.mask is not None else None, "normal": e(n, 3) if self.normal is not None else None}
        idx = torch.empty(n, dtype=torch.int64, device=dev)
        self._bad.zero_()
        _lib.call("msdf_pixel_batch", _lib.ptr(ids), n, _lib.ptr(self.poses), _lib.ptr(self.intrinsics), self.n_frames, self.H,
                  self.W, _lib.ptr(self.rgb), _lib.ptr(self.depth), _lib.ptr(self.mask), _lib.ptr(self.normal),
                  _lib.ptr(inp["ray_dirs"]), _lib.ptr(inp["ray_dirs_tmp"]), _lib.ptr(inp["ray_cam_loc"]), _lib.ptr(inp["ray_pose"]),
                  _lib.ptr(idx), _lib.ptr(gt["rgb"]), _lib.ptr(gt["depth"]), _lib.ptr(gt["mask"]), _lib.ptr(gt["normal"]),
                  _lib.ptr(self._bad), _lib.stream())
        return idx, inp, {k: v for k, v in gt.items() if v is not None}

    def check(self):
        """Raises if any batch since the last call named a ray id outside the bank (one host sync; debug / tests)."""
        if int(self._bad.item()) != 0:
            self._bad.zero_()
            raise IndexError("DevicePixelBank: ray id out of range [0, %d)" % self.total_pixels)

    def batches(self, num_pixels):
        """Iterates the epoch like DataLoader(batch_size=num_pixels, shuffle=False) over the sampled ids
        (monosdf_train.py:180-184): consecutive slices of sampling_idx, the last one ragged."""
        n = len(self)
        for lo in range(0, n, num_pixels):
            hi = min(lo + num_pixels, n)
            ids = self.sampling_idx[lo:hi] if self.sampling_idx is not None else torch.arange(lo, hi, device=self.device)
            yield self.batch(ids)
