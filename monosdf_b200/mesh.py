"""Device-side front end of mesh extraction ("next" row f4 of SURVEY.md section 8).

`sdf_volume_pyramid` produces, for one crop, the SDF volume that utils/plots.py:131-194 (get_surface_sliding) hands to
skimage's marching cubes: coarsest pyramid level evaluated densely, every finer level only where the parent cell had
|sdf| < threshold, unevaluated cells inheriting the parent's value.  Points are generated and compacted on the GPU
(csrc/sdfgrid.cu), the SDF network runs once per level on the compacted list (one launch sequence instead of one
`sdf(pnts).cpu()` per 100 000 points), and the volume leaves the device once.  Marching cubes itself (skimage / trimesh,
plots.py:196-221) is CPU post-processing outside the hot path.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def sdf_volume_pyramid(sdf, lo, hi, crop_n, levels=4, device="cuda", stats=None):
    """sdf: callable [M,3] cuda float tensor -> [M] or [M,1] SDF values (e.g. model.implicit_network.get_sdf_vals).
    lo, hi: 3 floats each (the crop's box); crop_n: samples per axis, divisible by 2**(levels-1).
    Returns the [crop_n, crop_n, crop_n] volume on the device (index order x, y, z like torch.meshgrid 'ij')."""
    dev = torch.device(device)
    lo3 = (ctypes.c_double * 3)(*[float(v) for v in lo])
    hi3 = (ctypes.c_double * 3)(*[float(v) for v in hi])
    # plots.py:162: threshold = 2 * (x_max - x_min) / cropN * 8, halved after every level (:190)
    threshold = 2.0 * (float(hi[0]) - float(lo[0])) / crop_n * 8.0
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    parent, parent_mask = None, None
    for pid in range(levels):
        s = levels - 1 - pid
        n = crop_n >> s
        cells = n * n * n
        slot = torch.empty(cells, dtype=torch.int32, device=dev)
        points = torch.empty(cells, 3, device=dev)
        counter.zero_()
        _lib.call("msdf_sdfgrid_level_points", lo3, hi3, crop_n, s, _lib.ptr(parent_mask), _lib.ptr(slot), _lib.ptr(points),
                  _lib.ptr(counter), _lib.stream())
        cnt = int(counter.item())                     # one 4-byte read-back per level sizes the network call
        if stats is not None:
            stats.append((n, cnt))
        values = None
        if cnt > 0:
            with torch.no_grad():
                values = sdf(points[:cnt]).reshape(-1).float().contiguous()
        level = torch.empty(cells, device=dev)
        mask = torch.empty(cells, dtype=torch.uint8, device=dev) if pid < levels - 1 else None
        _lib.call("msdf_sdfgrid_level_assemble", n, _lib.ptr(slot), _lib.ptr(values), _lib.ptr(parent), float(np.float32(threshold)),
                  _lib.ptr(level), _lib.ptr(mask), _lib.stream())
        parent, parent_mask = level, mask
        threshold /= 2.0
    return parent.reshape(crop_n, crop_n, crop_n)


def surface_volumes(sdf, resolution=512, grid_boundary=(-2.0, 2.0), device="cuda"):
    """The crops of get_surface_sliding (plots.py:110-128): yields (origin [3], spacing [3], volume as a numpy
    [cropN,cropN,cropN] float32 array) per crop -- what `measure.marching_cubes(volume, level, spacing)` consumes (:199-205)."""
    crop_n = 128 if resolution < 512 else 512
    assert resolution % crop_n == 0, "resolution: %d, cropN: %d" % (resolution, crop_n)
    N = resolution // crop_n
    xs = np.linspace(grid_boundary[0], grid_boundary[1], N + 1)
    for i in range(N):
        for j in range(N):
            for k in range(N):
                lo = (xs[i], xs[j], xs[k])
                hi = (xs[i + 1], xs[j + 1], xs[k + 1])
                vol = sdf_volume_pyramid(sdf, lo, hi, crop_n, device=device)
                spacing = tuple((hi[a] - lo[a]) / (crop_n - 1) for a in range(3))
                yield np.array(lo), spacing, vol.cpu().numpy()      # the one device -> host copy of the crop
