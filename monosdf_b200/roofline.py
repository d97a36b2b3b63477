"""Algorithmic work per unit of the hot path, derived from the network dimensions.

MACs per point:  A = SDF-only forward (sdf column of the last layer only), B = full forward (sdf + features),
C = reverse sweep for grad_x sdf (the reference: autograd through the same layers), D = colour-net forward.
FLOP = 2 MAC.  A training step costs 3x the forward for everything that is differentiated (forward + two backward
GEMMs per layer), so per ray with S render samples, 4 eikonal points and k sampler rounds of n_up points:

    FLOP / ray = S * 6 (B + C + D) + 4 * 6 (A + C) + n_up * k * 2 A

bench.py reports `step_algorithmic_tflops` from these; tests/test_roofline_constants.py pins them to SURVEY's figures.
"""


def sdf_layer_dims(d_in_enc, hidden, d_out, skip_in):
    """(in, out) of every Linear of ImplicitNetwork / ImplicitNetworkGrid (network.py:37-49, 205-216)."""
    dims = [d_in_enc] + list(hidden) + [d_out]
    layers = []
    for l in range(len(dims) - 1):
        out = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        layers.append((dims[l], out))
    return layers


def mlp_work(d_in_enc, hidden, feature_size, skip_in, color_in, color_hidden, color_out=3):
    """MACs per point {A, B, C, D} for an SDF net with `hidden` widths and a colour net color_in -> color_hidden -> 3."""
    layers = sdf_layer_dims(d_in_enc, hidden, 1 + feature_size, skip_in)
    body = sum(i * o for i, o in layers[:-1])
    last_in = layers[-1][0]
    A = body + last_in * 1
    B = body + last_in * (1 + feature_size)
    C = A                                   # the reverse sweep multiplies by the same matrices (sdf row of the last layer)
    cd = [color_in] + list(color_hidden) + [color_out]
    D = sum(cd[l] * cd[l + 1] for l in range(len(cd) - 1))
    return {"A": A, "B": B, "C": C, "D": D}


def gflop_per_ray(work, rounds, n_render=98, n_eik=4, n_up=128, train=True):
    A, B, C, D = work["A"], work["B"], work["C"], work["D"]
    if train:
        flop = n_render * 6 * (B + C + D) + n_eik * 6 * (A + C) + n_up * rounds * 2 * A
    else:
        flop = n_render * 2 * (B + C + D) + n_up * rounds * 2 * A
    return flop / 1e9


# scannet-MLP conf: PE 6 -> 39 inputs, 8 x 256, skip at 4, features 256; colour 289 -> 256 -> 256 -> 3 (a7, a10)
WORK_MLP = mlp_work(39, [256] * 8, 256, (4,), 289, [256, 256])
# kitchen-grids conf: 39 + 32 hash features, 2 x 256, no effective skip; same colour net (a8)
WORK_GRID = mlp_work(71, [256, 256], 256, (), 289, [256, 256])
GFLOP_PER_RAY_MLP = {k: gflop_per_ray(WORK_MLP, k) for k in range(1, 6)}
GFLOP_PER_RAY_GRID = {k: gflop_per_ray(WORK_GRID, k) for k in range(1, 6)}

# hash grid, bytes per point at element granularity (L = 16 levels, C = 2 features, D = 3): SURVEY 8d
HASH_BYTES = {"forward": 12 + 16 * 8 * 2 * 4 + 16 * 2 * 4, "backward": 12 + 16 * 2 * 4 + 2 * 16 * 8 * 2 * 4}
