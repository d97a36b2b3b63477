"""The work constants bench.py reports against (monosdf_b200/roofline.py) are derived from the layer dimensions and must
equal the figures SURVEY.md section 8(d) pins: A = 459 008, B = 524 544, C = 459 008, D = 140 288 MAC/point (MLP conf),
0.800 ... 1.270 GFLOP per training ray for 1 ... 5 sampler rounds; grid conf A = C = 83 968, B = 149 504."""
import pytest

from monosdf_b200 import roofline


def test_mlp_conf_macs_per_point():
    assert roofline.WORK_MLP == {"A": 459008, "B": 524544, "C": 459008, "D": 140288}


def test_grid_conf_macs_per_point():
    w = roofline.WORK_GRID
    assert (w["A"], w["B"], w["C"], w["D"]) == (83968, 149504, 83968, 140288)


def test_gflop_per_training_ray():
    for k, ref in {1: 0.800, 2: 0.918, 3: 1.035, 4: 1.153, 5: 1.270}.items():
        assert roofline.GFLOP_PER_RAY_MLP[k] == pytest.approx(ref, abs=6e-4)
    for k, ref in {1: 0.245, 3: 0.288, 5: 0.331}.items():
        assert roofline.GFLOP_PER_RAY_GRID[k] == pytest.approx(ref, abs=6e-4)
    # eval render: 0.338 + 0.1175 (k - 1) GFLOP per ray
    assert roofline.gflop_per_ray(roofline.WORK_MLP, 1, train=False) == pytest.approx(0.338, abs=1e-3)
    assert roofline.gflop_per_ray(roofline.WORK_MLP, 3, train=False) == pytest.approx(0.338 + 2 * 0.1175, abs=1e-3)


def test_hash_bytes_per_point():
    assert roofline.HASH_BYTES == {"forward": 1164, "backward": 2188}


def test_layer_dims_follow_the_skip_rule():
    # network.py:44-45: the layer before the skip has dims[l+1] - dims[0] outputs
    layers = roofline.sdf_layer_dims(39, [256] * 8, 257, (4,))
    assert layers[3] == (256, 217) and layers[4] == (256, 256) and layers[-1] == (256, 257)
