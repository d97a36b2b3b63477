"""CPU: the oracle (oracle/port.py, oracle/sampler_oracle.c) against the golden vectors produced by the unmodified
reference (tests/golden/*.pt, oracle/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

from oracle import port
from oracle.sampler_oracle import OracleSampler
from tests.helpers import build_model, frac_within, oracle_forward, params_of, rel_err

CASES = ["mlp_small", "gridmlp_small", "mlp_full", "spec_small", "spec_full"]
RAYS = lambda fx: port.synthetic_rays(fx["n_rays"], seed=1)   # noqa: E731


@pytest.mark.parametrize("case", CASES)
def test_constructor_reproduces_reference_weights(golden, case):
    """Same seed -> the drop-in's constructor yields the reference's state_dict (keys, shapes and values)."""
    fx = golden(case)
    sd = build_model(fx).state_dict()
    assert set(sd.keys()) == set(fx["state_checksum"].keys())
    for k, v in sd.items():
        assert float(v.double().sum()) == pytest.approx(float(fx["state_checksum"][k]), rel=1e-12, abs=1e-12), k
    if "state_dict" in fx:
        for k, v in fx["state_dict"].items():
            assert torch.equal(sd[k], v), k


@pytest.mark.parametrize("case", CASES)
def test_port_eval_matches_reference(golden, case):
    fx = golden(case)
    params = params_of(build_model(fx))
    out, _ = oracle_forward(fx, params, RAYS(fx), training=False)
    for k in ["z_vals", "sdf", "rgb", "depth_vals"]:
        assert rel_err(out[k], fx["eval"][k]) < 2e-5, k
    # composited quantities amplify fp32 rounding of the sdf by 1/beta (= 50..100) before the exp
    for k in ["weights", "rgb_values", "depth_values", "normal_map"]:
        assert rel_err(out[k], fx["eval"][k]) < 5e-4, k
    if case.startswith("spec"):      # diffuse/specular split (network.py:427-454, 576-582)
        assert rel_err(out["rgb_spec"], fx["eval"]["rgb_spec"]) < 2e-5
        assert rel_err(out["rgb_spec_values"], fx["eval"]["rgb_spec_values"]) < 3e-4


@pytest.mark.parametrize("case", ["mlp_small", "mlp_full"])
def test_port_uv_path_matches_reference(golden, case):
    fx = golden(case)
    params = params_of(build_model(fx))
    out, _ = oracle_forward(fx, params, fx["uv_input"], training=False, uv=True)
    assert rel_err(out["z_vals"], fx["uv_eval"]["z_vals"]) < 2e-5
    for k in ["rgb_values", "depth_values", "normal_map"]:
        assert rel_err(out[k], fx["uv_eval"][k]) < 3e-4, k


@pytest.mark.parametrize("case", ["mlp_small", "gridmlp_small", "spec_small"])
def test_port_train_step_matches_reference(golden, case):
    """Forward + MonoSDFLoss + backward with the CPU generator seeded like make_golden: loss and every gradient."""
    fx = golden(case)
    params = params_of(build_model(fx), requires_grad=True)
    out, _ = oracle_forward(fx, params, RAYS(fx), training=True, seed=fx["train_seed"])
    for k in ["z_vals", "grad_theta", "grad_theta_nei"]:
        assert rel_err(out[k], fx["train"][k]) < 2e-5, k
    assert rel_err(out["rgb_values"], fx["train"]["rgb_values"]) < 3e-4
    gt = port.synthetic_gt(fx["n_rays"], seed=2)
    loss = port.monosdf_loss(out, gt)
    if "rgb_spec_values" in out:     # make_golden.py gives the specular output a gradient of its own
        loss["loss"] = loss["loss"] + 0.25 * (out["rgb_spec_values"] * gt["rgb"].reshape(-1, 3)).mean()
    assert float(loss["loss"]) == pytest.approx(float(fx["train_loss"]["loss"]), rel=1e-5)
    loss["loss"].backward()
    for k, g in fx["train_grad"].items():
        assert params[k].grad is not None, k
        assert rel_err(params[k].grad, g) < 1e-3, k


@pytest.mark.parametrize("case", CASES)
def test_c_sampler_oracle_matches_reference(golden, case):
    """The bit-exact sampler restatement (own expf/expm1f, fixed scan order) lands on the reference's samples."""
    fx = golden(case)
    params = params_of(build_model(fx))
    cfg = port.cfg_from_conf(fx["conf"])
    rays = RAYS(fx)
    sc = cfg.sampler
    smp = OracleSampler(cfg.scene_bounding_sphere, sc.near, sc.N_samples, sc.N_samples_eval, sc.N_samples_extra, sc.eps,
                        sc.beta_iters, sc.max_total_iters, sc.add_tiny)
    beta0 = float(port.get_beta(params, cfg))
    trace = {}
    with torch.no_grad():
        z, _ = smp.get_z_vals(rays["ray_dirs"], rays["ray_cam_loc"], lambda p: port.sdf_vals(params, cfg, p), beta0, False, trace)
    assert z.shape == fx["eval"]["z_vals"].shape
    assert frac_within(z, fx["eval"]["z_vals"], 1e-4) > 0.995
