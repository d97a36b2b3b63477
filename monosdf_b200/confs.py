"""Model configurations of the benchmark workloads (shapes from the reference's conf files) and a dict-backed
stand-in for pyhocon's ConfigTree (`get_int/get_float/...`), so the package has no pyhocon dependency."""
import copy


class Conf(dict):
    def _get(self, key, default=None):
        if key in self:
            return self[key]
        if default is not None:
            return default
        raise KeyError(key)

    def get_int(self, k, default=None):
        return int(self._get(k, default))

    def get_float(self, k, default=None):
        return float(self._get(k, default))

    def get_bool(self, k, default=None):
        return bool(self[k]) if k in self else bool(default)

    def get_string(self, k, default=None):
        return str(self._get(k, default))

    def get_list(self, k, default=None):
        return list(self._get(k, default))

    def get_config(self, k, default=None):
        v = self._get(k, default)
        return v if isinstance(v, Conf) else to_conf(v)


def to_conf(d):
    out = Conf()
    for k, v in d.items():
        out[k] = to_conf(v) if isinstance(v, dict) else v
    return out


# code/confs/mp_jh4fc5c5qoQ_undist_scannetMLP.conf-shaped, upstream-style plain MLP (BASELINE.json configs[0..1])
SCANNET_MLP = {
    "feature_vector_size": 256,
    "scene_bounding_sphere": 1.1,
    "Grid_MLP": False,
    "implicit_network": {"d_in": 3, "d_out": 1, "dims": [256] * 8, "geometric_init": True, "bias": 0.9, "skip_in": [4],
                         "weight_norm": True, "multires": 6, "inside_outside": True},
    "rendering_network": {"mode": "idr", "d_in": 9, "d_out": 3, "dims": [256, 256], "weight_norm": True,
                          "multires_view": 4, "per_image_code": False},
    "density": {"params_init": {"beta": 0.1}, "beta_min": 0.0001},
    "ray_sampler": {"near": 0.0, "N_samples": 64, "N_samples_eval": 128, "N_samples_extra": 32, "eps": 0.1,
                    "beta_iters": 10, "max_total_iters": 5},
}

# code/confs/mi.conf (expname kitchen_HDR_grids): 16 x 2 hash grid, 2^19 entries/level, 16 -> 2048, 2 x 256 MLP
KITCHEN_GRIDS = copy.deepcopy(SCANNET_MLP)
KITCHEN_GRIDS["Grid_MLP"] = True
KITCHEN_GRIDS["implicit_network"].update(dims=[256, 256], skip_in=[4], use_grid_feature=True, divide_factor=1.1,
                                         base_size=16, end_size=2048, logmap=19, num_levels=16, level_dim=2)
KITCHEN_GRIDS["rendering_network"].update(per_image_code=True)
