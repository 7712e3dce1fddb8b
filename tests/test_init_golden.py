"""DNSplatterModel.populate_modules and the quaternion helpers against goldens produced by the REFERENCE's own
populate_modules / rotate_vector_to_vector / matrix_to_quaternion (tests/golden/make_golden_init.py)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("means", "scales", "quats", "features_dc", "features_rest", "opacities", "normals")


def _load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, name)).items()}


def _init(tag):
    from dn_splatter_b200.dn_model import DNSplatterModelConfig

    z = _load(f"dn_init_{tag}.npz")
    seeds = tuple(z["seed_" + k] for k in ("points", "rgb", "normals") if "seed_" + k in z)
    torch.manual_seed(4321)
    m = DNSplatterModelConfig(use_depth_loss=True, depth_lambda=0.2).setup(seed_points=seeds, num_train_data=10, device="cpu")
    return m, z


def test_init_from_seed_points_with_normals():
    m, z = _init("normals")
    for k in NAMES:
        torch.testing.assert_close(m.gauss_params[k].detach(), z["out_" + k], rtol=1e-5, atol=1e-6, msg=lambda s: f"{k}: {s}")
    torch.testing.assert_close(m.background_color, z["background_color"])
    assert abs(float(m.regularization_strategy.depth_lambda) - float(z["depth_lambda"])) < 1e-12


def test_init_from_seed_points_without_normals_uses_the_same_random_stream():
    m, z = _init("plain")
    for k in NAMES:
        torch.testing.assert_close(m.gauss_params[k].detach(), z["out_" + k], rtol=1e-5, atol=1e-6, msg=lambda s: f"{k}: {s}")


def test_rotation_helpers():
    from dn_splatter_b200.dn_model import matrix_to_quaternion, rotation_between

    z = _load("dn_init_helpers.npz")
    mat = rotation_between(z["v1"], z["v2"])
    torch.testing.assert_close(mat, z["mat"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(matrix_to_quaternion(z["mat"]), z["quat"], rtol=1e-5, atol=1e-6)
