"""oracle/sugar_ref.py against the golden produced by the REFERENCE's own get_density / get_sdf / get_density_grad /
get_sdf_weight / compute_level_surface_points / knn_sk (tests/golden/make_golden_sugar.py)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import dn_ref, sugar_ref as S

PARAMS = ("means", "quats", "scales", "opacities", "features_dc", "features_rest")


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dn_sugar_a.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _params(z):
    return {k: z["in_" + k] for k in PARAMS}


def test_knn_drops_the_nearest_neighbour_like_the_reference(gold):
    idx = S.knn_sk(gold["in_means"], gold["q_samples"], 16)
    assert torch.equal(idx, gold["q_idx"])
    d = (gold["q_samples"][:, None, :] - gold["in_means"][None]).norm(dim=-1)
    order = d.argsort(dim=1)
    assert torch.equal(idx, order[:, 1:17])  # ranks 2..17: the nearest one is discarded (knn.py:43)


def test_density_sdf_gradient_and_weight(gold):
    p = _params(gold)
    torch.testing.assert_close(S.get_density(gold["q_samples"], p, gold["q_idx"]), gold["q_density"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(S.get_sdf(gold["q_samples"], p, gold["q_idx"]), gold["q_sdf"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(S.get_density_grad(gold["q_samples"], p, gold["q_idx"]), gold["q_density_grad"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(S.get_sdf_weight(gold["q_idx"], p), gold["q_sdf_weight"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("mode", ["closest_gaussian", "analytical"])
def test_level_surface_points(gold, mode):
    p = _params(gold)
    fx, fy, cx, cy, W, H = [float(v) for v in gold["cam_intr"]]
    W, H = int(W), int(H)
    out = dn_ref.get_outputs(p, gold["cam_c2w"], fx, fy, cx, cy, W, H, torch.zeros(3))
    random.seed(11)
    res = S.compute_level_surface_points(p, out["gauss_normals"], out["depth"].detach(), out["rgb"].detach(), gold["cam_c2w"],
                                         fx, fy, cx, cy, W, H, num_samples=10_000, return_normal=mode)
    for level in (0.1, 0.3, 0.5):
        for k in ("points", "normals", "colors"):
            want = gold[f"level_{mode}_{level}_{k}"]
            assert res[level][k].shape == want.shape, (level, k, res[level][k].shape, want.shape)
            # analytical normals normalise a sum of 16 terms that partly cancel: allow a few 1e-5 of re-association noise
            atol = 2e-4 if (k == "normals" and mode == "analytical") else 2e-5
            torch.testing.assert_close(res[level][k], want, rtol=1e-4, atol=atol, msg=lambda s: f"{mode} {level} {k}: {s}")


def test_level_surface_subsampling_uses_pythons_random(gold):
    p = _params(gold)
    fx, fy, cx, cy, W, H = [float(v) for v in gold["cam_intr"]]
    out = dn_ref.get_outputs(p, gold["cam_c2w"], fx, fy, cx, cy, int(W), int(H), torch.zeros(3))
    random.seed(12)
    res = S.compute_level_surface_points(p, out["gauss_normals"], out["depth"].detach(), out["rgb"].detach(), gold["cam_c2w"],
                                         fx, fy, cx, cy, int(W), int(H), num_samples=25, surface_levels=(0.3,))
    torch.testing.assert_close(res[0.3]["points"], gold["level_sub25_points"], rtol=1e-4, atol=2e-5)
