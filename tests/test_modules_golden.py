"""The torch-executed API surface (losses.py modules, find_edges, AGSMeshRegularization) against goldens produced by the
reference's own classes (tests/golden/make_golden.py -> dn_reference_modules.npz)."""
import os

import numpy as np
import pytest
import torch

from dn_splatter_b200.losses import (DepthLoss, DepthLossType, EdgeAwareLogL1, EdgeAwareTV, HuberL1, L1, LogL1, NormalLoss,
                                     NormalLossType, PearsonDepthLoss, TVLoss)
from dn_splatter_b200.regularization_strategy import AGSMeshRegularization, find_edges, mean_angular_error

TOL = dict(rtol=1e-5, atol=1e-6)


@pytest.fixture(scope="module")
def z(golden_dir):
    d = np.load(os.path.join(golden_dir, "dn_reference_modules.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def test_loss_modules_match_reference(z):
    pd, gd, img = z["in_pd"], z["in_gd"], z["in_img"]
    mask = gd > 0.1
    torch.testing.assert_close(EdgeAwareLogL1(implementation="per-pixel")(pd, gd, img, mask), z["out_edge_aware_logl1_pp"], **TOL)
    torch.testing.assert_close(LogL1(implementation="per-pixel")(pd, gd), z["out_logl1_pp"], **TOL)
    torch.testing.assert_close(L1(implementation="per-pixel")(pd, gd), z["out_l1_pp"], **TOL)
    torch.testing.assert_close(HuberL1()(pd, gd), z["out_huber"], **TOL)
    torch.testing.assert_close(EdgeAwareTV()(pd[None], img[None]), z["out_edge_aware_tv"], **TOL)
    torch.testing.assert_close(PearsonDepthLoss()(pd, gd + 0.01), z["out_pearson"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(DepthLoss(DepthLossType.MSE)(pd, gd), z["out_mse"], **TOL)
    assert isinstance(NormalLoss(NormalLossType.Smooth).loss, TVLoss) and isinstance(NormalLoss(NormalLossType.L1).loss, L1)
    with pytest.raises(ValueError):
        DepthLoss("nope")


def test_find_edges_and_angular_error_match_reference(z):
    assert torch.equal(find_edges(z["in_gn"]).float(), z["out_find_edges_3"])
    assert torch.equal(find_edges(z["in_pd"].permute(2, 0, 1)).float(), z["out_find_edges_1"])
    torch.testing.assert_close(mean_angular_error(z["in_sn"], z["in_gn"]), z["out_mean_angular_error"], **TOL)


def test_ags_mesh_strategy_matches_reference(z):
    ags = AGSMeshRegularization()
    for step in (100, 8000, 16000):
        got = torch.as_tensor(ags.get_normal_loss(step, z["in_sn"], z["in_gn"], z["in_pn"])).float()
        torch.testing.assert_close(got, z[f"out_ags_normal_{step}"], **TOL)
    d = ags.get_depth_loss(step=100, pred_depth=z["in_pd"], gt_depth=z["in_gd"], confidence_map=z["in_conf"], gt_img=z["in_img"])
    torch.testing.assert_close(d, z["out_ags_depth_100"], **TOL)
    tot = ags(step=100, pred_depth=z["in_pd"], gt_depth=z["in_gd"], surf_normal=z["in_sn"], gt_normal=z["in_gn"],
              pred_normal=z["in_pn"], confidence_map=z["in_conf"], scales=z["in_scales"], gt_img=z["in_img"])
    torch.testing.assert_close(tot, z["out_ags_total_100"], **TOL)
