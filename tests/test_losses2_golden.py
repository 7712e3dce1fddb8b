"""The loss classes round 1 left as shells — DSSIML1, SensorDepthLoss, AdaptiveDepth, AdaptiveNormal,
LocalPearsonDepthLoss — against a golden produced by the reference's own classes
(tests/golden/make_golden_losses.py -> dn_reference_losses2.npz)."""
import os
import types

import numpy as np
import pytest
import torch

from dn_splatter_b200.losses import (AdaptiveDepth, AdaptiveNormal, DepthLoss, DepthLossType, DSSIML1, LocalPearsonDepthLoss,
                                     NormalLoss, NormalLossType, SensorDepthLoss)

TOL = dict(rtol=1e-5, atol=1e-6)


@pytest.fixture(scope="module")
def z(golden_dir):
    d = np.load(os.path.join(golden_dir, "dn_reference_losses2.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def test_dssim_l1_per_pixel(z):
    torch.testing.assert_close(DSSIML1()(z["in_dssim_a3"], z["in_dssim_b3"]), z["out_dssim_pp_3"], **TOL)
    torch.testing.assert_close(DSSIML1(kernel_size=5, alpha=0.6)(z["in_dssim_a1"], z["in_dssim_b1"]), z["out_dssim_pp_1"], **TOL)
    s = DSSIML1(implementation="scalar", kernel_size=11)(z["in_dssim_a3"], z["in_dssim_a3"])
    assert abs(float(s)) < 1e-6  # identical images: SSIM = 1, L1 = 0
    with pytest.raises(NotImplementedError):
        DSSIML1(implementation="scalar", single_resolution=False)


def test_sensor_depth_loss(z):
    rs = types.SimpleNamespace(frustums=types.SimpleNamespace(starts=z["in_sd_starts"]))
    l1, fs, sd = SensorDepthLoss(truncation=0.25)(
        {"sensor_depth": z["in_sd_sensor"]},
        {"depth": z["in_sd_depth_pred"], "ray_samples": rs, "field_outputs": {"sdf": z["in_sd_sdf"]},
         "directions_norm": z["in_sd_dnorm"]})
    torch.testing.assert_close(l1, z["out_sensor_l1"], **TOL)
    torch.testing.assert_close(fs, z["out_sensor_fs"], **TOL)
    torch.testing.assert_close(sd, z["out_sensor_sdf"], **TOL)


def test_adaptive_depth_and_normal(z):
    pd, gd, img, conf = z["in_ad_pd"], z["in_ad_gd"], z["in_ad_img"], z["in_ad_conf"]
    for step in (100, 9000):
        got = DepthLoss(DepthLossType.AdaptiveDepth)(pd, gd, img, gd > 0.1, conf, step)
        torch.testing.assert_close(got, z[f"out_adaptive_depth_{step}"], **TOL)
    for step in (100, 20000):
        got = NormalLoss(NormalLossType.AdaptiveNormal)(z["in_an_pn"], z["in_an_gn"], step)
        torch.testing.assert_close(got, z[f"out_adaptive_normal_{step}"], **TOL)
    assert isinstance(DepthLoss(DepthLossType.AdaptiveDepth).loss, AdaptiveDepth)
    assert isinstance(NormalLoss(NormalLossType.AdaptiveNormal).loss, AdaptiveNormal)


def test_local_pearson(z, monkeypatch):
    draws = [z["in_lp_x0"], z["in_lp_y0"]]
    real = torch.randint

    def replay(*a, **k):  # the reference's own random window corners, in the order it drew them
        want = draws.pop(0)
        assert real(*a, **k).shape == want.shape
        return want

    monkeypatch.setattr(torch, "randint", replay)
    got = DepthLoss(DepthLossType.LocalPearsonDepthLoss)(z["in_lp_pred"], z["in_lp_gt"], 24, 0.5)
    torch.testing.assert_close(got, z["out_local_pearson"], rtol=1e-4, atol=1e-6)
    assert isinstance(DepthLoss(DepthLossType.LocalPearsonDepthLoss).loss, LocalPearsonDepthLoss)
