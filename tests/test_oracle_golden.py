"""Pins oracle/dn_ref.py against golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py imports /root/reference/dn_splatter/{losses,regularization_strategy,
utils/normal_utils}.py unmodified).  fp32 tolerance 1e-6 abs / 1e-5 rel (same torch ops, same order up
to reduction order)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dn_ref

TOL = dict(rtol=1e-5, atol=1e-6)


def _cases(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "dn_reference_[ab].npz")))
    assert files, "golden fixtures missing"
    for f in files:
        z = np.load(f)
        yield os.path.basename(f), {k: torch.from_numpy(z[k]) for k in z.files}


def test_normal_from_depth_matches_reference(golden_dir):
    for name, z in _cases(golden_dir):
        fx, fy, cx, cy = [float(v) for v in z["in_intr"]]
        H, W, _ = z["in_pred_depth"].shape
        n = dn_ref.normal_from_depth_image(z["in_pred_depth"], fx, fy, cx, cy, W, H)
        torch.testing.assert_close(n, z["out_normal_from_depth"], **TOL)
        sn = dn_ref.surface_normal_output(z["in_pred_depth"], fx, fy, cx, cy, W, H)
        torch.testing.assert_close(sn, z["out_surface_normal_output"], **TOL)
        # border pixels are exactly 0.5 after the remap (quirk B5)
        assert torch.all(sn[0] == 0.5) and torch.all(sn[:, 0] == 0.5)


def test_losses_match_reference(golden_dir):
    for name, z in _cases(golden_dir):
        mask = z["in_gt_depth"] > 0.1
        torch.testing.assert_close(
            dn_ref.edge_aware_logl1(z["in_pred_depth"], z["in_gt_depth"], z["in_gt_img"], mask),
            z["out_edge_aware_logl1"], **TOL)
        torch.testing.assert_close(dn_ref.logl1_pp(z["in_pred_depth"][mask], z["in_gt_depth"][mask]).mean(),
                                   z["out_logl1"], **TOL)
        torch.testing.assert_close(dn_ref.l1_loss(z["in_pred_normal"], z["in_gt_normal"]), z["out_l1"], **TOL)
        torch.testing.assert_close(dn_ref.tv_loss(z["in_pred_normal"]), z["out_tv"], **TOL)


@pytest.mark.parametrize("key,kw", [
    ("dn_reg_lambda0.2", dict(depth_lambda=0.2)),
    ("dn_reg_lambda0.5", dict(depth_lambda=0.5)),
    ("dn_reg_LogL1", dict(depth_loss_type="LogL1")),
    ("dn_reg_L1", dict(depth_loss_type="L1")),
    ("dn_reg_mse", dict(depth_loss_type="MSE")),
    ("dn_reg_nodepth", dict(depth_loss_type=None)),
])
def test_dn_regularization_matches_reference(golden_dir, key, kw):
    for name, z in _cases(golden_dir):
        v = dn_ref.dn_regularization(z["in_pred_depth"], z["in_gt_depth"], z["in_pred_normal"], z["in_gt_normal"],
                                     z["in_scales"], z["in_gt_img"], **kw)
        torch.testing.assert_close(v, z["out_" + key], **TOL)


def test_dn_regularization_gradients_match_reference(golden_dir):
    for name, z in _cases(golden_dir):
        pd = z["in_pred_depth"].clone().requires_grad_(True)
        pn = z["in_pred_normal"].clone().requires_grad_(True)
        sc = z["in_scales"].clone().requires_grad_(True)
        dn_ref.dn_regularization(pd, z["in_gt_depth"], pn, z["in_gt_normal"], sc, z["in_gt_img"]).backward()
        torch.testing.assert_close(pd.grad, z["out_grad_pred_depth"], **TOL)
        torch.testing.assert_close(pn.grad, z["out_grad_pred_normal"], **TOL)
        torch.testing.assert_close(sc.grad, z["out_grad_scales"], **TOL)
