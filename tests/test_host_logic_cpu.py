"""CPU tests of the host-side mirror of the reference interface: cameras, view matrix convention, configs, synthetic
scenes, API error behaviour (no CPU compute path)."""
import math

import pytest
import torch

from dn_splatter_b200 import get_viewmat
from dn_splatter_b200.cameras import Cameras, is_camera
from dn_splatter_b200.dn_config import METHODS, TRAINER_DEFAULTS, optimizer_groups
from dn_splatter_b200.dn_model import (DNSplatterModelConfig, matrix_to_quaternion, num_sh_bases, quat_to_rotmat,
                                       random_quat_tensor, rotation_between, ssim)
from dn_splatter_b200.synthetic import make_scene, ring_cameras
from oracle import dn_ref


def test_viewmat_matches_oracle_and_is_rigid():
    for cam in ring_cameras(5, 64, 48):
        vm = get_viewmat(cam["c2w"])
        torch.testing.assert_close(vm, dn_ref.get_viewmat(cam["c2w"]), rtol=1e-6, atol=1e-6)
        R = vm[:3, :3]
        torch.testing.assert_close(R @ R.T, torch.eye(3), rtol=1e-5, atol=1e-5)
        # the camera centre maps to the origin, the look-at target (world origin) lands on +z (OpenCV: z forward)
        pos = cam["c2w"][:, 3]
        torch.testing.assert_close(R @ pos + vm[:3, 3], torch.zeros(3), rtol=0, atol=1e-5)
        tgt = vm[:3, 3]
        assert tgt[2] > 0 and abs(float(tgt[0])) < 1e-4 and abs(float(tgt[1])) < 1e-4


def test_cameras_duck_type():
    c = ring_cameras(3, 80, 60)[1]
    cam = Cameras(c["c2w"][None], c["fx"], c["fy"], c["cx"], c["cy"], 80, 60, metadata={"cam_idx": 4})
    assert is_camera(cam) and not is_camera("x") and cam.shape[0] == 1
    K = cam.get_intrinsics_matrices()
    assert K.shape == (1, 3, 3) and float(K[0, 0, 0]) == pytest.approx(0.9 * 80) and float(K[0, 1, 2]) == 30.0
    cam.rescale_output_resolution(0.5)
    assert int(cam.width) == 40 and int(cam.height) == 30 and float(cam.fx) == pytest.approx(0.45 * 80)
    cam.rescale_output_resolution(2.0)
    assert int(cam.width) == 80 and float(cam.cx) == pytest.approx(40.0)


def test_config_surface_matches_reference_names():
    cfg = DNSplatterModelConfig()
    # reference dn_model.py:55-123 fields and defaults that reach the hot path
    assert cfg.regularization_strategy == "dn-splatter" and cfg.predict_normals and cfg.use_normal_loss
    assert cfg.depth_lambda == 0.0 and not cfg.use_depth_loss and cfg.normal_supervision == "mono"
    assert cfg.warmup_length == 500 and cfg.stop_split_at == 15000 and cfg.num_downscales == 0
    for dead in ("use_depth_smooth_loss", "smooth_loss_lambda", "use_normal_cosine_loss", "use_normal_tv_loss", "normal_lambda",
                 "use_sparse_loss", "sparse_lambda", "sparse_loss_steps", "two_d_gaussians", "pearson_lambda",
                 "depth_tolerance", "output_depth_during_training"):
        assert hasattr(cfg, dead)
    assert set(METHODS) == {"dn-splatter", "ags-mesh", "dn-splatter-big"}
    assert METHODS["dn-splatter-big"]["model"]().cull_alpha_thresh == 0.005
    assert METHODS["ags-mesh"]["model"]().regularization_strategy == "ags-mesh"
    g = optimizer_groups()
    assert g["means"]["lr"] == 1.6e-4 and g["means"]["lr_final"] == 1.6e-6 and g["features_rest"]["lr"] == 0.0025 / 20
    assert g["opacities"]["lr"] == 0.05 and g["scales"]["lr"] == 0.005 and g["quats"]["lr"] == 0.001 and "normals" in g
    assert TRAINER_DEFAULTS["max_num_iterations"] == 30000 and TRAINER_DEFAULTS["mixed_precision"] is False


def test_quaternion_helpers():
    q = random_quat_tensor(64, generator=torch.Generator().manual_seed(0))
    torch.testing.assert_close(q.norm(dim=-1), torch.ones(64), rtol=1e-5, atol=1e-5)
    R = quat_to_rotmat(q)
    torch.testing.assert_close(R @ R.transpose(1, 2), torch.eye(3).expand(64, 3, 3), rtol=1e-5, atol=1e-5)
    q2 = matrix_to_quaternion(R)
    same = torch.minimum((q - q2).norm(dim=-1), (q + q2).norm(dim=-1))  # q and -q are the same rotation
    assert float(same.max()) < 1e-4
    z = torch.tensor([0.0, 0.0, 1.0]).repeat(8, 1)
    n = torch.nn.functional.normalize(torch.randn(8, 3, generator=torch.Generator().manual_seed(1)), dim=-1)
    torch.testing.assert_close((rotation_between(z, n) @ z[..., None]).squeeze(-1), n, rtol=1e-4, atol=1e-5)
    assert num_sh_bases(3) == 16 and num_sh_bases(0) == 1


def test_synthetic_scene_is_seeded_and_shaped():
    a, b = make_scene(100, seed=3), make_scene(100, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["features_rest"].shape == (100, 15, 3) and a["opacities"].shape == (100, 1)
    assert float(a["means"].abs().max()) <= 5.0
    cams = ring_cameras(4, 32, 24)
    assert abs(float(cams[0]["c2w"][:, 3].norm()) - 8.0) < 1e-5 and cams[0]["fx"] == pytest.approx(0.9 * 32)


def test_ssim_of_identical_images_is_one():
    x = torch.rand(1, 3, 40, 48, generator=torch.Generator().manual_seed(0))
    assert float(ssim(x, x)) == pytest.approx(1.0, abs=1e-5)
    assert float(ssim(x, torch.rand(1, 3, 40, 48, generator=torch.Generator().manual_seed(1)))) < 0.2


def test_fused_ops_refuse_cpu_tensors():
    from dn_splatter_b200._lib import DnrError
    from dn_splatter_b200.regularization_strategy import DNRegularization, FusedL1
    from dn_splatter_b200.utils.normal_utils import normal_from_depth_image

    with pytest.raises(DnrError):
        FusedL1.apply(torch.rand(4, 4, 3), torch.rand(4, 4, 3))
    with pytest.raises(DnrError):
        DNRegularization()(pred_depth=torch.rand(4, 4, 1), gt_depth=torch.rand(4, 4, 1) + 1, pred_normal=torch.rand(4, 4, 3),
                           gt_normal=torch.rand(4, 4, 3), scales=torch.zeros(3, 3), gt_img=torch.rand(4, 4, 3))
    with pytest.raises(DnrError):
        normal_from_depth_image(torch.rand(4, 4, 1), 3.0, 3.0, 2.0, 2.0, (4, 4), torch.eye(4), torch.device("cpu"))
