"""Host logic of dn_splatter_b200.sugar (the mirror of the reference's SuGaR-style queries) on the CPU: the three kernel
calls are served by the oracle (TEST INFRASTRUCTURE, like tests/cpu_proxy.py) and the result is compared with the
golden produced by the reference's own compute_level_surface_points / get_density / get_sdf / get_density_grad."""
import os
import random
from contextlib import contextmanager

import numpy as np
import pytest
import torch

from oracle import sugar_ref as S
from tests.cpu_proxy import cpu_proxy

PARAMS = ("means", "quats", "scales", "opacities", "features_dc", "features_rest")


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dn_sugar_a.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@contextmanager
def sugar_proxy():
    import dn_splatter_b200.sugar as SG

    saved = (SG.KnnIndex, SG._density_call, SG.ray_densities, SG._need_cuda)

    def params_of(model):
        return {k: model.gauss_params[k].detach() for k in ("means", "scales", "quats", "opacities")}

    class Index:
        def __init__(self, points):
            self.points = points

        def query(self, q, k, skip_first=True, return_distances=False):
            assert skip_first and not return_distances
            return S.knn_sk(self.points, q, k)

    def density_call(samples, idx, model, per_row, clamp_min):
        assert per_row == 1
        return S.raw_density(samples, idx, params_of(model)).clamp(min=clamp_min)

    def ray_densities(model, points, idx, cam_pos, n_range=21, range_size=3.0):
        return S.ray_densities(points, idx, params_of(model), cam_pos, n_range, range_size)

    SG.KnnIndex, SG._density_call, SG.ray_densities, SG._need_cuda = Index, density_call, ray_densities, lambda *a: None
    try:
        yield SG
    finally:
        SG.KnnIndex, SG._density_call, SG.ray_densities, SG._need_cuda = saved


def _model(gold):
    from dn_splatter_b200.cameras import Cameras
    from dn_splatter_b200.dn_model import DNSplatterModelConfig

    m = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black", ssim_lambda=0.0).setup(device="cpu")
    m.load_gaussians({k: gold["in_" + k] for k in PARAMS})
    m.background_color = torch.zeros(3)
    m.step = 30000
    m.eval()
    fx, fy, cx, cy, W, H = [float(v) for v in gold["cam_intr"]]
    return m, Cameras(gold["cam_c2w"][None], fx, fy, cx, cy, int(W), int(H))


@pytest.mark.parametrize("mode", ["closest_gaussian", "analytical"])
def test_compute_level_surface_points_matches_reference(gold, mode):
    with cpu_proxy(), sugar_proxy() as SG:
        m, cam = _model(gold)
        random.seed(11)
        res = SG.compute_level_surface_points(m, cam, num_samples=10_000, return_normal=mode)
        for level in (0.1, 0.3, 0.5):
            for k in ("points", "normals", "colors"):
                want = gold[f"level_{mode}_{level}_{k}"]
                assert res[level][k].shape == want.shape
                atol = 2e-4 if (k == "normals" and mode == "analytical") else 2e-5
                torch.testing.assert_close(res[level][k], want, rtol=1e-4, atol=atol)


def test_point_queries_match_reference(gold):
    with cpu_proxy(), sugar_proxy() as SG:
        m, _ = _model(gold)
        q, idx = gold["q_samples"], gold["q_idx"]
        assert torch.equal(SG.get_closest_gaussians(m, q), idx)
        torch.testing.assert_close(SG.get_density(m, q), gold["q_density"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(SG.get_sdf(m, q, idx), gold["q_sdf"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(SG.get_density_grad(m, q, closest_gaussians=idx), gold["q_density_grad"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(SG.get_sdf_weight(m, idx), gold["q_sdf_weight"], rtol=1e-6, atol=1e-8)


def test_no_cpu_path():
    import dn_splatter_b200.sugar as SG

    with pytest.raises(Exception, match="no CPU path"):
        SG.knn_gpu(torch.zeros(4, 3), torch.zeros(2, 3), 2)


def test_sampling_ideal_sdf_and_loss_weights_match_reference(gold):
    """sample_points_in_gaussians / get_ideal_sdf / get_sdf_loss_weight are device-agnostic torch: run as they are."""
    import dn_splatter_b200.sugar as SG

    with cpu_proxy():
        m, cam = _model(gold)
        out = m.get_outputs(cam)  # sets m.camera (used by the "std" weights) and m.radii
        vis = torch.where(m.radii > 0)[0][::2]
        assert torch.equal(vis, gold["samp_vis_indices"])
        for tag, vi in (("all", None), ("vis", vis)):
            torch.manual_seed(99)
            pts, ids = SG.sample_points_in_gaussians(m, 200, vis_indices=vi)
            assert torch.equal(ids, gold[f"samp_{tag}_ids"])
            torch.testing.assert_close(pts.detach(), gold[f"samp_{tag}_points"], rtol=1e-5, atol=1e-6)
        depth = gold["ideal_depth_map"]
        torch.testing.assert_close(out["depth"], depth, rtol=1e-5, atol=1e-6)
        for tag, mk in (("nomask", None), ("mask", gold["ideal_mask"])):
            sdf, valid = SG.get_ideal_sdf(m, gold["samp_all_points"], depth, cam, mask=mk)
            assert torch.equal(valid, gold[f"ideal_{tag}_valid"])
            torch.testing.assert_close(sdf, gold[f"ideal_{tag}_sdf"], rtol=1e-5, atol=1e-6)
        for mode in ("area", "std"):
            torch.testing.assert_close(SG.get_sdf_loss_weight(m, gold["samp_all_ids"], mode=mode), gold[f"lossw_{mode}"],
                                       rtol=1e-5, atol=1e-7)
