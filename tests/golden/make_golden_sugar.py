"""Generates tests/golden/dn_sugar_*.npz by executing the REFERENCE's own SuGaR-style queries
(/root/reference/dn_splatter/dn_model.py: get_density :1077-1135, get_sdf :1137-1158, get_density_grad :1449-1494,
compute_level_surface_points :1207-1447, and utils/knn.py: knn_sk), unmodified, on a small scene.  The depth map the
level-set search starts from comes from the reference's own get_outputs with gsplat served by oracle/gsplat_ref.py
(see make_golden_model.py); sklearn (the reference's KNN backend) is present in this container and used as is.
tests/test_sugar_golden.py checks oracle/sugar_ref.py against the files; the CUDA kernels are checked against
oracle/sugar_ref.py on the GPU.

Run only where /root/reference exists:   python tests/golden/make_golden_sugar.py
"""
import os
import random
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import make_golden_model as G0  # noqa: E402


def surface_scene(n, seed):
    """Flat-ish Gaussians on a wavy sheet in front of the cameras: the level sets of their density form a surface."""
    g = torch.Generator().manual_seed(seed)
    xy = (torch.rand(n, 2, generator=g) - 0.5) * 3.0
    z = 0.15 * torch.sin(2.0 * xy[:, :1]) * torch.cos(1.5 * xy[:, 1:]) + 0.01 * torch.randn(n, 1, generator=g)
    means = torch.cat([xy, z], dim=1)
    scales = torch.log(torch.cat([0.05 + 0.05 * torch.rand(n, 2, generator=g), 0.004 + 0.004 * torch.rand(n, 1, generator=g)], dim=1))
    quats = torch.tensor([1.0, 0.0, 0.0, 0.0]).repeat(n, 1) + 0.15 * torch.randn(n, 4, generator=g)
    return {"means": means, "quats": quats, "scales": scales, "opacities": torch.logit(0.3 + 0.65 * torch.rand(n, 1, generator=g)),
            "features_dc": torch.rand(n, 3, generator=g), "features_rest": 0.05 * torch.randn(n, 15, 3, generator=g)}


def camera_dict(W, H):
    # nerfstudio OpenGL camera at z = +3 looking down -z (towards the sheet), y up
    c2w = torch.tensor([[1.0, 0.0, 0.0, 0.1], [0.0, 1.0, 0.0, -0.05], [0.0, 0.0, 1.0, 3.0]])
    return {"c2w": c2w, "fx": 0.9 * W, "fy": 0.9 * W, "cx": W / 2.0, "cy": H / 2.0, "width": W, "height": H}


def main():
    M = G0.install()
    n, W, H = 500, 40, 32
    params = surface_scene(n, seed=3)
    cam = camera_dict(W, H)
    m = G0.make_model(M, params, torch.tensor([0.0, 0.0, 0.0]), ssim_lambda=0.0, num_downscales=0)
    m.eval()
    camera = G0.make_camera(M, cam)
    captured = {}
    real_knn = M.knn_sk

    def knn_spy(x, y, k):
        out = real_knn(x, y, k)
        captured["knn_x"], captured["knn_y"], captured["knn_idx"] = x.clone(), y.clone(), out.clone()
        return out

    M.knn_sk = knn_spy
    # get_outputs sets gauss_params["normals"]; compute_level_surface_points calls it itself
    z = {"in_" + k: v.numpy() for k, v in params.items()}
    z["cam_c2w"] = cam["c2w"].numpy()
    z["cam_intr"] = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"], W, H], dtype=np.float64)
    for mode in ("closest_gaussian", "analytical"):
        random.seed(11)
        res = M.DNSplatterModel.compute_level_surface_points(m, camera, num_samples=10_000, return_normal=mode)
        for level, o in res.items():
            for k in ("points", "normals", "colors"):
                z[f"level_{mode}_{level}_{k}"] = o[k].detach().numpy()
        print(mode, {lv: tuple(o["points"].shape) for lv, o in res.items()})
    z["knn_points"], z["knn_idx"] = captured["knn_y"].numpy(), captured["knn_idx"].numpy()
    random.seed(12)
    sub = M.DNSplatterModel.compute_level_surface_points(m, camera, num_samples=25, surface_levels=(0.3,))
    z["level_sub25_points"] = sub[0.3]["points"].numpy()
    # density / sdf / density gradient at free samples
    g = torch.Generator().manual_seed(5)
    samples = params["means"][torch.randint(0, n, (300,), generator=g)] + 0.03 * torch.randn(300, 3, generator=g)
    with torch.no_grad():
        idx = real_knn(m.means.data, samples, 16)
        z["q_samples"], z["q_idx"] = samples.numpy(), idx.numpy()
        z["q_density"] = M.DNSplatterModel.get_density(m, samples, closest_gaussians=idx).numpy()
        m.get_density = lambda **kw: M.DNSplatterModel.get_density(m, **kw)
        z["q_sdf"] = M.DNSplatterModel.get_sdf(m, samples, closest_gaussians=idx).numpy()
        z["q_density_grad"] = M.DNSplatterModel.get_density_grad(m, samples, closest_gaussians=idx).numpy()
        z["q_sdf_weight"] = M.DNSplatterModel.get_sdf_weight(m, idx).numpy()
        # sampling / ideal sdf / loss weights (dn_model.py:954-1058, 1167-1204); m.camera was set by get_outputs
        m.training = False
        vis = torch.where(m.radii > 0)[0][::2].clone()
        for tag, vi in (("all", None), ("vis", vis)):
            torch.manual_seed(99)
            pts, ids = M.DNSplatterModel.sample_points_in_gaussians(m, 200, vis_indices=vi)
            z[f"samp_{tag}_points"], z[f"samp_{tag}_ids"] = pts.numpy(), ids.numpy()
        z["samp_vis_indices"] = vis.numpy()
        depth = M.DNSplatterModel.get_outputs(m, camera)["depth"]
        z["ideal_depth_map"] = depth.numpy()
        mask = (torch.rand(H, W, 1, generator=g) > 0.3)
        for tag, mk in (("nomask", None), ("mask", mask)):
            sdf, valid = M.DNSplatterModel.get_ideal_sdf(m, torch.from_numpy(z["samp_all_points"]), depth, camera, mask=mk)
            z[f"ideal_{tag}_sdf"], z[f"ideal_{tag}_valid"] = sdf.numpy(), valid.numpy()
        z["ideal_mask"] = mask.numpy()
        for mode in ("area", "std"):
            z[f"lossw_{mode}"] = M.DNSplatterModel.get_sdf_loss_weight(m, torch.from_numpy(z["samp_all_ids"]), mode=mode).numpy()
    np.savez_compressed(os.path.join(OUT, "dn_sugar_a.npz"), **z)
    print("density range", float(z["q_density"].min()), float(z["q_density"].max()), " >=1:", int((z["q_density"] >= 0.99999).sum()))


if __name__ == "__main__":
    main()
