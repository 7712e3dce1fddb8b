"""Generates tests/golden/dn_reference_*.npz by importing the REFERENCE's own modules
(/root/reference/dn_splatter/{losses,regularization_strategy,utils/normal_utils,utils/camera_utils}.py)
unmodified through a stub shim.  Run only where /root/reference exists:

    python tests/golden/make_golden.py

The committed .npz files pin oracle/dn_ref.py (tests/test_oracle_golden.py) and, through it, the CUDA
surface-normal / loss kernels.  gsplat itself is absent everywhere, so the rasterizer has no such pin.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_shim():
    pkg = types.ModuleType("dn_splatter")
    pkg.__path__ = [os.path.join(REF, "dn_splatter")]
    sys.modules["dn_splatter"] = pkg
    utils = types.ModuleType("dn_splatter.utils")
    utils.__path__ = [os.path.join(REF, "dn_splatter", "utils")]
    sys.modules["dn_splatter.utils"] = utils

    class _Dummy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    tm = types.ModuleType("torchmetrics")
    tmi = types.ModuleType("torchmetrics.image")
    for n in ("StructuralSimilarityIndexMeasure", "MultiScaleStructuralSimilarityIndexMeasure", "PeakSignalNoiseRatio"):
        setattr(tmi, n, _Dummy)
    tml = types.ModuleType("torchmetrics.image.lpip")
    tml.LearnedPerceptualImagePatchSimilarity = _Dummy
    tmf = types.ModuleType("torchmetrics.functional")
    tmf.mean_squared_error = lambda a, b: ((a - b) ** 2).mean()
    tm.image, tmi.lpip, tm.functional = tmi, tml, tmf
    sys.modules.update({"torchmetrics": tm, "torchmetrics.image": tmi, "torchmetrics.image.lpip": tml,
                        "torchmetrics.functional": tmf})
    ns = types.ModuleType("nerfstudio")
    fc = types.ModuleType("nerfstudio.field_components")
    fh = types.ModuleType("nerfstudio.field_components.field_heads")

    class FieldHeadNames:
        pass

    fh.FieldHeadNames = FieldHeadNames
    sys.modules.update({"nerfstudio": ns, "nerfstudio.field_components": fc,
                        "nerfstudio.field_components.field_heads": fh})


def main():
    install_shim()
    from dn_splatter.losses import DepthLoss, DepthLossType, EdgeAwareLogL1, L1, LogL1, TVLoss
    from dn_splatter.regularization_strategy import DNRegularization
    from dn_splatter.utils.normal_utils import normal_from_depth_image

    g = torch.Generator().manual_seed(1234)
    cases = {}
    for tag, (H, W) in {"a": (24, 32), "b": (17, 21)}.items():
        pred_depth = 0.5 + 4 * torch.rand(H, W, 1, generator=g)
        gt_depth = pred_depth + 0.3 * torch.randn(H, W, 1, generator=g)
        gt_depth[torch.rand(H, W, 1, generator=g) < 0.15] = 0.0  # invalid pixels (mask = gt > 0.1)
        pred_normal = torch.rand(H, W, 3, generator=g)
        gt_normal = torch.rand(H, W, 3, generator=g)
        gt_img = torch.rand(H, W, 3, generator=g).clamp(min=10 / 255.0)
        scales = torch.randn(50, 3, generator=g) - 3
        fx, fy, cx, cy = 0.9 * W, 0.85 * W, W / 2.0 + 0.25, H / 2.0 - 0.5
        out = {}
        # surface normals exactly as dn_model.py:589-603 calls it
        n = normal_from_depth_image(depths=pred_depth, fx=fx, fy=fy, cx=cx, cy=cy, img_size=(W, H),
                                    c2w=torch.eye(4), device=torch.device("cpu"), smooth=False)
        out["normal_from_depth"] = n
        sn = n @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))
        out["surface_normal_output"] = (1 + sn) / 2
        mask = gt_depth > 0.1
        out["edge_aware_logl1"] = EdgeAwareLogL1()(pred_depth, gt_depth, gt_img, mask)
        out["logl1"] = LogL1()(pred_depth[mask], gt_depth[mask])
        out["l1"] = L1()(pred_normal, gt_normal)
        out["tv"] = TVLoss()(pred_normal)
        for lam in (0.2, 0.5):
            reg = DNRegularization(depth_lambda=lam)
            out[f"dn_reg_lambda{lam}"] = reg(pred_depth=pred_depth, gt_depth=gt_depth, pred_normal=pred_normal,
                                             gt_normal=gt_normal, scales=scales, gt_img=gt_img)
        for t in (DepthLossType.LogL1, DepthLossType.L1, DepthLossType.MSE):
            reg = DNRegularization(depth_lambda=0.2)
            reg.depth_loss_type = t
            reg.depth_loss = DepthLoss(t)
            out[f"dn_reg_{t.value}"] = reg(pred_depth=pred_depth, gt_depth=gt_depth, pred_normal=pred_normal,
                                           gt_normal=gt_normal, scales=scales, gt_img=gt_img)
        reg = DNRegularization(depth_lambda=0.2)
        reg.depth_loss = None
        out["dn_reg_nodepth"] = reg(pred_depth=pred_depth, gt_depth=gt_depth, pred_normal=pred_normal,
                                    gt_normal=gt_normal, scales=scales, gt_img=gt_img)
        # gradients of the default regulariser w.r.t. the rendered maps (what the fused backward must produce)
        pd = pred_depth.clone().requires_grad_(True)
        pn = pred_normal.clone().requires_grad_(True)
        sc = scales.clone().requires_grad_(True)
        DNRegularization(depth_lambda=0.2)(pred_depth=pd, gt_depth=gt_depth, pred_normal=pn, gt_normal=gt_normal,
                                           scales=sc, gt_img=gt_img).backward()
        out["grad_pred_depth"], out["grad_pred_normal"], out["grad_scales"] = pd.grad, pn.grad, sc.grad
        inp = dict(pred_depth=pred_depth, gt_depth=gt_depth, pred_normal=pred_normal, gt_normal=gt_normal,
                   gt_img=gt_img, scales=scales, intr=torch.tensor([fx, fy, cx, cy], dtype=torch.float64))
        cases[tag] = (inp, out)
    # ---- the nn.Module surface of losses.py and the AGS-Mesh strategy (torch paths of this repo) -------------------
    from dn_splatter.losses import EdgeAwareTV, HuberL1, PearsonDepthLoss
    from dn_splatter.regularization_strategy import AGSMeshRegularization, find_edges, mean_angular_error

    H, W = 20, 28
    pd = 0.5 + 4 * torch.rand(H, W, 1, generator=g)
    gd = pd + 0.3 * torch.randn(H, W, 1, generator=g)
    gd[torch.rand(H, W, 1, generator=g) < 0.15] = 0.0
    img = torch.rand(H, W, 3, generator=g).clamp(min=10 / 255.0)
    mask = gd > 0.1
    mod_in = dict(pd=pd, gd=gd, img=img)
    mod_out = {
        "edge_aware_logl1_pp": EdgeAwareLogL1(implementation="per-pixel")(pd, gd, img, mask),
        "logl1_pp": LogL1(implementation="per-pixel")(pd, gd),
        "l1_pp": L1(implementation="per-pixel")(pd, gd),
        "huber": HuberL1()(pd, gd),
        "edge_aware_tv": EdgeAwareTV()(pd[None], img[None]),
        "pearson": PearsonDepthLoss()(pd, gd + 0.01),
        "mse": DepthLoss(DepthLossType.MSE)(pd, gd),
    }
    # AGS-Mesh: [C,H,W] normals in [-1,1]
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    gn = torch.stack([0.3 + 0.1 * torch.sin(xx / 3), 0.4 + 0.1 * torch.cos(yy / 4), 0.8 + 0 * xx])
    gn[:, :, W // 2:] = gn[:, :, W // 2:].flip(0)  # one crease: find_edges should fire there and only there
    gn = torch.nn.functional.normalize(gn, dim=0)
    sn = torch.nn.functional.normalize(gn + 0.05 * torch.randn(3, H, W, generator=g), dim=0)
    pn = torch.nn.functional.normalize(torch.randn(3, H, W, generator=g), dim=0)
    conf = (torch.rand(H, W, 1, generator=g) > 0.3).float()
    scales = torch.randn(40, 3, generator=g) - 3
    mod_in.update(sn=sn, gn=gn, pn=pn, conf=conf, scales=scales)
    mod_out["find_edges_3"] = find_edges(gn).float()
    mod_out["find_edges_1"] = find_edges(pd.permute(2, 0, 1)).float()
    mod_out["mean_angular_error"] = mean_angular_error(sn, gn)
    ags = AGSMeshRegularization()
    for step in (100, 8000, 16000):
        mod_out[f"ags_normal_{step}"] = torch.as_tensor(ags.get_normal_loss(step, sn, gn, pn)).float()
    mod_out["ags_depth_100"] = ags.get_depth_loss(step=100, pred_depth=pd, gt_depth=gd, confidence_map=conf, gt_img=img)
    mod_out["ags_total_100"] = ags(step=100, pred_depth=pd, gt_depth=gd, surf_normal=sn, gt_normal=gn, pred_normal=pn,
                                   confidence_map=conf, scales=scales, gt_img=img)
    arrs = {f"in_{k}": v.detach().numpy() for k, v in mod_in.items()}
    arrs.update({f"out_{k}": v.detach().numpy() for k, v in mod_out.items()})
    np.savez_compressed(os.path.join(OUT, "dn_reference_modules.npz"), **arrs)
    print("modules", {k: tuple(v.shape) for k, v in arrs.items() if k.startswith("out_")})

    for tag, (inp, out) in cases.items():
        arrs = {f"in_{k}": v.detach().numpy() for k, v in inp.items()}
        arrs.update({f"out_{k}": v.detach().numpy() for k, v in out.items()})
        np.savez_compressed(os.path.join(OUT, f"dn_reference_{tag}.npz"), **arrs)
        print(tag, {k: tuple(v.shape) for k, v in arrs.items()})


if __name__ == "__main__":
    main()
