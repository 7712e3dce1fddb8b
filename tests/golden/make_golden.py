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
    for tag, (inp, out) in cases.items():
        arrs = {f"in_{k}": v.detach().numpy() for k, v in inp.items()}
        arrs.update({f"out_{k}": v.detach().numpy() for k, v in out.items()})
        np.savez_compressed(os.path.join(OUT, f"dn_reference_{tag}.npz"), **arrs)
        print(tag, {k: tuple(v.shape) for k, v in arrs.items()})


if __name__ == "__main__":
    main()
