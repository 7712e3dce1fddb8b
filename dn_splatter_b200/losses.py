"""Loss surface of the reference (dn_splatter/losses.py) — same class / enum names and call signatures,
re-implemented.  The terms on the hot path (L1, LogL1, EdgeAwareLogL1, MSE depth; L1 + TV normals) are
evaluated by the fused CUDA kernels via regularization_strategy.DNRegularization; the nn.Modules here are
the plain-torch API objects the reference exposes (used for types the kernels do not cover and for
losses built by user code).  Global-statistic losses (Pearson, Huber) and the losses of the NeuS baselines /
AGS variants (DSSIML1, SensorDepthLoss, LocalPearsonDepthLoss, AdaptiveDepth, AdaptiveNormal) stay plain torch
 and are pinned to goldens produced by the reference's own classes
(tests/golden/make_golden_losses.py); unlike the reference they run on whatever device their inputs live on (the
reference hard-codes "cuda")."""
from __future__ import annotations

import math
from enum import Enum
from typing import Literal, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn


class DepthLossType(Enum):
    """reference losses.py:20-32"""

    MSE = "mse"
    L1 = "L1"
    LogL1 = "LogL1"
    HuberL1 = "HuberL1"
    TV = "TV"
    EdgeAwareLogL1 = "EdgeAwareLogL1"
    EdgeAwareTV = "EdgeAwareTV"
    PearsonDepth = "PearsonDepth"
    LocalPearsonDepthLoss = "LocalPearsonDepthLoss"
    AdaptiveDepth = "AdaptiveDepth"


class NormalLossType(Enum):
    """reference losses.py:355-361"""

    L1 = "L1"
    Smooth = "Smooth"
    AdaptiveNormal = "AdaptiveNormal"


class L1(nn.Module):
    """reference losses.py:155-168"""

    def __init__(self, implementation: str = "scalar", **kwargs):
        super().__init__()
        self.implementation = implementation

    def forward(self, pred, gt):
        d = (pred - gt).abs()
        return d.mean() if self.implementation == "scalar" else d


class LogL1(nn.Module):
    """reference losses.py:171-184"""

    def __init__(self, implementation: str = "scalar", **kwargs):
        super().__init__()
        self.implementation = implementation

    def forward(self, pred, gt):
        v = torch.log(1 + (pred - gt).abs())
        return v.mean() if self.implementation == "scalar" else v


def _edge_weights(rgb: Tensor):
    wx = torch.exp(-(rgb[..., :, :-1, :] - rgb[..., :, 1:, :]).abs().mean(-1, keepdim=True))
    wy = torch.exp(-(rgb[..., :-1, :, :] - rgb[..., 1:, :, :]).abs().mean(-1, keepdim=True))
    return wx, wy


class EdgeAwareLogL1(nn.Module):
    """reference losses.py:187-224"""

    def __init__(self, implementation: str = "scalar", **kwargs):
        super().__init__()
        self.implementation = implementation

    def forward(self, pred: Tensor, gt: Tensor, rgb: Tensor, mask: Optional[Tensor]):
        ll = torch.log(1 + (pred - gt).abs())
        wx, wy = _edge_weights(rgb)
        lx, ly = wx * ll[..., :, :-1, :], wy * ll[..., :-1, :, :]
        if self.implementation == "per-pixel":
            if mask is not None:
                lx = lx * mask[..., :, :-1, :]
                ly = ly * mask[..., :-1, :, :]
            return lx[..., :-1, :, :] + ly[..., :, :-1, :]
        if mask is not None:
            assert mask.shape[:2] == pred.shape[:2]
            lx, ly = lx[mask[..., :, :-1, :]], ly[mask[..., :-1, :, :]]
        return lx.mean() + ly.mean()


class HuberL1(nn.Module):
    """reference losses.py:227-248"""

    def __init__(self, tresh=0.2, implementation: str = "scalar", **kwargs):
        super().__init__()
        self.tresh, self.implementation = tresh, implementation

    def forward(self, pred, gt):
        m = gt != 0
        l1 = (pred[m] - gt[m]).abs()
        d = self.tresh * l1.max()
        loss = torch.where(l1 < d, ((pred - gt) ** 2 + d**2) / (2 * d), l1)
        return loss.mean() if self.implementation == "scalar" else loss


class EdgeAwareTV(nn.Module):
    """reference losses.py:251-276"""

    def forward(self, depth: Tensor, rgb: Tensor):
        wx, wy = _edge_weights(rgb)
        gx = (depth[..., :, :-1, :] - depth[..., :, 1:, :]).abs() * wx
        gy = (depth[..., :-1, :, :] - depth[..., 1:, :, :]).abs() * wy
        return gx.mean() + gy.mean()


class TVLoss(nn.Module):
    """reference losses.py:279-295"""

    def forward(self, pred):
        return (pred[..., :, :-1, :] - pred[..., :, 1:, :]).abs().mean() + \
            (pred[..., :-1, :, :] - pred[..., 1:, :, :]).abs().mean()


class PearsonDepthLoss(nn.Module):
    """reference losses.py:428-452: 1 - mean(z(pred) * z(gt)) with z(x) = (x - mean) / (unbiased std + 1e-6)."""

    def forward(self, pred, gt):
        zp = (pred - pred.mean()) / (pred.std() + 1e-6)
        zg = (gt - gt.mean()) / (gt.std() + 1e-6)
        co = (zp * zg).mean()
        assert not torch.any(torch.isnan(co))
        return 1 - co


def _gaussian_window(size: int, sigma: float, device, dtype) -> Tensor:
    x = torch.arange(size, device=device, dtype=dtype) - (size - 1) / 2
    g = torch.exp(-(x * x) / (2 * sigma * sigma))
    return g / g.sum()


def ssim(img1: Tensor, img2: Tensor, kernel_size: int = 11, sigma: float = 1.5, data_range: float = 1.0) -> Tensor:
    """Mean SSIM of [1,C,H,W] images, Gaussian window, reflect padding then crop — what
    torchmetrics.StructuralSimilarityIndexMeasure(data_range=1.0, kernel_size=11) computes [EXT]
    (reference dn_model.py:180).  Plain torch; the CUDA path of the model uses csrc/ssim.cu (FusedSSIM)."""
    C = img1.shape[1]
    pad = (kernel_size - 1) // 2
    g = _gaussian_window(kernel_size, sigma, img1.device, img1.dtype)
    win = (g[:, None] * g[None, :]).expand(C, 1, kernel_size, kernel_size).contiguous()
    a, b = F.pad(img1, (pad,) * 4, mode="reflect"), F.pad(img2, (pad,) * 4, mode="reflect")
    stack = torch.cat([a, b, a * a, b * b, a * b], dim=0)
    mu = F.conv2d(stack, win, groups=C)
    mu1, mu2, s11, s22, s12 = mu.chunk(5, dim=0)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    v1, v2, v12 = s11 - mu1 * mu1, s22 - mu2 * mu2, s12 - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * v12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (v1 + v2 + c2))
    return m[..., pad:-pad, pad:-pad].mean()


class DSSIML1(nn.Module):
    """alpha * DSSIM + (1 - alpha) * L1 (reference losses.py:73-152).  'per-pixel': 3x3 average-pool SSIM on reflect-padded
    images, clamp((1 - SSIM) / 2, 0, 1), channel mean — pinned to the reference.  'scalar': the reference delegates to
    torchmetrics' Gaussian SSIM (absent here; `ssim()` above restates it, data_range taken from the inputs as torchmetrics
    does when none is given); the multi-scale variant is not restated."""

    def __init__(self, kernel_size: int = 3, alpha: float = 0.85, single_resolution: bool = True,
                 implementation: Literal["scalar", "per-pixel"] = "per-pixel", **kwargs):
        super().__init__()
        self.implementation, self.kernel_size, self.alpha = implementation, kernel_size, alpha
        if implementation == "scalar" and not single_resolution:
            raise NotImplementedError("multi-scale SSIM (torchmetrics MultiScaleStructuralSimilarityIndexMeasure) is not "
                                      "restated: torchmetrics is absent from this image, so it could not be pinned")
        self.C1, self.C2 = 0.01 ** 2, 0.03 ** 2

    def ssim_per_pixel(self, pred, gt):
        k, p = self.kernel_size, int((self.kernel_size - 1) / 2)
        x, y = F.pad(pred, (p,) * 4, mode="reflect"), F.pad(gt, (p,) * 4, mode="reflect")
        pool = lambda t: F.avg_pool2d(t, k, 1)  # noqa: E731
        mu_x, mu_y = pool(x), pool(y)
        sigma_x, sigma_y, sigma_xy = pool(x ** 2) - mu_x ** 2, pool(y ** 2) - mu_y ** 2, pool(x * y) - mu_x * mu_y
        n = (2 * mu_x * mu_y + self.C1) * (2 * sigma_xy + self.C2)
        d = (mu_x ** 2 + mu_y ** 2 + self.C1) * (sigma_x + sigma_y + self.C2)
        return torch.clamp((1 - n / d) / 2, 0, 1)

    def forward(self, pred: Tensor, gt: Tensor):
        if (pred.shape[-1] == 1 or pred.shape[-1] == 3) and pred.dim() == 3:
            pred = pred.permute(2, 0, 1).unsqueeze(0)
        if (gt.shape[-1] == 1 or pred.shape[-1] == 3) and gt.dim() == 3:  # (sic: the reference tests pred here)
            gt = gt.permute(2, 0, 1).unsqueeze(0)
        abs_diff = torch.abs(pred - gt)
        if self.implementation == "scalar":
            dr = float(max(pred.max() - pred.min(), gt.max() - gt.min()))
            s = ssim(pred, gt, kernel_size=self.kernel_size, data_range=dr)
            return self.alpha * (1 - s) / 2 + (1 - self.alpha) * abs_diff.mean()
        return self.alpha * self.ssim_per_pixel(pred, gt).mean(1, True) + (1 - self.alpha) * abs_diff.mean(1, True)


def _sdf_key():
    try:
        from nerfstudio.field_components.field_heads import FieldHeadNames  # type: ignore

        return FieldHeadNames.SDF
    except Exception:  # noqa: BLE001 — nerfstudio is optional
        return "sdf"


class SensorDepthLoss(nn.Module):
    """L1 + free-space + SDF terms against a sensor depth map (reference losses.py:297-352; NeuS-style baselines)."""

    def __init__(self, truncation: float):
        super().__init__()
        self.truncation = truncation

    def forward(self, batch, outputs):
        depth_pred = outputs["depth"]
        depth_gt = batch["sensor_depth"].to(depth_pred.device)[..., None]
        valid = depth_gt > 0.0
        l1_loss = torch.sum(valid * torch.abs(depth_gt - depth_pred)) / (valid.sum() + 1e-6)
        fo = outputs["field_outputs"]
        pred_sdf = (fo[_sdf_key()] if _sdf_key() in fo else fo["sdf"])[..., 0]
        z_vals = outputs["ray_samples"].frustums.starts[..., 0] / outputs["directions_norm"]
        t = self.truncation
        front = valid & (z_vals < (depth_gt - t))
        back = valid & (z_vals > (depth_gt + t))
        sdf_mask = valid & (~front) & (~back)
        n_fs, n_sdf = front.sum(), sdf_mask.sum()
        n = n_fs + n_sdf + 1e-6
        fs_weight, sdf_weight = 1.0 - n_fs / n, 1.0 - n_sdf / n
        free_space_loss = torch.mean((F.relu(t - pred_sdf) * front) ** 2) * fs_weight
        sdf_loss = torch.mean(((z_vals + pred_sdf) - depth_gt) ** 2 * sdf_mask) * sdf_weight
        return l1_loss, free_space_loss, sdf_loss


def _mean_angular_error(pred: Tensor, gt: Tensor) -> Tensor:
    """reference metrics.py:59-74: [B,C,H,W] x2 -> [B,H,W] angle in radians."""
    return torch.acos(torch.clamp(torch.sum(gt * pred, dim=1), -1.0, 1.0))


class AdaptiveDepth(nn.Module):
    """EdgeAwareLogL1 that, from step 7000 on, drops pixels whose confidence is not positive (reference losses.py:386-401)."""

    def __init__(self, implementation: Literal["scalar", "per-pixel"] = "scalar", **kwargs):
        super().__init__()
        self.edgeaware = EdgeAwareLogL1(implementation=implementation)

    def forward(self, pred, gt, gt_image, mask, confidence_map, step):
        if step < 7_000:
            return self.edgeaware(pred, gt, gt_image, mask)
        gt = torch.where(confidence_map > 0, gt, torch.zeros_like(gt))
        return self.edgeaware(pred, gt, gt_image, gt > 0.1)


class AdaptiveNormal(nn.Module):
    """L1 on normals; from step 15000 on only where the angular error is <= 0.1 rad (reference losses.py:404-424)."""

    def __init__(self, implementation: Literal["scalar", "per-pixel"] = "scalar", **kwargs):
        super().__init__()
        self.implementation = implementation
        self.L1 = L1(implementation=self.implementation)

    def forward(self, pred, gt, step):
        if step < 15_000:
            return self.L1(pred, gt)
        diff = _mean_angular_error((pred * 2 - 1).permute(2, 0, 1).unsqueeze(0), (gt * 2 - 1).permute(2, 0, 1).unsqueeze(0))
        keep = ((1 - (diff > 0.1).float()) > 0).squeeze(0)
        return self.L1(pred[keep, :], gt[keep, :])


class LocalPearsonDepthLoss(nn.Module):
    """Mean Pearson loss over int(p_corr * floor(H/box) * floor(W/box)) random box_p x box_p windows (reference
    losses.py:455-485).  The window corners come from torch.randint on the inputs' device (the reference draws them on
    "cuda"); pass `generator` for reproducible draws."""

    def __init__(self):
        super().__init__()
        self.pearson_depth_loss = PearsonDepthLoss()

    def forward(self, depth_pred, depth_gt, box_p=128, p_corr=0.5, generator=None):
        dev = depth_pred.device
        num_box_h, num_box_w = math.floor(depth_pred.shape[0] / box_p), math.floor(depth_pred.shape[1] / box_p)
        max_h, max_w = depth_pred.shape[0] - box_p, depth_pred.shape[1] - box_p
        n_corr = int(p_corr * num_box_h * num_box_w)
        x_0 = torch.randint(0, max_h, size=(n_corr,), device=dev, generator=generator)
        y_0 = torch.randint(0, max_w, size=(n_corr,), device=dev, generator=generator)
        loss = torch.tensor(0.0, device=dev)
        for x0, y0 in zip(x_0.tolist(), y_0.tolist()):
            loss = loss + self.pearson_depth_loss(depth_pred[x0:x0 + box_p, y0:y0 + box_p].reshape(-1),
                                                  depth_gt[x0:x0 + box_p, y0:y0 + box_p].reshape(-1))
        return loss / n_corr


class DepthLoss(nn.Module):
    """Factory (reference losses.py:35-70)."""

    def __init__(self, depth_loss_type: DepthLossType, **kwargs):
        super().__init__()
        self.depth_loss_type = depth_loss_type
        table = {
            DepthLossType.MSE: lambda: nn.MSELoss(), DepthLossType.L1: lambda: L1(**kwargs),
            DepthLossType.LogL1: lambda: LogL1(**kwargs), DepthLossType.HuberL1: lambda: HuberL1(**kwargs),
            DepthLossType.EdgeAwareLogL1: lambda: EdgeAwareLogL1(**kwargs), DepthLossType.EdgeAwareTV: EdgeAwareTV,
            DepthLossType.TV: TVLoss, DepthLossType.PearsonDepth: PearsonDepthLoss,
            DepthLossType.LocalPearsonDepthLoss: LocalPearsonDepthLoss, DepthLossType.AdaptiveDepth: AdaptiveDepth,
        }
        if depth_loss_type not in table:
            raise ValueError(f"Unsupported loss type: {depth_loss_type}")
        self.loss = table[depth_loss_type]()

    def forward(self, *args) -> Tensor:
        return self.loss(*args)


class NormalLoss(nn.Module):
    """Factory (reference losses.py:363-384)."""

    def __init__(self, normal_loss_type: NormalLossType, **kwargs):
        super().__init__()
        self.normal_loss_type = normal_loss_type
        if normal_loss_type == NormalLossType.L1:
            self.loss = L1(**kwargs)
        elif normal_loss_type == NormalLossType.Smooth:
            self.loss = TVLoss()
        elif normal_loss_type == NormalLossType.AdaptiveNormal:
            self.loss = AdaptiveNormal(**kwargs)
        else:
            raise ValueError(f"Unsupported loss type: {normal_loss_type}")

    def forward(self, *args) -> Tensor:
        return self.loss(*args)
