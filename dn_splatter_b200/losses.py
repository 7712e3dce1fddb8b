"""Loss surface of the reference (dn_splatter/losses.py) — same class / enum names and call signatures,
re-implemented.  The terms on the hot path (L1, LogL1, EdgeAwareLogL1, MSE depth; L1 + TV normals) are
evaluated by the fused CUDA kernels via regularization_strategy.DNRegularization; the nn.Modules here are
the plain-torch API objects the reference exposes (used for types the kernels do not cover and for
losses built by user code).  Global-statistic losses (Pearson, Huber) stay torch;
DSSIML1 / SensorDepthLoss / LocalPearson / Adaptive* (NeuS baselines, CUDA-hard-coded paths) are out of
scope and raise NotImplementedError."""
from __future__ import annotations

from enum import Enum
from typing import Optional

import torch
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


def _unsupported(name):
    class _U(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the accelerated hot path (SURVEY.md §2.1 #3)")

    _U.__name__ = name
    return _U


LocalPearsonDepthLoss = _unsupported("LocalPearsonDepthLoss")
AdaptiveDepth = _unsupported("AdaptiveDepth")
AdaptiveNormal = _unsupported("AdaptiveNormal")
DSSIML1 = _unsupported("DSSIML1")
SensorDepthLoss = _unsupported("SensorDepthLoss")


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
