"""csrc/ssim.cu restated step by step in torch (zero-padded separable filter, interior mask, the three partial
derivative maps, transposed filter) and checked against autograd of dn_model.ssim() — pins the closed-form
derivatives the kernel hard-codes; the kernel itself is compared with dn_model.ssim() on the GPU
(tests/test_gpu_model.py::test_fused_ssim_matches_torch)."""
import torch
import torch.nn.functional as F

from dn_splatter_b200.dn_model import ssim


def _filter(img_hwc, win):
    """Zero-padded 'same' separable filter of a [H,W,C] image, as ld_img + separable<> do."""
    t = img_hwc.permute(2, 0, 1)[:, None]  # [C,1,H,W]
    t = F.conv2d(F.pad(t, (5, 5, 0, 0)), win.view(1, 1, 1, 11))
    t = F.conv2d(F.pad(t, (0, 0, 5, 5)), win.view(1, 1, 11, 1))
    return t[:, 0].permute(1, 2, 0)


def kernel_restatement(x, y):
    H, W, C = x.shape
    k = torch.arange(11, dtype=x.dtype) - 5
    win = torch.exp(-(k * k) / (2 * 1.5 * 1.5))
    win = win / win.sum()
    mx, my, exx, eyy, exy = (_filter(q, win) for q in (x, y, x * x, y * y, x * y))
    interior = torch.zeros(H, W, 1, dtype=torch.bool)
    interior[5:H - 5, 5:W - 5] = True
    C1, C2 = 0.0001, 0.0009
    sx, sy, sxy = exx - mx * mx, eyy - my * my, exy - mx * my
    A1, A2, B1, B2 = 2 * mx * my + C1, 2 * sxy + C2, mx * mx + my * my + C1, sx + sy + C2
    inv = 1.0 / (B1 * B2)
    s = A1 * A2 * inv
    zero = torch.zeros_like(s)
    d_xx = torch.where(interior, -s / B2, zero)
    d_xy = torch.where(interior, 2 * A1 * inv, zero)
    d_mu = torch.where(interior, 2 * my * (A2 - A1) * inv - 2 * mx * s / B1 + 2 * mx * s / B2, zero)
    count = (H - 10) * (W - 10) * C
    mean = torch.where(interior, s, zero).sum() / count
    grad = (_filter(d_mu, win) + 2 * x * _filter(d_xx, win) + y * _filter(d_xy, win)) / count
    return mean, grad


def test_ssim_kernel_formulas_match_autograd():
    g = torch.Generator().manual_seed(5)
    for (H, W) in ((23, 37), (16, 16), (40, 11 + 2)):
        x = torch.rand(H, W, 3, generator=g, dtype=torch.float64).requires_grad_(True)
        y = (x.detach() * 0.7 + 0.3 * torch.rand(H, W, 3, generator=g, dtype=torch.float64)).clamp(0, 1)
        ref = ssim(y.permute(2, 0, 1)[None], x.permute(2, 0, 1)[None])
        (gref,) = torch.autograd.grad(ref, x)
        mean, grad = kernel_restatement(x.detach(), y)
        assert torch.allclose(mean, ref.detach(), rtol=1e-12, atol=1e-12)
        assert torch.allclose(grad, gref, rtol=1e-9, atol=1e-12)
