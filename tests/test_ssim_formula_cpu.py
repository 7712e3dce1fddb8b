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


def _tile_mirror(x, y, v=1.0):
    """csrc/ssim.cu transcribed loop for loop (16x16 tiles, 26x26 halo, horizontal then vertical 11-tap pass, interior
    mask, partial maps, second pass over the maps) in numpy float64 — pins the tile / halo index arithmetic."""
    import numpy as np

    H, W, C = x.shape
    k = np.arange(11) - 5.0
    win = np.exp(-(k * k) / 4.5)
    win /= win.sum()
    T, R, HT = 16, 5, 26

    def ld(img, i, j, c):
        return img[i, j, c] if (0 <= i < H and 0 <= j < W) else 0.0

    def separable(s_in):  # s_in [Q,26,26] -> out [Q,16,16]
        Q = s_in.shape[0]
        mid = np.zeros((Q, HT, T))
        for r in range(HT):
            for c in range(T):
                for q in range(Q):
                    mid[q, r, c] = sum(win[kk] * s_in[q, r, c + kk] for kk in range(11))
        out = np.zeros((Q, T, T))
        for ty in range(T):
            for tx in range(T):
                for q in range(Q):
                    out[q, ty, tx] = sum(win[kk] * mid[q, ty + kk, tx] for kk in range(11))
        return out

    dmaps = np.zeros((3, H, W, C))
    total = 0.0
    tiles_y, tiles_x = (H + T - 1) // T, (W + T - 1) // T
    for c in range(C):
        for by in range(tiles_y):
            for bx in range(tiles_x):
                i0, j0 = by * T - R, bx * T - R
                s_in = np.zeros((5, HT, HT))
                for r in range(HT):
                    for q in range(HT):
                        xv, yv = ld(x, i0 + r, j0 + q, c), ld(y, i0 + r, j0 + q, c)
                        s_in[:, r, q] = (xv, yv, xv * xv, yv * yv, xv * yv)
                f = separable(s_in)
                for ty in range(T):
                    for tx in range(T):
                        i, j = by * T + ty, bx * T + tx
                        interior = R <= i < H - R and R <= j < W - R
                        s = d_mu = d_xx = d_xy = 0.0
                        if interior:
                            mx, my = f[0, ty, tx], f[1, ty, tx]
                            sx, sy, sxy = f[2, ty, tx] - mx * mx, f[3, ty, tx] - my * my, f[4, ty, tx] - mx * my
                            A1, A2, B1, B2 = 2 * mx * my + 1e-4, 2 * sxy + 9e-4, mx * mx + my * my + 1e-4, sx + sy + 9e-4
                            inv = 1.0 / (B1 * B2)
                            s = A1 * A2 * inv
                            d_xx, d_xy = -s / B2, 2 * A1 * inv
                            d_mu = 2 * my * (A2 - A1) * inv - 2 * mx * s / B1 + 2 * mx * s / B2
                        if i < H and j < W:
                            dmaps[:, i, j, c] = (d_mu, d_xx, d_xy)
                        total += s
    count = (H - 2 * R) * (W - 2 * R) * C
    grad = np.zeros((H, W, C))
    for c in range(C):
        for by in range(tiles_y):
            for bx in range(tiles_x):
                i0, j0 = by * T - R, bx * T - R
                s_in = np.zeros((3, HT, HT))
                for r in range(HT):
                    for q in range(HT):
                        for mth in range(3):
                            s_in[mth, r, q] = ld(dmaps[mth], i0 + r, j0 + q, c)
                f = separable(s_in)
                for ty in range(T):
                    for tx in range(T):
                        i, j = by * T + ty, bx * T + tx
                        if i < H and j < W:
                            grad[i, j, c] = (v / count) * (f[0, ty, tx] + 2 * x[i, j, c] * f[1, ty, tx] + y[i, j, c] * f[2, ty, tx])
    return total / count, grad


def test_ssim_tile_mirror_matches_autograd():
    g = torch.Generator().manual_seed(9)
    H, W = 19, 35  # ragged: 2 x 3 tiles, last ones partial
    x = torch.rand(H, W, 1, generator=g, dtype=torch.float64).requires_grad_(True)
    y = (x.detach() * 0.5 + 0.5 * torch.rand(H, W, 1, generator=g, dtype=torch.float64))
    ref = ssim(y.permute(2, 0, 1)[None], x.permute(2, 0, 1)[None])
    (gref,) = torch.autograd.grad(ref, x)
    mean, grad = _tile_mirror(x.detach().numpy(), y.numpy(), v=1.0)
    assert abs(mean - float(ref)) < 1e-12
    assert float((torch.from_numpy(grad) - gref).abs().max()) < 1e-12
