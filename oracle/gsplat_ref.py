"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the
gsplat==1.0.0 rasterization semantics that dn-splatter calls.

PARITY UNPINNED: gsplat 1.0.0 (pyproject.toml:8 of the reference) is an un-vendored
third-party dependency that is absent from /root/reference and from this image, and
the reference ships no tests / golden vectors for this path.
This file therefore restates gsplat's published algorithm
anchored on the reference's call sites:

  * dn_splatter/dn_model.py:495-516   gsplat.rendering.rasterization(...)   "RGB+ED"
  * dn_splatter/dn_model.py:564-575   gsplat.rasterize_gaussians(...)       normals

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path never does.

Numerics: everything is written as explicit element-wise torch expressions in a
FIXED operation order (no matmul, no fused multiply-add), so that the CUDA
projection kernel — compiled with -fmad=false and the same order — produces
bit-identical radii / tile boxes / sort keys in fp32.  The same code runs in fp64
(pass fp64 tensors) for finite-difference-grade gradient references.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.999
T_STOP = 1e-4

# ----------------------------------------------------------------------------
# per-Gaussian geometry  (gsplat fully_fused_projection)
# ----------------------------------------------------------------------------


def normalize_quat(q: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """wxyz quaternion -> unit components; order: ((w*w + x*x) + y*y) + z*z."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    n2 = ((w * w + x * x) + y * y) + z * z
    inv = 1.0 / torch.sqrt(n2)
    return w * inv, x * inv, y * inv, z * inv


def quat_to_rotmat_entries(q: Tensor):
    """Rows of R(q) (dn_model.py:34 quat_to_rotmat [EXT], wxyz, normalised inside)."""
    w, x, y, z = normalize_quat(q)
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R = [
        [1.0 - 2.0 * (y2 + z2), 2.0 * (xy - wz), 2.0 * (xz + wy)],
        [2.0 * (xy + wz), 1.0 - 2.0 * (x2 + z2), 2.0 * (yz - wx)],
        [2.0 * (xz - wy), 2.0 * (yz + wx), 1.0 - 2.0 * (x2 + y2)],
    ]
    return R


def quat_to_rotmat(q: Tensor) -> Tensor:
    R = quat_to_rotmat_entries(q)
    return torch.stack([torch.stack(r, dim=-1) for r in R], dim=-2)


def _dot3(a0, b0, a1, b1, a2, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2


def project_gaussians(
    means: Tensor,  # [N,3]
    quats: Tensor,  # [N,4] wxyz (any norm)
    scales: Tensor,  # [N,3] ACTIVATED (exp applied)
    viewmat: Tensor,  # [4,4] world->camera
    K: Tensor,  # [3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    fov_size=None,
) -> Dict[str, Tensor]:
    """radii[N] i32, means2d[N,2], depths[N], conics[N,3], compensations[N].
    Culled Gaussians have radius 0 and zeroed float outputs (gsplat leaves them
    uninitialised/zero; nothing downstream reads them)."""
    dt = means.dtype
    W = [[viewmat[i, j] for j in range(3)] for i in range(3)]
    t = [viewmat[i, 3] for i in range(3)]
    px, py, pz = means[:, 0], means[:, 1], means[:, 2]
    # camera-space mean
    x = _dot3(W[0][0], px, W[0][1], py, W[0][2], pz) + t[0]
    y = _dot3(W[1][0], px, W[1][1], py, W[1][2], pz) + t[1]
    z = _dot3(W[2][0], px, W[2][1], py, W[2][2], pz) + t[2]
    in_depth = (z >= near_plane) & (z <= far_plane)

    # world covariance  Sigma = M M^T,  M = R(q) diag(s)
    R = quat_to_rotmat_entries(quats)
    s = [scales[:, 0], scales[:, 1], scales[:, 2]]
    M = [[R[i][j] * s[j] for j in range(3)] for i in range(3)]
    S = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(i, 3):
            S[i][j] = _dot3(M[i][0], M[j][0], M[i][1], M[j][1], M[i][2], M[j][2])
            S[j][i] = S[i][j]
    # camera covariance  Sc = W S W^T
    A = [[_dot3(W[i][0], S[0][j], W[i][1], S[1][j], W[i][2], S[2][j]) for j in range(3)] for i in range(3)]
    Sc = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(i, 3):
            Sc[i][j] = _dot3(A[i][0], W[j][0], A[i][1], W[j][1], A[i][2], W[j][2])
            Sc[j][i] = Sc[i][j]

    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # fov_size = (W, H) of the frame the 1.3 tan(fov) clamp refers to: set it to the FULL frame when rendering a crop window
    # of a larger frame (tests/test_gpu_fullsize.py), so that the window's Gaussians see the full frame's Jacobian clamp
    fw, fh = (width, height) if fov_size is None else fov_size
    tan_fovx = (0.5 * fw) / fx
    tan_fovy = (0.5 * fh) / fy
    lim_x = 1.3 * tan_fovx
    lim_y = 1.3 * tan_fovy
    zs = torch.where(in_depth, z, torch.ones_like(z))  # keep culled lanes finite
    rz = 1.0 / zs
    rz2 = rz * rz
    tx = zs * torch.minimum(lim_x, torch.maximum(-lim_x, x * rz))
    ty = zs * torch.minimum(lim_y, torch.maximum(-lim_y, y * rz))
    J00 = fx * rz
    J02 = -(fx * tx) * rz2
    J11 = fy * rz
    J12 = -(fy * ty) * rz2
    B00 = J00 * Sc[0][0] + J02 * Sc[2][0]
    B01 = J00 * Sc[0][1] + J02 * Sc[2][1]
    B02 = J00 * Sc[0][2] + J02 * Sc[2][2]
    B11 = J11 * Sc[1][1] + J12 * Sc[2][1]
    B12 = J11 * Sc[1][2] + J12 * Sc[2][2]
    a = B00 * J00 + B02 * J02
    b = B01 * J11 + B02 * J12
    c = B11 * J11 + B12 * J12
    mx = (fx * x) * rz + cx
    my = (fy * y) * rz + cy

    det_orig = a * c - b * b
    a = a + eps2d
    c = c + eps2d
    det = a * c - b * b
    ok_det = det > 0
    dets = torch.where(ok_det, det, torch.ones_like(det))
    comp = torch.sqrt(torch.clamp(det_orig / dets, min=0.0))
    inv_det = 1.0 / dets
    conic = torch.stack([c * inv_det, -(b * inv_det), a * inv_det], dim=-1)

    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.01))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    ok = in_depth & ok_det & (radius > radius_clip)
    inside = ~(
        (mx + radius <= 0) | (mx - radius >= width) | (my + radius <= 0) | (my - radius >= height)
    )
    ok = ok & inside
    radii = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)
    okf = ok.to(dt)
    return {
        "radii": radii,
        "means2d": torch.stack([mx, my], dim=-1) * okf[:, None],
        "depths": z * okf,
        "conics": conic * okf[:, None],
        "compensations": comp * okf,
        "mean_cam": torch.stack([x, y, z], dim=-1),
    }


# ----------------------------------------------------------------------------
# spherical harmonics  (gsplat spherical_harmonics, Appendix A1.4)
# ----------------------------------------------------------------------------

SH_C0 = 0.2820947917738781
SH_C1 = 0.48860251190292


def num_sh_bases(degree: int) -> int:
    """dn_model.py:35,139 — (degree+1)^2."""
    return (degree + 1) ** 2


def eval_sh(degree: int, dirs: Tensor, coeffs: Tensor) -> Tensor:
    """dirs[N,3] (un-normalised), coeffs[N,K,3] -> [N,3]; Sloan's fast polynomial form."""
    res = SH_C0 * coeffs[:, 0, :]
    if degree < 1:
        return res
    inorm = 1.0 / torch.sqrt((dirs[:, 0] * dirs[:, 0] + dirs[:, 1] * dirs[:, 1]) + dirs[:, 2] * dirs[:, 2])
    x = (dirs[:, 0] * inorm)[:, None]
    y = (dirs[:, 1] * inorm)[:, None]
    z = (dirs[:, 2] * inorm)[:, None]
    res = res + SH_C1 * ((-y * coeffs[:, 1] + z * coeffs[:, 2]) - x * coeffs[:, 3])
    if degree < 2:
        return res
    z2 = z * z
    fTmp0B = -1.092548430592079 * z
    fC1 = x * x - y * y
    fS1 = 2.0 * x * y
    p6 = 0.9461746957575601 * z2 - 0.3153915652525201
    p7 = fTmp0B * x
    p5 = fTmp0B * y
    p8 = 0.5462742152960395 * fC1
    p4 = 0.5462742152960395 * fS1
    res = res + ((((p4 * coeffs[:, 4] + p5 * coeffs[:, 5]) + p6 * coeffs[:, 6]) + p7 * coeffs[:, 7]) + p8 * coeffs[:, 8])
    if degree < 3:
        return res
    fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658
    fTmp1B = 1.445305721320277 * z
    fC2 = x * fC1 - y * fS1
    fS2 = x * fS1 + y * fC1
    p12 = z * (1.865881662950577 * z2 - 1.119528997770346)
    p13 = fTmp0C * x
    p11 = fTmp0C * y
    p14 = fTmp1B * fC1
    p10 = fTmp1B * fS1
    p15 = -0.5900435899266435 * fC2
    p9 = -0.5900435899266435 * fS2
    res = res + (
        (((((p9 * coeffs[:, 9] + p10 * coeffs[:, 10]) + p11 * coeffs[:, 11]) + p12 * coeffs[:, 12]) + p13 * coeffs[:, 13])
          + p14 * coeffs[:, 14])
        + p15 * coeffs[:, 15]
    )
    return res


# ----------------------------------------------------------------------------
# tile intersection + sort  (gsplat isect_tiles / isect_offset_encode, A3)
# ----------------------------------------------------------------------------


def tile_bounds(means2d: Tensor, radii: Tensor, tile: int, tile_w: int, tile_h: int):
    """tile_min inclusive, tile_max exclusive; (u32)floor / (u32)ceil saturate at 0."""
    r = radii.to(means2d.dtype) / tile
    tcx = means2d[:, 0] / tile
    tcy = means2d[:, 1] / tile
    x0 = torch.clamp(torch.floor(tcx - r), min=0).clamp(max=tile_w).to(torch.int64)
    y0 = torch.clamp(torch.floor(tcy - r), min=0).clamp(max=tile_h).to(torch.int64)
    x1 = torch.clamp(torch.ceil(tcx + r), min=0).clamp(max=tile_w).to(torch.int64)
    y1 = torch.clamp(torch.ceil(tcy + r), min=0).clamp(max=tile_h).to(torch.int64)
    vis = radii > 0
    z = torch.zeros_like(x0)
    x0, y0, x1, y1 = [torch.where(vis, v, z) for v in (x0, y0, x1, y1)]
    return x0, y0, x1, y1


def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile: int, width: int, height: int):
    """Returns tiles_per_gauss[N] i32, isect_ids[I] i64 (sorted), flatten_ids[I] i32 (sorted),
    isect_offsets[tile_h*tile_w] i32.  Keys = (tile_id << 32) | bitcast_i32(depth_fp32);
    stable sort => ties keep ascending Gaussian index (cub::DeviceRadixSort is stable)."""
    tile_w = (width + tile - 1) // tile
    tile_h = (height + tile - 1) // tile
    x0, y0, x1, y1 = tile_bounds(means2d.detach(), radii, tile, tile_w, tile_h)
    nx = x1 - x0
    ny = y1 - y0
    tpg = (nx * ny).to(torch.int64)
    N = means2d.shape[0]
    total = int(tpg.sum())
    gid = torch.repeat_interleave(torch.arange(N, dtype=torch.int64), tpg)
    first = torch.cumsum(tpg, 0) - tpg
    local = torch.arange(total, dtype=torch.int64) - first[gid]
    nxg = torch.clamp(nx[gid], min=1)
    ty = y0[gid] + local // nxg
    txx = x0[gid] + local % nxg
    tile_id = ty * tile_w + txx
    dbits = depths.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = (tile_id << 32) | dbits[gid]
    order = torch.sort(keys, stable=True).indices
    isect_ids = keys[order]
    flatten_ids = gid[order].to(torch.int32)
    counts = torch.bincount((isect_ids >> 32), minlength=tile_w * tile_h)
    offsets = (torch.cumsum(counts, 0) - counts).to(torch.int32)
    return tpg.to(torch.int32), isect_ids, flatten_ids, offsets, (tile_w, tile_h)


# ----------------------------------------------------------------------------
# per-tile front-to-back compositing  (gsplat rasterize_to_pixels fwd, A4;
# legacy rasterize_forward for the normal pass, A6 — same per-pixel rule)
# ----------------------------------------------------------------------------


def rasterize_tiles(
    means2d: Tensor,  # [N,2]
    conics: Tensor,  # [N,3]
    opacities: Tensor,  # [N]
    feats: Tensor,  # [N,C]
    width: int,
    height: int,
    tile: int,
    isect_offsets: Tensor,
    flatten_ids: Tensor,
    chunk: int = 256,
    collect_absgrad: bool = False,
):
    """Returns (out[H,W,C] = sum_i feat_i * alpha_i * T_i, alpha[H,W] = 1 - T_final,
    last_ids[H,W] i32 (index into the sorted list), hooks).  Differentiable w.r.t.
    means2d / conics / opacities / feats through plain autograd; the skip / stop rules
    are (A4): skip if sigma<0 or alpha<1/255; stop BEFORE accumulating when
    T*(1-alpha) <= 1e-4."""
    dt = means2d.dtype
    tile_w = (width + tile - 1) // tile
    tile_h = (height + tile - 1) // tile
    C = feats.shape[1]
    last_ids = torch.zeros(height, width, dtype=torch.int32)
    n_isects = flatten_ids.shape[0]
    offs = isect_offsets.tolist() + [n_isects]
    hooks = []  # (gaussian ids, alpha tensor [P,G] with retain_grad, dx, dy, sigma) for absgrad
    rows_out, rows_alpha = [], []
    for ty in range(tile_h):
        row_out, row_alpha = [], []
        for tx in range(tile_w):
            tid = ty * tile_w + tx
            lo, hi = offs[tid], offs[tid + 1]
            y0, x0 = ty * tile, tx * tile
            y1, x1 = min(y0 + tile, height), min(x0 + tile, width)
            h, w_ = y1 - y0, x1 - x0
            if hi <= lo:
                row_out.append(torch.zeros(h, w_, C, dtype=dt))
                row_alpha.append(torch.zeros(h, w_, dtype=dt))
                continue
            ys = torch.arange(y0, y1, dtype=dt) + 0.5
            xs = torch.arange(x0, x1, dtype=dt) + 0.5
            py = ys[:, None].expand(h, w_).reshape(-1)
            px = xs[None, :].expand(h, w_).reshape(-1)
            P = px.shape[0]
            T = torch.ones(P, dtype=dt)
            done = torch.zeros(P, dtype=torch.bool)
            acc = torch.zeros(P, C, dtype=dt)
            last = torch.zeros(P, dtype=torch.int64)
            for s in range(lo, hi, chunk):
                if bool(done.all()):
                    break
                e = min(s + chunk, hi)
                g = flatten_ids[s:e].long()
                dx = means2d[g, 0][None, :] - px[:, None]
                dy = means2d[g, 1][None, :] - py[:, None]
                ca, cb, cc = conics[g, 0][None, :], conics[g, 1][None, :], conics[g, 2][None, :]
                sigma = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
                alpha = torch.clamp(opacities[g][None, :] * torch.exp(-sigma), max=ALPHA_MAX)
                if collect_absgrad and alpha.requires_grad:
                    alpha.retain_grad()
                    hooks.append((g, alpha, dx.detach(), dy.detach(), sigma.detach()))
                valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
                a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
                om = 1.0 - a_eff
                seq = torch.cat([T[:, None], om], dim=1)
                cp = torch.cumprod(seq, dim=1)  # sequential product, same order as the kernel
                T_before = cp[:, :-1]
                T_after = cp[:, 1:]
                stop = valid & (T_after.detach() <= T_STOP)
                stopped = torch.cummax(stop.to(torch.int8), dim=1).values.bool()
                live = valid & ~stopped & ~done[:, None]
                w = torch.where(live, a_eff * T_before, torch.zeros_like(a_eff))
                acc = acc + w @ feats[g]
                # last contributing index (into the sorted list)
                idx = torch.arange(s, e, dtype=torch.int64)[None, :].expand(P, -1)
                cand = torch.where(live, idx, torch.full_like(idx, -1)).max(dim=1).values
                last = torch.where(cand >= 0, cand, last)
                # transmittance leaving this chunk
                any_stop = stopped[:, -1]
                first_stop = torch.argmax(stop.to(torch.int8), dim=1)
                T_at_stop = torch.gather(T_before, 1, first_stop[:, None])[:, 0]
                T_new = torch.where(any_stop, T_at_stop, T_after[:, -1])
                T = torch.where(done, T, T_new)
                done = done | any_stop
            row_out.append(acc.view(h, w_, C))
            row_alpha.append((1.0 - T).view(h, w_))
            last_ids[y0:y1, x0:x1] = last.view(h, w_).to(torch.int32)
        rows_out.append(torch.cat(row_out, dim=1))
        rows_alpha.append(torch.cat(row_alpha, dim=1))
    out = torch.cat(rows_out, dim=0)
    alpha_img = torch.cat(rows_alpha, dim=0)
    return out, alpha_img, last_ids, hooks


def absgrad_from_hooks(hooks, conics: Tensor, opacities: Tensor, n_gauss: int) -> Tensor:
    """gsplat absgrad (A5): sum over pixels of |v_xy contribution|, v_xy = v_sigma * conic*delta,
    v_sigma = -opac*vis*v_alpha, only where opac*vis <= 0.999 and the pixel accumulated."""
    out = torch.zeros(n_gauss, 2, dtype=conics.dtype)
    for g, alpha, dx, dy, sigma in hooks:
        if alpha.grad is None:
            continue
        va = alpha.grad  # zero where clamped / not used
        ov = opacities[g][None, :].detach() * torch.exp(-sigma)
        v_sigma = -ov * va
        ca, cb, cc = conics[g, 0].detach()[None], conics[g, 1].detach()[None], conics[g, 2].detach()[None]
        vx = (v_sigma * (ca * dx + cb * dy)).abs().sum(0)
        vy = (v_sigma * (cb * dx + cc * dy)).abs().sum(0)
        out.index_add_(0, g, torch.stack([vx, vy], dim=-1))
    return out


# ----------------------------------------------------------------------------
# gsplat.rendering.rasterization(...) as dn-splatter calls it (A1)
# ----------------------------------------------------------------------------


def rasterization(
    means: Tensor,
    quats: Tensor,
    scales: Tensor,  # activated
    opacities: Tensor,  # activated [N]
    colors: Tensor,  # [N,K,3] SH coefficients
    viewmat: Tensor,
    K: Tensor,
    width: int,
    height: int,
    tile_size: int = 16,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    sh_degree: int = 3,
    rasterize_mode: str = "classic",
    eps2d: float = 0.3,
    collect_absgrad: bool = False,
    fov_size=None,
):
    """render[H,W,4] (rgb premultiplied, ED), alpha[H,W,1], info — dn_model.py:495-516."""
    proj = project_gaussians(means, quats, scales, viewmat, K, width, height, eps2d, near_plane, far_plane, fov_size=fov_size)
    radii = proj["radii"]
    opac = opacities
    if rasterize_mode == "antialiased":
        opac = opacities * proj["compensations"]
    tpg, isect_ids, flatten_ids, offsets, (tw, th) = isect_tiles(
        proj["means2d"], radii, proj["depths"], tile_size, width, height
    )
    Rwc = viewmat[:3, :3]
    cam_pos = -(Rwc.T @ viewmat[:3, 3])  # == inverse(viewmat)[:3,3]
    dirs = means - cam_pos[None, :]
    rgb = eval_sh(sh_degree, dirs, colors)
    rgb = torch.where((radii > 0)[:, None], rgb, torch.zeros_like(rgb))
    rgb = torch.clamp(rgb + 0.5, min=0.0)
    feats = torch.cat([rgb, proj["depths"][:, None]], dim=-1)
    out, alpha, last_ids, hooks = rasterize_tiles(
        proj["means2d"], proj["conics"], opac, feats, width, height, tile_size, offsets, flatten_ids,
        collect_absgrad=collect_absgrad,
    )
    ed = out[..., 3:4] / torch.clamp(alpha[..., None], min=1e-10)
    render = torch.cat([out[..., :3], ed], dim=-1)
    info = {
        "radii": radii, "means2d": proj["means2d"], "depths": proj["depths"], "conics": proj["conics"],
        "opacities": opac, "tiles_per_gauss": tpg, "isect_ids": isect_ids, "flatten_ids": flatten_ids,
        "isect_offsets": offsets, "tile_width": tw, "tile_height": th, "last_ids": last_ids,
        "colors": rgb, "compensations": proj["compensations"], "hooks": hooks,
    }
    return render, alpha[..., None], info


def rasterize_gaussians_legacy(
    xys: Tensor, conics: Tensor, colors: Tensor, opacity: Tensor, height: int, width: int, tile: int,
    isect_offsets: Tensor, flatten_ids: Tensor, background: Optional[Tensor] = None,
) -> Tensor:
    """gsplat.rasterize_gaussians (legacy, A6) on the SAME binning as the colour pass
    (documented deviation: the legacy emitter's bbox differs in measure-zero cases).
    background=None -> ones(D).  out = sum + T_final * background."""
    out, alpha, _, _ = rasterize_tiles(xys, conics, opacity, colors, width, height, tile, isect_offsets, flatten_ids)
    if background is None:
        background = torch.ones(colors.shape[1], dtype=colors.dtype)
    return out + (1.0 - alpha)[..., None] * background
