"""Pins the vectorised oracle (oracle/gsplat_ref.py::rasterize_tiles, autograd backward) to a LITERAL scalar
transcription of the algorithm statement in SURVEY.md Appendix A4 (forward: skip / stop rules, last index) and A5
(backward: back-to-front replay with ra = 1/(1-alpha), buffer, v_alpha, v_sigma, clamp rule, absgrad).  Tiny sizes,
float64, so any disagreement is a logic error, not rounding."""
import math

import torch

from oracle import gsplat_ref as G


def _scene(seed=0, n=14, W=24, H=20):
    g = torch.Generator().manual_seed(seed)
    means2d = torch.stack([torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H], -1).double()
    # conics from random SPD covariances (2..8 px sigmas), a few near-opaque splats to trigger the T <= 1e-4 stop
    s1, s2 = 2 + 6 * torch.rand(n, generator=g), 2 + 6 * torch.rand(n, generator=g)
    th = math.pi * torch.rand(n, generator=g)
    c, s = torch.cos(th), torch.sin(th)
    a = c * c * s1 * s1 + s * s * s2 * s2
    b = c * s * (s1 * s1 - s2 * s2)
    d = s * s * s1 * s1 + c * c * s2 * s2
    det = a * d - b * b
    conics = torch.stack([d / det, -b / det, a / det], -1).double()
    opac = (0.3 + 0.699 * torch.rand(n, generator=g)).double()
    opac[:5] = 0.9995  # > 0.999 before the clamp: exercises the `opac * vis <= 0.999` gradient gate
    feats = torch.rand(n, 4, generator=g).double()
    depths = (1 + 9 * torch.rand(n, generator=g)).float()
    radii = torch.full((n,), 40, dtype=torch.int32)  # every splat covers the whole image: long lists, stops happen
    return means2d, conics, opac, feats, depths, radii, W, H


def _literal_forward(means2d, conics, opac, feats, order, W, H):
    C = feats.shape[1]
    out = torch.zeros(H, W, C, dtype=torch.float64)
    alpha_img = torch.zeros(H, W, dtype=torch.float64)
    last = torch.zeros(H, W, dtype=torch.int64)
    for i in range(H):
        for j in range(W):
            px, py = j + 0.5, i + 0.5
            T, acc, cur = 1.0, [0.0] * C, 0
            for pos, gidx in enumerate(order):
                dx, dy = float(means2d[gidx, 0]) - px, float(means2d[gidx, 1]) - py
                ca, cb, cc = (float(v) for v in conics[gidx])
                sigma = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
                alpha = min(0.999, float(opac[gidx]) * math.exp(-sigma))
                if sigma < 0 or alpha < 1 / 255:
                    continue
                next_T = T * (1 - alpha)
                if next_T <= 1e-4:
                    break
                vis = alpha * T
                for k in range(C):
                    acc[k] += float(feats[gidx, k]) * vis
                cur, T = pos, next_T
            out[i, j] = torch.tensor(acc, dtype=torch.float64)
            alpha_img[i, j] = 1 - T
            last[i, j] = cur
    return out, alpha_img, last


def _literal_backward(means2d, conics, opac, feats, order, W, H, alpha_img, last, v_out, v_alpha_out):
    n, C = feats.shape
    v_xy, v_abs = torch.zeros(n, 2, dtype=torch.float64), torch.zeros(n, 2, dtype=torch.float64)
    v_conic, v_feat, v_opac = torch.zeros(n, 3, dtype=torch.float64), torch.zeros(n, C, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    for i in range(H):
        for j in range(W):
            px, py = j + 0.5, i + 0.5
            T_final = 1 - float(alpha_img[i, j])
            T, buf = T_final, [0.0] * C
            for pos in range(int(last[i, j]), -1, -1):
                gidx = order[pos]
                dx, dy = float(means2d[gidx, 0]) - px, float(means2d[gidx, 1]) - py
                ca, cb, cc = (float(v) for v in conics[gidx])
                sigma = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
                vis = math.exp(-sigma)
                op = float(opac[gidx])
                alpha = min(0.999, op * vis)
                if sigma < 0 or alpha < 1 / 255:
                    continue
                ra = 1 / (1 - alpha)
                T *= ra
                fac = alpha * T
                v_a = 0.0
                for k in range(C):
                    v_feat[gidx, k] += fac * float(v_out[i, j, k])
                    v_a += (float(feats[gidx, k]) * T - buf[k] * ra) * float(v_out[i, j, k])
                v_a += T_final * ra * float(v_alpha_out[i, j])
                if op * vis <= 0.999:
                    v_sigma = -op * vis * v_a
                    v_conic[gidx] += torch.tensor([0.5 * v_sigma * dx * dx, v_sigma * dx * dy, 0.5 * v_sigma * dy * dy], dtype=torch.float64)
                    gx, gy = v_sigma * (ca * dx + cb * dy), v_sigma * (cb * dx + cc * dy)
                    v_xy[gidx] += torch.tensor([gx, gy], dtype=torch.float64)
                    v_abs[gidx] += torch.tensor([abs(gx), abs(gy)], dtype=torch.float64)
                    v_opac[gidx] += vis * v_a
                for k in range(C):
                    buf[k] += float(feats[gidx, k]) * fac
    return v_xy, v_abs, v_conic, v_feat, v_opac


def test_vectorised_oracle_equals_literal_appendix_a4_a5():
    means2d, conics, opac, feats, depths, radii, W, H = _scene()
    tpg, isect_ids, flat, offs, (tw, th) = G.isect_tiles(means2d, radii, depths, 16, W, H)
    m, c, o, f = (t.clone().requires_grad_(True) for t in (means2d, conics, opac, feats))
    out, alpha, last_ids, hooks = G.rasterize_tiles(m, c, o, f, W, H, 16, offs, flat, chunk=5, collect_absgrad=True)
    offs_l = offs.tolist() + [flat.shape[0]]
    g = torch.Generator().manual_seed(5)
    v_out, v_al = torch.rand(H, W, 4, generator=g).double(), torch.rand(H, W, generator=g).double()
    ((out * v_out).sum() + (alpha * v_al).sum()).backward()
    stopped = 0
    for ty in range(th):
        for tx in range(tw):
            t = ty * tw + tx
            order = flat[offs_l[t]:offs_l[t + 1]].tolist()
            y0, x0 = ty * 16, tx * 16
            y1, x1 = min(y0 + 16, H), min(x0 + 16, W)
            # literal pass over this tile's pixels with the tile's sorted list (pixel coordinates are absolute)
            sub = lambda a: a[y0:y1, x0:x1]  # noqa: E731
            lo, la, ll = _literal_forward(means2d - torch.tensor([x0, y0], dtype=torch.float64), conics, opac, feats, order, x1 - x0, y1 - y0)
            torch.testing.assert_close(sub(out.detach()), lo, rtol=1e-12, atol=1e-12)
            torch.testing.assert_close(sub(alpha.detach()), la, rtol=1e-12, atol=1e-12)
            assert torch.equal(sub(last_ids).long(), ll + offs_l[t])
            stopped += int((la > 1 - 1.1e-4 * 10).sum())
            vx, vabs, vc, vf, vo = _literal_backward(means2d - torch.tensor([x0, y0], dtype=torch.float64), conics, opac, feats, order, x1 - x0, y1 - y0,
                                                     la, ll, sub(v_out), sub(v_al))
            if t == 0:
                tot = [vx, vabs, vc, vf, vo]
            else:
                tot = [a + b for a, b in zip(tot, [vx, vabs, vc, vf, vo])]
    assert stopped > 0, "the scene should contain saturated pixels (T <= 1e-4 stop rule exercised)"
    torch.testing.assert_close(m.grad, tot[0], rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(c.grad, tot[2], rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(f.grad, tot[3], rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(o.grad, tot[4], rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(G.absgrad_from_hooks(hooks, conics, opac, means2d.shape[0]), tot[1], rtol=1e-9, atol=1e-12)


def test_sh_polynomial_form_equals_the_canonical_real_sh_table():
    """eval_sh uses Sloan's factored polynomials; the canonical (Inria / plenoxels) table of real SH up to degree 3 is the
    independent statement: result = C0 c0 - C1 y c1 + C1 z c2 - C1 x c3 + C2[0] xy c4 + C2[1] yz c5 + ..."""
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    g = torch.Generator().manual_seed(0)
    dirs = torch.randn(50, 3, generator=g).double()
    coeffs = torch.randn(50, 16, 3, generator=g).double()
    u = dirs / dirs.norm(dim=-1, keepdim=True)
    x, y, z = u[:, 0:1], u[:, 1:2], u[:, 2:3]
    sh = coeffs
    want0 = C0 * sh[:, 0]
    want1 = want0 - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    want2 = (want1 + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
             + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    want3 = (want2 + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10] + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11]
             + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12] + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13]
             + C3[5] * z * (xx - yy) * sh[:, 14] + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    for deg, want in ((0, want0), (1, want1), (2, want2), (3, want3)):
        torch.testing.assert_close(G.eval_sh(deg, dirs, coeffs), want, rtol=1e-9, atol=1e-12)


def test_projection_entry_formulas_equal_the_matrix_statement():
    """project_gaussians spells every matrix product out element by element (fixed rounding order for the CUDA kernel);
    the independent statement is Appendix A2 in matrix form: Sigma = R S S^T R^T, Sigma_c = W Sigma W^T,
    cov2d = J Sigma_c J^T + 0.3 I, conic = inverse, radius = ceil(3 sqrt(lambda_max))."""
    from oracle import dn_ref
    from tests.helpers import scene_and_camera

    params, cam = scene_and_camera(300, 96, 64, view=2)
    W, H = 96, 64
    vm = dn_ref.get_viewmat(cam["c2w"]).double()
    K = dn_ref.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"], torch.float64)
    means, quats, scales = params["means"].double(), params["quats"].double(), torch.exp(params["scales"]).double()
    got = G.project_gaussians(means, quats, scales, vm, K, W, H)
    qn = quats / quats.norm(dim=-1, keepdim=True)
    w, x, y, z = qn.unbind(-1)
    R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                     torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)
    M = R * scales[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    Wm, t = vm[:3, :3], vm[:3, 3]
    pc = means @ Wm.T + t
    Sc = Wm @ Sigma @ Wm.T
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    X, Y, Z = pc.unbind(-1)
    limx, limy = 1.3 * 0.5 * W / fx, 1.3 * 0.5 * H / fy
    tx, ty = Z * torch.clamp(X / Z, -limx, limx), Z * torch.clamp(Y / Z, -limy, limy)
    J = torch.zeros(len(means), 2, 3, dtype=torch.float64)
    J[:, 0, 0], J[:, 0, 2], J[:, 1, 1], J[:, 1, 2] = fx / Z, -fx * tx / Z**2, fy / Z, -fy * ty / Z**2
    cov2 = J @ Sc @ J.transpose(1, 2) + 0.3 * torch.eye(2, dtype=torch.float64)
    con = torch.linalg.inv(cov2)
    lam = torch.linalg.eigvalsh(cov2)[:, 1]
    radius = torch.ceil(3 * torch.sqrt(lam))
    mean2d = torch.stack([fx * X / Z + cx, fy * Y / Z + cy], -1)
    ok = (Z >= 0.01) & ~((mean2d[:, 0] + radius <= 0) | (mean2d[:, 0] - radius >= W) | (mean2d[:, 1] + radius <= 0) | (mean2d[:, 1] - radius >= H))
    assert torch.equal(got["radii"] > 0, ok)
    v = ok
    # lambda_max here is exact; the kernel statement uses mid + sqrt(max(0.01, mid^2 - det)) which is the same number
    assert torch.equal(got["radii"][v].double(), radius[v])
    torch.testing.assert_close(got["means2d"][v], mean2d[v], rtol=1e-10, atol=1e-9)
    torch.testing.assert_close(got["depths"][v], Z[v], rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(got["conics"][v], torch.stack([con[:, 0, 0], con[:, 0, 1], con[:, 1, 1]], -1)[v], rtol=1e-8, atol=1e-10)
