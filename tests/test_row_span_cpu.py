"""Conservativeness of the precise-hit emission formula (csrc/binning.cu: load_hit_gauss + row_span), restated in numpy
float32 and checked against brute force: every tile that contains a pixel centre with sigma <= lim must lie inside the
emitted span of its row — for ordinary, needle-like, huge and off-screen splats.  (The CUDA kernel itself is covered
on the GPU by test_precise_hit_lists_render_bit_identical_images.)"""
import numpy as np
import pytest

f32 = np.float32
TILE = 16


def spans(mx, my, A, B, C, L, x0, y0, nx, ny):
    """Mirror of load_hit_gauss + row_span for one Gaussian; returns [(lo, hi)] per tile row of the box."""
    mx, my, A, B, C, L = map(f32, (mx, my, A, B, C, L))
    det = f32(A * C - B * B)
    if not (L > 0) or not (det > 0):
        return [(x0, x0)] * ny
    invA, bac, twoAL = f32(1) / A, -det, f32(2) * A * L
    ex_max = np.sqrt(f32(2) * L * C / det, dtype=f32) * f32(1.0001) + f32(0.01)
    ey_max = np.sqrt(f32(2) * L * A / det, dtype=f32) * f32(1.0001) + f32(0.01)
    ey_star = -B * ex_max / C
    out = []
    for r in range(ny):
        ty = y0 + r
        lo, hi = x0, x0 + nx
        e0 = f32(ty * TILE + 0.5) - my - f32(0.01)
        e1 = f32(ty * TILE + TILE - 1 + 0.5) - my + f32(0.01)
        if e0 > ey_max or e1 < -ey_max:
            out.append((lo, lo))
            continue
        e0, e1 = max(e0, -ey_max), min(e1, ey_max)
        s0 = np.sqrt(max(f32(bac * e0 * e0 + twoAL), f32(0)), dtype=f32)
        s1 = np.sqrt(max(f32(bac * e1 * e1 + twoAL), f32(0)), dtype=f32)
        xmax = max((-B * e0 + s0) * invA, (-B * e1 + s1) * invA)
        xmin = min((-B * e0 - s0) * invA, (-B * e1 - s1) * invA)
        if e0 <= ey_star <= e1:
            xmax = ex_max
        if e0 <= -ey_star <= e1:
            xmin = -ex_max
        xmax = xmax + f32(0.01) + f32(1e-5) * abs(xmax)
        xmin = xmin - f32(0.01) - f32(1e-5) * abs(xmin)
        t_lo = int(np.ceil((mx + xmin - f32(15.5)) / f32(TILE)))
        t_hi = int(np.floor((mx + xmax - f32(0.5)) / f32(TILE))) + 1
        out.append((max(lo, t_lo), min(hi, t_hi)))
    return out


def brute_force_tiles(mx, my, A, B, C, L, tiles_x, tiles_y):
    xs = np.arange(tiles_x * TILE, dtype=np.float64) + 0.5
    ys = np.arange(tiles_y * TILE, dtype=np.float64) + 0.5
    dx, dy = mx - xs[None, :], my - ys[:, None]
    sigma = 0.5 * (A * dx * dx + C * dy * dy) + B * dx * dy
    hit = sigma <= L
    return hit.reshape(tiles_y, TILE, tiles_x, TILE).any(axis=(1, 3))


@pytest.mark.parametrize("seed", range(6))
def test_row_spans_cover_every_reachable_tile(seed):
    rng = np.random.default_rng(seed)
    tiles_x, tiles_y = 20, 12
    kept = total = 0
    for _ in range(300):
        # covariance from random axes; every third splat is a needle (axis ratio up to 1:200)
        s1 = rng.uniform(0.6, 60.0)
        s2 = s1 / rng.uniform(1.0, 200.0 if rng.random() < 0.33 else 6.0)
        s2 = max(s2, 0.55)
        th = rng.uniform(0, np.pi)
        c, s = np.cos(th), np.sin(th)
        cov = np.array([[c * c * s1 * s1 + s * s * s2 * s2, c * s * (s1 * s1 - s2 * s2)],
                        [c * s * (s1 * s1 - s2 * s2), s * s * s1 * s1 + c * c * s2 * s2]])
        con = np.linalg.inv(cov)
        A, B, C = con[0, 0], con[0, 1], con[1, 1]
        mx, my = rng.uniform(-40, tiles_x * TILE + 40), rng.uniform(-40, tiles_y * TILE + 40)
        opac = rng.uniform(0.005, 1.0)
        L = np.log(255.0 * opac) + 0.1  # cull_lim of project_fwd
        want = brute_force_tiles(mx, my, A, B, C, np.log(255.0 * opac), tiles_x, tiles_y)  # true reach (no margin)
        got = spans(mx, my, A, B, C, L, 0, 0, tiles_x, tiles_y)
        for ty in range(tiles_y):
            lo, hi = got[ty]
            need = np.nonzero(want[ty])[0]
            if need.size:
                assert lo <= need.min() and hi > need.max(), (seed, mx, my, A, B, C, opac, ty, (lo, hi), need)
            kept += max(hi - lo, 0)
        total += int(want.sum())
    # and the spans are tight: not more than ~1.6x the truly reachable tiles on this mix
    assert kept <= 1.6 * total + 50, (kept, total)
