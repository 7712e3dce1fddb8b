"""optim.FusedAdam: the kernel's update rule (restated by FusedAdam.reference_step) against torch.optim.Adam, the
Optimizer surface densification relies on, and the no-CPU-path error."""
import pytest
import torch

from dn_splatter_b200.optim import FusedAdam, bias_corrections


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = {"means": (50, 3), "scales": (50, 3), "quats": (50, 4), "features_rest": (50, 15, 3), "opacities": (50, 1)}
    return {k: torch.nn.Parameter(torch.randn(*s, generator=g)) for k, s in shapes.items()}


LRS = {"means": 1.6e-4, "scales": 5e-3, "quats": 1e-3, "features_rest": 2.5e-3 / 20, "opacities": 5e-2}


def test_update_rule_matches_torch_adam_over_steps():
    a, b = _params(1), _params(1)
    fused = FusedAdam([{"params": [p], "lr": LRS[k], "eps": 1e-15, "name": k} for k, p in a.items()])
    ref = {k: torch.optim.Adam([p], lr=LRS[k], eps=1e-15) for k, p in b.items()}
    g = torch.Generator().manual_seed(2)
    for step in range(25):
        for k in a:
            grad = torch.randn(a[k].shape, generator=g) * 10.0 ** (-(len(k) % 6))  # a fixed magnitude per group
            if step % 7 == 3:
                grad = grad * (torch.rand(grad.shape, generator=g) > 0.8)  # mostly-zero gradients (invisible Gaussians)
            a[k].grad, b[k].grad = grad.clone(), grad.clone()
        fused.reference_step()
        for o in ref.values():
            o.step()
        if step == 10:  # the means schedule changes lr between steps
            for grp in fused.param_groups:
                if grp["name"] == "means":
                    grp["lr"] = 1e-5
            ref["means"].param_groups[0]["lr"] = 1e-5
    for k in a:
        st_a, st_b = fused.state[a[k]], ref[k].state[b[k]]
        assert int(st_a["step"]) == int(st_b["step"]) == 25
        for x, y in ((st_a["exp_avg"], st_b["exp_avg"]), (st_a["exp_avg_sq"], st_b["exp_avg_sq"]), (a[k].data, b[k].data)):
            # fp32 rounding of a 25-step recurrence (lerp vs m + w (g - m)); with cancellation; a wrong formula is off by >> 1e-5
            assert float((x - y).abs().max()) <= 1e-5 * float(y.abs().max()), (k, float((x - y).abs().max()))


def test_bias_corrections():
    bc1, bc2s = bias_corrections(1, 0.9, 0.999)
    assert abs(bc1 - 0.1) < 1e-12 and abs(bc2s - 0.001 ** 0.5) < 1e-9


def test_optimizer_surface_for_densification():
    from dn_splatter_b200.densify import _resize_adam_state

    a = _params(3)
    fused = FusedAdam([{"params": [p], "lr": LRS[k], "eps": 1e-15, "name": k} for k, p in a.items()])
    for p in a.values():
        p.grad = torch.ones_like(p)
    fused.reference_step()
    old = a["means"]
    new = torch.nn.Parameter(torch.cat([old.data, old.data[:5]]))
    _resize_adam_state(fused, old, new, lambda t: torch.cat([t, torch.zeros_like(t[:5])]))
    assert new in fused.state and old not in fused.state
    assert fused.state[new]["exp_avg"].shape == new.shape
    assert any(new is q for grp in fused.param_groups for q in grp["params"])
    assert set(fused.as_dict(None)) == set(a)
    sd = fused.state_dict()
    assert len(sd["param_groups"]) == len(a)


def test_no_cpu_path():
    a = _params(4)
    fused = FusedAdam([{"params": [p], "lr": 1e-3, "name": k} for k, p in a.items()])
    for p in a.values():
        p.grad = torch.ones_like(p)
    with pytest.raises(Exception, match="no CPU path"):
        fused.step()


def test_too_many_tensors():
    ps = [{"params": [torch.nn.Parameter(torch.zeros(2))]} for _ in range(17)]
    with pytest.raises(ValueError):
        FusedAdam(ps)


def _peer_gather_mirror(rows, touched, width):
    """numpy mirror of csrc/adam.cu adam_reduce_kernel's gather: `rows[k]` = rank k's flat gradient segment (rows of
    untouched Gaussians are exactly zero), `touched[k]` = its per-Gaussian flags.  Per float4 i the union of the masks of
    the Gaussians its four elements belong to selects the ranks to read; the sum runs in rank order."""
    import numpy as np

    world, n = len(rows), rows[0].size
    mask = np.zeros(touched[0].size, dtype=np.uint32)
    for k in range(world):
        mask |= (touched[k] != 0).astype(np.uint32) << k
    out = np.zeros(n, dtype=np.float32)
    for i in range(n // 4):
        e = 4 * i
        mk = mask[e // width] | mask[(e + 1) // width] | mask[(e + 2) // width] | mask[(e + 3) // width]
        g = np.zeros(4, dtype=np.float32)
        for k in range(world):
            if (mk >> k) & 1:
                g = g + rows[k][e:e + 4]
        out[e:e + 4] = g
    for j in range(n // 4 * 4, n):
        for k in range(world):
            if (mask[j // width] >> k) & 1:
                out[j] += rows[k][j]
    return out


@pytest.mark.parametrize("width", [1, 3, 4, 45])
def test_masked_peer_gather_equals_the_dense_sum(width):
    """The exchange of dnr_adam_step_reduce reads only the rows of ranks that touched a Gaussian; because untouched rows are
    zero that must equal the dense sum over ranks bit for bit — also where one float4 spans several Gaussians (width 1: four
    of them; width 3 / 45: rows straddle float4 boundaries)."""
    import numpy as np

    rng = np.random.default_rng(width)
    world, n_gauss = 3, 37
    touched = [(rng.random(n_gauss) < 0.35).astype(np.uint8) for _ in range(world)]
    rows = [(rng.standard_normal((n_gauss, width)).astype(np.float32) * touched[k][:, None]).reshape(-1) for k in range(world)]
    dense = np.zeros(n_gauss * width, dtype=np.float32)
    for k in range(world):
        dense = dense + rows[k]
    got = _peer_gather_mirror(rows, touched, width)
    assert np.array_equal(got, dense)
    # the per-Gaussian mask alone (the bug this guards against) would drop neighbours' rows inside a shared float4
    if width == 1:
        assert any(touched[k][g] and not touched[k][g - g % 4] for k in range(world) for g in range(n_gauss))
