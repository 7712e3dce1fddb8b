"""densify.refinement_after against goldens produced by the REFERENCE's own DNSplatterModel.refinement_after
(tests/golden/make_golden_refine.py): same surviving Gaussians in the same order, same parameters, same Adam moments,
for the warm-up, densify (early / after stop_screen_size_at), opacity-reset, post-densification-cull and idle steps."""
import glob
import os
import types

import numpy as np
import pytest
import torch

from dn_splatter_b200.densify import PARAM_NAMES, DensifyConfig, DensifyState, refinement_after

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dn_refine_*.npz")))


class _Model:
    _bucket = None

    def __init__(self, params):
        self.gauss_params = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
        self.last_size = (480, 640)

    num_points = property(lambda self: self.gauss_params["means"].shape[0])


def test_goldens_exist():
    assert len(FILES) == 6


@pytest.mark.parametrize("f", FILES, ids=[os.path.basename(f) for f in FILES])
def test_refinement_matches_reference(f):
    z = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(f).items()}
    step = int(z["step"])
    m = _Model({k: z["in_" + k] for k in PARAM_NAMES})
    opts = {}
    for k in PARAM_NAMES:
        o = torch.optim.Adam([m.gauss_params[k]], lr=1e-3, eps=1e-15)
        o.state[m.gauss_params[k]] = {"step": torch.tensor(5.0), "exp_avg": z["in_exp_avg_" + k].clone(),
                                      "exp_avg_sq": z["in_exp_avg_sq_" + k].clone()}
        opts[k] = o
    st = DensifyState()
    st.xys_grad_norm, st.vis_counts, st.max_2Dsize = z["in_xys_grad_norm"].clone(), z["in_vis_counts"].clone(), z["in_max_2Dsize"].clone()
    # the reference's dn-splatter config: warmup_length 500, stop_split_at 15000, continue_cull_post_densification True
    cfg = DensifyConfig()
    refinement_after(m, opts, step, st, cfg, num_train_data=50, generator=torch.Generator().manual_seed(1234))
    for k in PARAM_NAMES:
        got, want = m.gauss_params[k].detach(), z["out_" + k]
        assert got.shape == want.shape, (k, got.shape, want.shape)
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7, msg=lambda s: f"{k}: {s}")
        state = opts[k].state[opts[k].param_groups[0]["params"][0]]
        torch.testing.assert_close(state["exp_avg"], z["out_exp_avg_" + k], rtol=0, atol=0)
        torch.testing.assert_close(state["exp_avg_sq"], z["out_exp_avg_sq_" + k], rtol=0, atol=0)
    assert bool(z["stats_reset"]) == (st.xys_grad_norm is None)
