"""CPU tests of the densification logic: statistics accumulation, split / duplicate / cull decisions
with the reference's schedule (dn_model.py:271-386), and Adam-state surgery."""
import torch

from dn_splatter_b200.densify import DensifyConfig, DensifyState, build_optimizers, exponential_lr, refinement_after


class _Model:
    def __init__(self, n, k=16):
        g = torch.Generator().manual_seed(0)
        P = torch.nn.Parameter
        self.gauss_params = torch.nn.ParameterDict({
            "means": P(torch.randn(n, 3, generator=g)), "scales": P(torch.full((n, 3), -6.0)),
            "quats": P(torch.randn(n, 4, generator=g)), "features_dc": P(torch.rand(n, 3, generator=g)),
            "features_rest": P(torch.zeros(n, k - 1, 3)), "opacities": P(torch.full((n, 1), 2.0)),
            "normals": P(torch.randn(n, 3, generator=g)),
        })
        self.last_size = (100, 200)
        self._bucket = None

    @property
    def num_points(self):
        return self.gauss_params["means"].shape[0]


def _adam_step(model, opts):
    for name, o in opts.items():
        model.gauss_params[name].grad = torch.ones_like(model.gauss_params[name])
        o.step()


def test_after_train_accumulates_only_visible():
    st = DensifyState()
    radii = torch.tensor([0, 5, 10, 0], dtype=torch.int32)
    absgrad = torch.tensor([[9.0, 9.0], [3.0, 4.0], [0.0, 1.0], [1.0, 1.0]])
    st.after_train(absgrad, radii, (100, 200))
    st.after_train(absgrad, radii, (100, 200))
    assert st.xys_grad_norm.tolist() == [0.0, 10.0, 2.0, 0.0]
    assert st.vis_counts.tolist() == [1.0, 3.0, 3.0, 1.0]
    assert torch.allclose(st.max_2Dsize, torch.tensor([0.0, 5 / 200, 10 / 200, 0.0]))


def test_split_dup_cull_and_adam_state():
    n = 10
    m = _Model(n)
    cfg = DensifyConfig()
    opts = build_optimizers(m)
    assert set(opts) == {"means", "scales", "quats", "features_dc", "features_rest", "opacities", "normals"}
    _adam_step(m, opts)
    with torch.no_grad():
        m.gauss_params["scales"][0:3] = -2.0      # exp(-2)=0.135 > densify_size_thresh -> split candidates
        m.gauss_params["opacities"][9] = -5.0     # sigmoid < 0.1 -> culled
    st = DensifyState()
    st.xys_grad_norm = torch.tensor([1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0])  # high grads: 0,1,3,4
    st.vis_counts = torch.ones(n)
    st.max_2Dsize = torch.zeros(n)
    means_before = m.gauss_params["means"].detach().clone()
    exp_avg_before = opts["means"].state[m.gauss_params["means"]]["exp_avg"].clone()
    step = 700  # > warmup, 700 % 3000 = 700 > num_train_data + refine_every
    info = refinement_after(m, opts, step, st, cfg, num_train_data=50, generator=torch.Generator().manual_seed(1))
    # splits: 0,1 (big & high grad) -> 2 children each, parents removed; dups: 3,4; culled: 9 (+ parents 0,1)
    assert info["split"] == 2 and info["dup"] == 2 and info["culled"] == 3
    assert m.num_points == n - 3 + 4 + 2 == 13
    for name, p in m.gauss_params.items():
        assert p.shape[0] == 13, name
        stt = opts[name].state[p]
        assert stt["exp_avg"].shape == p.shape and stt["exp_avg_sq"].shape == p.shape
        assert opts[name].param_groups[0]["params"][0] is p
    # survivors keep their values and moments (original indices 2..8 come first), new Gaussians start with zero moments
    assert torch.equal(m.gauss_params["means"][:7].detach(), means_before[2:9])
    assert torch.equal(opts["means"].state[m.gauss_params["means"]]["exp_avg"][:7], exp_avg_before[2:9])
    assert float(opts["means"].state[m.gauss_params["means"]]["exp_avg"][7:].abs().max()) == 0.0
    # children of a split are scaled down by 1.6 and the duplicates are exact copies
    assert torch.allclose(m.gauss_params["scales"][7:11].detach(), torch.full((4, 3), -2.0 - torch.log(torch.tensor(1.6)).item()))
    assert torch.equal(m.gauss_params["means"][11:13].detach(), means_before[3:5])
    assert st.xys_grad_norm is None  # statistics restart after every refinement
    # the optimizers still work on the re-created parameters
    _adam_step(m, opts)


def test_schedule_gates():
    m = _Model(6)
    cfg = DensifyConfig()
    opts = build_optimizers(m)
    st = DensifyState()
    assert refinement_after(m, opts, 400, st, cfg, 50)["n"] == 6  # warm-up: nothing happens
    # step % reset_interval <= num_train_data + refine_every: no densification, but opacity reset at == refine_every
    st.xys_grad_norm, st.vis_counts, st.max_2Dsize = torch.zeros(6), torch.ones(6), torch.zeros(6)
    _adam_step(m, opts)
    refinement_after(m, opts, 3100, st, cfg, 50)
    cap = torch.logit(torch.tensor(0.2)).item()
    assert float(m.gauss_params["opacities"].max()) <= cap + 1e-6
    assert float(opts["opacities"].state[m.gauss_params["opacities"]]["exp_avg"].abs().max()) == 0.0
    # after stop_split_at only culling continues
    with torch.no_grad():
        m.gauss_params["opacities"][0] = -6.0
    info = refinement_after(m, opts, 15100, st, cfg, 50)
    assert info["culled"] == 1 and m.num_points == 5


def test_exponential_lr_endpoints():
    assert abs(exponential_lr(1.6e-4, 1.6e-6, 0, 30000) - 1.6e-4) < 1e-10
    assert abs(exponential_lr(1.6e-4, 1.6e-6, 30000, 30000) - 1.6e-6) < 1e-12
    assert abs(exponential_lr(1.6e-4, 1.6e-6, 15000, 30000) - 1.6e-5) < 1e-9
