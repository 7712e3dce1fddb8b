"""Generates tests/golden/dn_refine_*.npz by executing the REFERENCE's own `DNSplatterModel.refinement_after`
(/root/reference/dn_splatter/dn_model.py:271-386, unmodified) on small parameter sets.  The schedule (warm-up, the
do_densification window, the post-densification cull, the opacity reset), the split / dup masks and their ORDER (the dup
mask is taken after split_gaussians has already shrunk the parents), the concatenation order, the cull-after-split and
the optimizer-state handling calls are reference code; the helpers it inherits from nerfstudio 1.1.3's SplatfactoModel
(split_gaussians, dup_gaussians, cull_gaussians, dup_in_all_optim, remove_from_all_optim) are absent from this
container and restated below [EXT].  tests/test_densify_golden.py checks dn_splatter_b200.densify against the files.

Run only where /root/reference exists:   python tests/golden/make_golden_refine.py
"""
import os
import sys
import types

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import make_golden_model as G0  # noqa: E402  (stubs + the restated nerfstudio base class)
from oracle import gsplat_ref as G  # noqa: E402

NAMES = ("means", "scales", "quats", "features_dc", "features_rest", "opacities", "normals")


class Parent(G0.SplatfactoModel):
    """nerfstudio 1.1.3 SplatfactoModel densification helpers [EXT], restated."""

    num_points = property(lambda self: self.means.shape[0])

    def split_gaussians(self, split_mask, samps):
        n_splits = int(split_mask.sum().item())
        centered_samples = torch.randn((samps * n_splits, 3))
        scaled_samples = torch.exp(self.scales[split_mask].repeat(samps, 1)) * centered_samples
        quats = self.quats[split_mask] / self.quats[split_mask].norm(dim=-1, keepdim=True)
        rots = G.quat_to_rotmat(quats.repeat(samps, 1))
        rotated_samples = torch.bmm(rots, scaled_samples[..., None]).squeeze(-1)
        new_means = rotated_samples + self.means[split_mask].repeat(samps, 1)
        size_fac = 1.6
        new_scales = torch.log(torch.exp(self.scales[split_mask]) / size_fac).repeat(samps, 1)
        self.scales[split_mask] = torch.log(torch.exp(self.scales[split_mask]) / size_fac)
        out = {"means": new_means, "features_dc": self.features_dc[split_mask].repeat(samps, 1),
               "features_rest": self.features_rest[split_mask].repeat(samps, 1, 1),
               "opacities": self.opacities[split_mask].repeat(samps, 1), "scales": new_scales,
               "quats": self.quats[split_mask].repeat(samps, 1)}
        for name, param in self.gauss_params.items():
            if name not in out:
                out[name] = param[split_mask].repeat(samps, 1)
        return out

    def dup_gaussians(self, dup_mask):
        return {name: param[dup_mask] for name, param in self.gauss_params.items()}

    def cull_gaussians(self, extra_cull_mask=None):
        culls = (torch.sigmoid(self.opacities) < self.config.cull_alpha_thresh).squeeze()
        if extra_cull_mask is not None:
            culls = culls | extra_cull_mask
        if self.step > self.config.refine_every * self.config.reset_alpha_every:
            toobigs = (torch.exp(self.scales).max(dim=-1).values > self.config.cull_scale_thresh).squeeze()
            if self.step < self.config.stop_screen_size_at and self.max_2Dsize is not None:
                toobigs = toobigs | (self.max_2Dsize > self.config.cull_screen_size).squeeze()
            culls = culls | toobigs
        for name, param in self.gauss_params.items():
            self.gauss_params[name] = torch.nn.Parameter(param[~culls])
        return culls

    def get_gaussian_param_groups(self):
        return {name: [self.gauss_params[name]] for name in NAMES}

    def remove_from_all_optim(self, optimizers, deleted_mask):
        for group, new_params in self.get_gaussian_param_groups().items():
            optimizer = optimizers.optimizers[group]
            param = optimizer.param_groups[0]["params"][0]
            state = optimizer.state[param]
            del optimizer.state[param]
            if "exp_avg" in state:
                state["exp_avg"] = state["exp_avg"][~deleted_mask]
                state["exp_avg_sq"] = state["exp_avg_sq"][~deleted_mask]
            optimizer.param_groups[0]["params"] = new_params
            optimizer.state[new_params[0]] = state

    def dup_in_all_optim(self, optimizers, dup_mask, n):
        for group, new_params in self.get_gaussian_param_groups().items():
            optimizer = optimizers.optimizers[group]
            param = optimizer.param_groups[0]["params"][0]
            state = optimizer.state[param]
            if "exp_avg" in state:
                rep = (n,) + tuple(1 for _ in range(state["exp_avg"].dim() - 1))
                for key in ("exp_avg", "exp_avg_sq"):
                    state[key] = torch.cat([state[key], torch.zeros_like(state[key][dup_mask.squeeze()]).repeat(*rep)], dim=0)
            del optimizer.state[param]
            optimizer.state[new_params[0]] = state
            optimizer.param_groups[0]["params"] = new_params


@G0.dataclasses.dataclass
class ParentConfig(G0.SplatfactoModelConfig):
    stop_split_at: int = 15000
    stop_screen_size_at: int = 4000
    densify_grad_thresh: float = 0.0008
    densify_size_thresh: float = 0.01
    n_split_samples: int = 2
    split_screen_size: float = 0.05
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    cull_screen_size: float = 0.15
    continue_cull_post_densification: bool = True


def scene(n, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    p = {"means": torch.randn(n, 3, generator=g), "quats": torch.randn(n, 4, generator=g),
         # log-scales straddling densify_size_thresh = 0.01 and 0.01 * 1.6; a few huge ones (cull_scale_thresh = 0.5)
         "scales": torch.log(0.004 + 0.02 * r(n, 3) ** 2 + (r(n, 1) > 0.96).float() * 0.8),
         "features_dc": r(n, 3), "features_rest": r(n, 15, 3) * 0.1, "opacities": torch.logit(0.02 + 0.96 * r(n, 1)),
         "normals": torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)}
    stats = {"xys_grad_norm": r(n) * 0.02, "vis_counts": 1.0 + torch.floor(r(n) * 20), "max_2Dsize": r(n) ** 3 * 0.3}
    return p, stats


def main():
    G0.install()
    sf = sys.modules["nerfstudio.models.splatfacto"]
    sf.SplatfactoModel, sf.SplatfactoModelConfig = Parent, ParentConfig
    sys.modules.pop("dn_splatter.dn_model", None)  # re-import on top of the densification-capable parent
    import dn_splatter.dn_model as M

    cases = {"warmup": 400, "densify": 700, "densify_late": 5200, "reset": 3100, "post_cull": 15100, "idle": 3000}
    for tag, step in cases.items():
        n = 240
        params, stats = scene(n, seed=step)
        m = M.DNSplatterModel.__new__(M.DNSplatterModel)
        torch.nn.Module.__init__(m)
        m.config = M.DNSplatterModelConfig()
        m.step, m.num_train_data, m.last_size = step, 50, (480, 640)
        m.gauss_params = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
        m.xys_grad_norm, m.vis_counts, m.max_2Dsize = (stats[k].clone() for k in ("xys_grad_norm", "vis_counts", "max_2Dsize"))
        opts = {}
        g = torch.Generator().manual_seed(step + 1)
        for k in NAMES:
            o = torch.optim.Adam([m.gauss_params[k]], lr=1e-3, eps=1e-15)
            o.state[m.gauss_params[k]] = {"step": torch.tensor(5.0), "exp_avg": torch.randn(params[k].shape, generator=g),
                                          "exp_avg_sq": torch.rand(params[k].shape, generator=g)}
            opts[k] = o
        z = {"in_" + k: v.numpy() for k, v in params.items()}
        z.update({"in_" + k: v.numpy() for k, v in stats.items()})
        z.update({"in_exp_avg_" + k: opts[k].state[m.gauss_params[k]]["exp_avg"].numpy().copy() for k in NAMES})
        z.update({"in_exp_avg_sq_" + k: opts[k].state[m.gauss_params[k]]["exp_avg_sq"].numpy().copy() for k in NAMES})
        torch.manual_seed(1234)  # the split samples
        M.DNSplatterModel.refinement_after(m, types.SimpleNamespace(optimizers=opts), step)
        for k in NAMES:
            p = m.gauss_params[k]
            z["out_" + k] = p.detach().numpy()
            st = opts[k].state[opts[k].param_groups[0]["params"][0]]
            assert opts[k].param_groups[0]["params"][0] is p or step <= 500 or tag in ("idle", "reset")
            z["out_exp_avg_" + k], z["out_exp_avg_sq_" + k] = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
        z["step"] = np.array(step)
        z["stats_reset"] = np.array(m.xys_grad_norm is None and m.vis_counts is None and m.max_2Dsize is None)
        np.savez_compressed(os.path.join(OUT, f"dn_refine_{tag}.npz"), **z)
        print(tag, step, "n:", n, "->", m.gauss_params["means"].shape[0], "stats reset:", bool(z["stats_reset"]))


if __name__ == "__main__":
    main()
