"""Densification / culling around the hot path: the statistics the rasterizer feeds
(`means2d.absgrad`, `radii`) and `DNSplatterModel.refinement_after` of the reference
(/root/reference/dn_splatter/dn_model.py:271-386), whose helpers live in nerfstudio's SplatfactoModel [EXT 1.1.3:
after_train, split_gaussians, dup_gaussians, cull_gaussians, dup_in_all_optim, remove_from_all_optim].

Everything here is device-agnostic torch on whole tensors (no host loops, one boolean-count sync per refinement
step, which runs every `refine_every` = 100 steps), so it is unit-tested on the CPU (tests/test_densify_cpu.py).
Multi-GPU: all ranks must apply identical decisions — reduce the statistics with
`parallel.all_reduce_densification_stats` first and seed `torch.Generator` identically (the split samples are the only
random draw).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import Tensor

PARAM_NAMES = ("means", "scales", "quats", "features_dc", "features_rest", "opacities", "normals")


@dataclass
class DensifyConfig:
    """SplatfactoModelConfig fields [EXT nerfstudio 1.1.3 defaults] read by refinement_after; dn-splatter overrides
    warmup_length / stop_split_at (dn_model.py:102,112) and, for dn-splatter-big, cull_alpha_thresh and
    continue_cull_post_densification (dn_config.py:151-152)."""

    warmup_length: int = 500
    refine_every: int = 100
    reset_alpha_every: int = 30
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
    split_size_factor: float = 1.6


class DensifyState:
    """Running statistics between two refinements (SplatfactoModel.after_train [EXT])."""

    def __init__(self):
        self.xys_grad_norm: Optional[Tensor] = None
        self.vis_counts: Optional[Tensor] = None
        self.max_2Dsize: Optional[Tensor] = None

    def reset(self):
        self.xys_grad_norm = self.vis_counts = self.max_2Dsize = None

    @torch.no_grad()
    def all_reduce_(self, group=None) -> None:
        """Multi-GPU: every rank saw different views; before a refinement the statistics become those of ALL views
        (sum of the gradient norms and visibility counts, max of the screen sizes) so that every rank takes the same
        split / dup / cull decisions.  vis_counts starts at one on each rank: the extra ones are removed again."""
        import torch.distributed as dist

        from .parallel import all_reduce_densification_stats

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or self.vis_counts is None:
            return
        all_reduce_densification_stats(self.xys_grad_norm, self.vis_counts, self.max_2Dsize, group)
        self.vis_counts -= float(dist.get_world_size(group) - 1)

    def all_reduce_before_refinement(self, step: int, warmup_length: int, group=None) -> bool:
        """What a multi-rank training loop calls at a refine_every boundary, right before refinement_after: reduces ONLY
        when that call will consume and reset the statistics (step > warmup_length).  During warm-up refinement_after
        returns without resetting; an in-place reduction there would be reduced again at the next boundary and the early
        steps would end up weighted by world_size^k."""
        if step <= warmup_length:
            return False
        self.all_reduce_(group)
        return True

    @torch.no_grad()
    def after_train(self, absgrad: Tensor, radii: Tensor, last_size) -> None:
        """absgrad [N,2] (means2d.absgrad of the view just trained), radii [N] int32, last_size (H, W).  Mask-free
        formulation (no boolean indexing, hence no sync): invisible Gaussians add zero."""
        vis = radii > 0
        n = radii.shape[0]
        if self.xys_grad_norm is None or self.xys_grad_norm.shape[0] != n:
            self.xys_grad_norm = torch.zeros(n, device=radii.device, dtype=torch.float32)
            self.vis_counts = torch.ones(n, device=radii.device, dtype=torch.float32)
            self.max_2Dsize = torch.zeros(n, device=radii.device, dtype=torch.float32)
        zero = torch.zeros((), device=radii.device, dtype=torch.float32)
        self.vis_counts += vis.to(torch.float32)
        self.xys_grad_norm += torch.where(vis, absgrad.norm(dim=-1), zero)  # where(): never touches invisible rows
        rel = radii.to(torch.float32) / float(max(last_size[0], last_size[1]))
        self.max_2Dsize = torch.maximum(self.max_2Dsize, torch.where(vis, rel, zero))


def _quat_to_rotmat(q: Tensor) -> Tensor:
    w, x, y, z = torch.unbind(torch.nn.functional.normalize(q, dim=-1), dim=-1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1),
    ], dim=-2)


@torch.no_grad()
def split_gaussians(params: Dict[str, Tensor], mask: Tensor, samps: int, size_fac: float = 1.6,
                    generator: Optional[torch.Generator] = None) -> Dict[str, Tensor]:
    """`samps` children per selected Gaussian, sampled from the parent's own distribution; parents and children get
    their scale divided by `size_fac` (the parents are culled afterwards by the caller)."""
    n = int(mask.sum())
    dev = params["means"].device
    noise = torch.randn((samps * n, 3), generator=generator, device="cpu" if generator is not None else dev).to(dev)
    scaled = torch.exp(params["scales"][mask].repeat(samps, 1)) * noise
    rots = _quat_to_rotmat(params["quats"][mask].repeat(samps, 1))
    out = {k: v[mask].repeat(samps, *([1] * (v.dim() - 1))) for k, v in params.items()}
    out["means"] = torch.bmm(rots, scaled[..., None]).squeeze(-1) + params["means"][mask].repeat(samps, 1)
    shrunk = torch.log(torch.exp(params["scales"][mask]) / size_fac)
    out["scales"] = shrunk.repeat(samps, 1)
    params["scales"][mask] = shrunk
    return out


@torch.no_grad()
def dup_gaussians(params: Dict[str, Tensor], mask: Tensor) -> Dict[str, Tensor]:
    return {k: v[mask].clone() for k, v in params.items()}


def _resize_adam_state(optimizer: torch.optim.Optimizer, old: torch.nn.Parameter, new: torch.nn.Parameter, fn) -> None:
    """Moves the optimizer state of `old` to `new`, transforming exp_avg / exp_avg_sq with `fn`."""
    state = optimizer.state.pop(old, None)
    if state is not None:
        for key in ("exp_avg", "exp_avg_sq"):
            if key in state:
                state[key] = fn(state[key])
        optimizer.state[new] = state
    for group in optimizer.param_groups:
        group["params"] = [new if p is old else p for p in group["params"]]


@torch.no_grad()
def refinement_after(model, optimizers: Dict[str, torch.optim.Optimizer], step: int, state: DensifyState,
                     cfg: DensifyConfig, num_train_data: int, generator: Optional[torch.Generator] = None) -> Dict[str, int]:
    """Split / duplicate / cull / opacity reset with the reference's schedule; returns counts for logging.
    `optimizers`: one optimizer per gauss_params name (first param of the first group), as dn_config.py builds them."""
    info = {"split": 0, "dup": 0, "culled": 0, "n": model.num_points}
    if step <= cfg.warmup_length:
        return info
    gp = model.gauss_params
    reset_interval = cfg.reset_alpha_every * cfg.refine_every
    do_densify = step < cfg.stop_split_at and step % reset_interval > num_train_data + cfg.refine_every
    keep: Optional[Tensor] = None
    names = [k for k in PARAM_NAMES if k in gp]
    if do_densify:
        assert state.xys_grad_norm is not None and state.vis_counts is not None and state.max_2Dsize is not None
        H, W = model.last_size
        avg = (state.xys_grad_norm / state.vis_counts) * 0.5 * max(H, W)
        high = avg > cfg.densify_grad_thresh
        big = gp["scales"].exp().max(dim=-1).values > cfg.densify_size_thresh
        splits = big.clone()
        if step < cfg.stop_screen_size_at:
            splits |= state.max_2Dsize > cfg.split_screen_size
        splits &= high
        data = {k: gp[k].data for k in names}
        new_split = split_gaussians(data, splits, cfg.n_split_samples, cfg.split_size_factor, generator)
        # the reference takes the dup mask AFTER split_gaussians has shrunk the parents in place (dn_model.py:309-316):
        # a split parent that falls below densify_size_thresh after the /1.6 is duplicated as well (the copy survives,
        # the parent itself is culled below) — pinned by tests/test_densify_golden.py
        dups = (data["scales"].exp().max(dim=-1).values <= cfg.densify_size_thresh) & high
        new_dup = dup_gaussians(data, dups)
        n_split, n_dup = int(splits.sum()), int(dups.sum())
        info["split"], info["dup"] = n_split, n_dup
        grown = {k: torch.cat([data[k], new_split[k], new_dup[k]], dim=0) for k in names}
        n_new = cfg.n_split_samples * n_split + n_dup
        state.max_2Dsize = torch.cat([state.max_2Dsize, torch.zeros(n_new, device=state.max_2Dsize.device)])
        # the split parents are pruned together with the low-opacity / oversized ones
        cull = torch.cat([splits, torch.zeros(n_new, dtype=torch.bool, device=splits.device)])
        cull |= _cull_mask(grown, state.max_2Dsize, step, cfg)
        keep = ~cull

        def grow_then_keep(t):
            pad = torch.zeros((n_new,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            return torch.cat([t, pad], dim=0)[keep]

        _replace_params(model, optimizers, {k: grown[k][keep] for k in names}, grow_then_keep)
        info["culled"] = int(cull.sum())
    elif step >= cfg.stop_split_at and cfg.continue_cull_post_densification:
        data = {k: gp[k].data for k in names}
        cull = _cull_mask(data, state.max_2Dsize, step, cfg)
        keep = ~cull
        _replace_params(model, optimizers, {k: data[k][keep] for k in names}, lambda t: t[keep])
        info["culled"] = int(cull.sum())
    if step < cfg.stop_split_at and step % reset_interval == cfg.refine_every:
        # reset opacity to twice the cull threshold and restart its Adam moments (dn_model.py:364-379)
        cap = torch.logit(torch.tensor(cfg.cull_alpha_thresh * 2.0)).item()
        model.gauss_params["opacities"].data.clamp_(max=cap)
        opt = optimizers.get("opacities")
        if opt is not None:
            st = opt.state.get(model.gauss_params["opacities"])
            if st:
                st["exp_avg"].zero_()
                st["exp_avg_sq"].zero_()
    state.reset()
    info["n"] = model.num_points
    return info


def _cull_mask(data: Dict[str, Tensor], max_2Dsize: Optional[Tensor], step: int, cfg: DensifyConfig) -> Tensor:
    """SplatfactoModel.cull_gaussians [EXT]: low opacity, and (after the first opacity reset) oversized in world or
    screen space."""
    culls = torch.sigmoid(data["opacities"]).squeeze(-1) < cfg.cull_alpha_thresh
    if step > cfg.refine_every * cfg.reset_alpha_every:
        toobig = data["scales"].exp().max(dim=-1).values > cfg.cull_scale_thresh
        if step < cfg.stop_screen_size_at and max_2Dsize is not None:
            toobig |= max_2Dsize > cfg.cull_screen_size
        culls |= toobig
    return culls


def _replace_params(model, optimizers: Dict[str, torch.optim.Optimizer], new_data: Dict[str, Tensor], state_fn) -> None:
    """Re-creates every Parameter (the Gaussian count changed) and carries the Adam moments over."""
    for name, t in new_data.items():
        old = model.gauss_params[name]
        new = torch.nn.Parameter(t.contiguous(), requires_grad=old.requires_grad)
        opt = optimizers.get(name)
        if opt is not None:
            _resize_adam_state(opt, old, new, state_fn)
        model.gauss_params[name] = new
    if getattr(model, "_bucket", None) is not None:
        # the flat gradient bucket must follow the new parameter set; a peer bucket is re-allocated in symmetric memory (a
        # collective: every rank refines at the same step with identical decisions, see Trainer)
        peer, group = getattr(model, "_bucket_mode", (False, None))
        model.enable_flat_grads(peer=peer, group=group)


def build_optimizers(model, groups: Optional[Dict[str, Dict]] = None) -> Dict[str, torch.optim.Optimizer]:
    """One Adam per gauss_params group with the reference's learning rates (dn_config.py:29-68)."""
    from .dn_config import optimizer_groups

    groups = groups or optimizer_groups()
    out = {}
    for name, p in model.gauss_params.items():
        if name in groups:
            g = groups[name]
            out[name] = torch.optim.Adam([p], lr=g["lr"], eps=g["eps"])
    return out


def exponential_lr(lr_init: float, lr_final: float, step: int, max_steps: int) -> float:
    """nerfstudio ExponentialDecayScheduler [EXT] without warm-up (means: 1.6e-4 -> 1.6e-6, dn_config.py:30-35)."""
    t = min(max(step / max_steps, 0.0), 1.0)
    return float(torch.tensor(lr_init).log().mul(1 - t).add(torch.tensor(lr_final).log().mul(t)).exp())
