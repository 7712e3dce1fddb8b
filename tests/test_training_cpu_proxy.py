"""Host-side training logic end to end on the CPU (oracle substituted for the CUDA entry points, tests/cpu_proxy.py):
model -> losses -> backward into the flat bucket -> per-group Adam -> after_train -> refinement.  The rasterizer's
gradients equal the oracle's (GPU parity tests), so this exercises exactly the code a GPU run executes on the host."""
import torch

from tests.cpu_proxy import cpu_proxy


def test_trainer_fits_a_small_scene_and_survives_refinement():
    from dn_splatter_b200.cameras import Cameras
    from dn_splatter_b200.dn_model import DNSplatterModelConfig
    from dn_splatter_b200.losses import DepthLossType
    from dn_splatter_b200.synthetic import make_scene, ring_cameras
    from dn_splatter_b200.trainer import Trainer

    torch.manual_seed(0)
    W, H, n_views = 48, 32, 3
    cams = [Cameras(c["c2w"][None], c["fx"], c["fy"], c["cx"], c["cy"], W, H, metadata={"cam_idx": i})
            for i, c in enumerate(ring_cameras(n_views, W, H))]
    cfg = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black", use_depth_loss=True, depth_lambda=0.2,
                                depth_loss_type=DepthLossType.EdgeAwareLogL1, ssim_lambda=0.0, warmup_length=25, refine_every=13,
                                densify_grad_thresh=1e-6, sh_degree_interval=1)
    with cpu_proxy():
        target = cfg.setup(device="cpu", num_train_data=n_views)
        target.load_gaussians(make_scene(80, seed=4))
        target.step = 100
        target.eval()
        batches = []
        with torch.no_grad():
            for c in cams:
                o = target.get_outputs(c)
                batches.append({"image": (o["rgb"] * 255).round().to(torch.uint8), "mono_depth": o["depth"].clone(),
                                "normal": o["normal"].clone()})
        model = cfg.setup(device="cpu", num_train_data=n_views)
        start = make_scene(80, seed=4)
        g = torch.Generator().manual_seed(9)
        start["features_dc"] = torch.rand(start["features_dc"].shape, generator=g)
        start["means"] = start["means"] + 0.03 * torch.randn(start["means"].shape, generator=g)
        model.load_gaussians(start)
        model.num_train_data = n_views
        tr = Trainer(model, lambda s: (cams[s % n_views], dict(batches[s % n_views])), max_steps=100)
        losses, counts, refines = [], [], []
        for _ in range(30):  # refinement hooks fire at steps 13 (inside the warm-up: no-op) and 26 (densifies)
            out = tr.train_iteration()
            assert torch.isfinite(out["loss"])
            losses.append(float(out["loss"]))
            counts.append(model.num_points)
            if out["refine"] is not None:
                refines.append(out["refine"])
    # phase 1 (no refinement yet): Adam with the reference's learning rates brings the loss down
    first, last = sum(losses[:6]) / 6, sum(losses[18:24]) / 6
    assert last < 0.9 * first, (first, last, losses)
    # phase 2: the densification at step 26 (threshold set absurdly low on purpose: everything splits) changes the Gaussian
    # count; the loop keeps running on the re-created parameters / optimizer state / gradient bucket
    assert [r["split"] + r["dup"] > 0 for r in refines] == [False, True], refines
    assert counts[-1] > counts[0]
    n = model.num_points
    for name, opt in tr.optimizers.items():
        p = model.gauss_params[name]
        assert p.shape[0] == n and opt.param_groups[0]["params"][0] is p
    assert n * 59 <= model._bucket.flat.numel() < n * 59 + 24  # segments padded to 16 B


def test_trainer_with_fused_adam_follows_the_per_group_adam_trainer(monkeypatch):
    """Trainer(fused_adam=True): one FusedAdam for all groups (its kernel launch replaced by FusedAdam.reference_step, the
    torch restatement of csrc/adam.cu) must walk the same trajectory as the default per-group torch.optim.Adam trainer,
    through a refinement that re-creates the parameters (learning-rate schedule by group name, moment surgery)."""
    from dn_splatter_b200.cameras import Cameras
    from dn_splatter_b200.dn_model import DNSplatterModelConfig
    from dn_splatter_b200.losses import DepthLossType
    from dn_splatter_b200.optim import FusedAdam
    from dn_splatter_b200.synthetic import make_scene, ring_cameras
    from dn_splatter_b200.trainer import Trainer

    monkeypatch.setattr(FusedAdam, "step", lambda self, closure=None: self.reference_step())
    W, H, n_views = 40, 32, 2
    cams = [Cameras(c["c2w"][None], c["fx"], c["fy"], c["cx"], c["cy"], W, H) for c in ring_cameras(n_views, W, H)]
    cfg = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black", use_depth_loss=True, depth_lambda=0.2,
                                depth_loss_type=DepthLossType.LogL1, ssim_lambda=0.0, warmup_length=3, refine_every=5,
                                densify_grad_thresh=1e-6, sh_degree_interval=1)
    g = torch.Generator().manual_seed(2)
    batches = [{"image": (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8), "mono_depth": 2 + 4 * torch.rand(H, W, 1, generator=g),
                "normal": torch.rand(H, W, 3, generator=g)} for _ in range(n_views)]
    results = []
    with cpu_proxy():
        for fused in (False, True):
            model = cfg.setup(device="cpu", num_train_data=n_views)
            model.load_gaussians(make_scene(50, seed=6))
            model.num_train_data = n_views
            tr = Trainer(model, lambda s: (cams[s % n_views], dict(batches[s % n_views])), max_steps=50, fused_adam=fused, seed=3)
            losses = [float(tr.train_iteration()["loss"]) for _ in range(12)]  # refinements at steps 5 and 10
            results.append((losses, {k: v.detach().clone() for k, v in model.gauss_params.items()}))
            if fused:
                assert len({id(o) for o in tr.optimizers.values()}) == 1  # one optimizer object behind every group name
                for grp in tr.fused.param_groups:
                    assert grp["params"][0] is model.gauss_params[grp["name"]]
    (la, pa), (lb, pb) = results
    assert pa["means"].shape == pb["means"].shape and pa["means"].shape[0] != 50  # both densified identically
    for x, y in zip(la, lb):
        assert abs(x - y) <= 1e-4 * max(1.0, abs(x)), (la, lb)
    for k in pa:
        assert float((pa[k] - pb[k]).abs().max()) <= 1e-3 * float(pb[k].abs().max() + 1e-12), k
