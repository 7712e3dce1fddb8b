"""world_size-2 gloo test (CPU) of the multi-GPU host logic: per-camera sharding + one flat all-reduce must
reproduce the single-process gradient over the union of the views (SURVEY.md §8e determinism check)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dn_splatter_b200.parallel import GRAD_PARAMS, FlatGradBucket, shard_views


def _fake_view_grads(params, view):
    """Stand-in for one view's backward: a deterministic function of (parameters, view id)."""
    out = {}
    for i, (k, p) in enumerate(params.items()):
        out[k] = torch.sin(p.detach() * (view + 1) + i)
    return out


def _make_params():
    g = torch.Generator().manual_seed(0)
    shapes = {"means": (50, 3), "scales": (50, 3), "quats": (50, 4), "features_dc": (50, 3), "features_rest": (50, 15, 3),
              "opacities": (50, 1), "normals": (50, 3)}
    return {k: torch.nn.Parameter(torch.randn(*s, generator=g)) for k, s in shapes.items()}


def _worker(rank, world, port, n_views, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = _make_params()
    bucket = FlatGradBucket(params)
    assert "normals" not in bucket.names and 50 * 59 <= bucket.flat.numel() < 50 * 59 + 24
    bucket.zero_()
    for v in shard_views(n_views, rank, world):
        for k, g in _fake_view_grads(params, v).items():
            if k in bucket.views:
                bucket.views[k] += g  # what the rasterizer's grad-sink / autograd accumulation does
    bucket.all_reduce()
    if rank == 0:
        ret.put({k: params[k].grad.numpy().copy() for k in bucket.names})  # numpy: pickled by value, no fd hand-off
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_rank():
    n_views, world = 5, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    params = _make_params()
    want = {k: torch.zeros_like(params[k]) for k in GRAD_PARAMS}
    for v in range(n_views):
        for k, g in _fake_view_grads(params, v).items():
            if k in want:
                want[k] += g
    for k in GRAD_PARAMS:
        torch.testing.assert_close(torch.from_numpy(got[k]), want[k], rtol=1e-6, atol=1e-6)
    assert sorted(shard_views(5, 0, 2) + shard_views(5, 1, 2)) == list(range(5))


def _stats_worker(rank, world, port, ret):
    from dn_splatter_b200.densify import DensifyState

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = DensifyState()
    for v in shard_views(6, rank, world):
        absgrad, radii = _fake_view_stats(v)
        st.after_train(absgrad, radii, (48, 64))
    st.all_reduce_()
    if rank == 1:
        ret.put(tuple(t.numpy().copy() for t in (st.xys_grad_norm, st.vis_counts, st.max_2Dsize)))
    dist.barrier()
    dist.destroy_process_group()


def _fake_view_stats(view):
    g = torch.Generator().manual_seed(100 + view)
    radii = (torch.rand(40, generator=g) * 6).int() * (torch.rand(40, generator=g) > 0.3).int()
    return torch.randn(40, 2, generator=g), radii


def test_two_rank_densification_statistics_equal_single_rank():
    from dn_splatter_b200.densify import DensifyState

    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    st = DensifyState()
    for v in range(6):
        st.after_train(*_fake_view_stats(v), (48, 64))
    got = [torch.from_numpy(a) for a in got]
    torch.testing.assert_close(got[0], st.xys_grad_norm, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(got[1], st.vis_counts)
    torch.testing.assert_close(got[2], st.max_2Dsize)


def _warmup_worker(rank, world, port, ret):
    """The Trainer's schedule around the warm-up boundary: after_train every step, a refine boundary every 2 steps,
    warm-up of 5 steps; refinement_after is modelled by what it does to the statistics (reset once step > warm-up)."""
    from dn_splatter_b200.densify import DensifyState

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st, warmup, consumed = DensifyState(), 5, []
    for step in range(9):
        absgrad, radii = _fake_view_stats(step * world + rank)
        st.after_train(absgrad, radii, (48, 64))
        if step > 0 and step % 2 == 0:
            if st.all_reduce_before_refinement(step, warmup):
                consumed.append((step, st.xys_grad_norm.numpy().copy(), st.vis_counts.numpy().copy(), st.max_2Dsize.numpy().copy()))
                st.reset()
    if rank == 0:
        ret.put(consumed)
    dist.barrier()
    dist.destroy_process_group()


def test_statistics_are_not_re_reduced_during_warmup():
    """reducing in place at every boundary of the warm-up (where nothing resets the statistics) weighted the
    first steps by world_size^k.  The statistics the first real refinement consumes must equal the single-process ones."""
    from dn_splatter_b200.densify import DensifyState

    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_warmup_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert [g[0] for g in got] == [6, 8]  # boundaries 2 and 4 lie inside the warm-up: no reduction there
    st = DensifyState()
    for step in range(7):  # steps 0..6 on both ranks feed the first refinement
        for r in range(world):
            st.after_train(*_fake_view_stats(step * world + r), (48, 64))
    torch.testing.assert_close(torch.from_numpy(got[0][1]), st.xys_grad_norm, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(torch.from_numpy(got[0][2]), st.vis_counts)
    torch.testing.assert_close(torch.from_numpy(got[0][3]), st.max_2Dsize)


def _trainer_worker(rank, world, port, ret):
    """Trainer(world_size=2) on the CPU proxy: per-camera sharding, flat all-reduce (gloo), statistics reduced before the
    refinement, identical split samples — the replicas must hold identical parameters after a densification."""
    from dn_splatter_b200.cameras import Cameras
    from dn_splatter_b200.dn_model import DNSplatterModelConfig
    from dn_splatter_b200.losses import DepthLossType
    from dn_splatter_b200.synthetic import make_scene, ring_cameras
    from dn_splatter_b200.trainer import Trainer
    from tests.cpu_proxy import cpu_proxy

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    W, H, n_views = 40, 32, 4
    cams = [Cameras(c["c2w"][None], c["fx"], c["fy"], c["cx"], c["cy"], W, H, metadata={"cam_idx": i})
            for i, c in enumerate(ring_cameras(n_views, W, H))]
    g = torch.Generator().manual_seed(3)
    batches = [{"image": (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8),
                "mono_depth": 2 + 6 * torch.rand(H, W, 1, generator=g),
                "normal": torch.rand(H, W, 3, generator=g)} for _ in range(n_views)]
    cfg = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black", use_depth_loss=True, depth_lambda=0.2,
                                depth_loss_type=DepthLossType.LogL1, ssim_lambda=0.0, warmup_length=3, refine_every=3,
                                densify_grad_thresh=1e-6, sh_degree_interval=1)
    with cpu_proxy():
        m = cfg.setup(device="cpu", num_train_data=2)
        m.load_gaussians(make_scene(60, seed=4))
        m.num_train_data = 2

        def next_train(step):  # rank r renders views {i : i mod world == r}
            v = (step * world + rank) % n_views
            return cams[v], dict(batches[v])

        tr = Trainer(m, next_train, max_steps=100, world_size=world)
        counts = []
        for _ in range(8):  # boundary 6: past the warm-up and step % reset > num_train_data + refine_every -> densifies
            out = tr.train_iteration()
            assert torch.isfinite(out["loss"])
            counts.append(m.num_points)
    names = sorted(k for k in m.gauss_params if k != "normals")
    mine = torch.cat([m.gauss_params[k].detach().reshape(-1) for k in names])
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()]))
    same = len({int(s) for s in sizes}) == 1
    if same:
        theirs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(theirs, mine)
        same = all(torch.equal(t, mine) for t in theirs)
    if rank == 0:
        ret.put((counts, same))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_trainer_keeps_replicas_identical_through_a_refinement():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_trainer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, same = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(set(counts)) > 1, counts
    assert same, "replicas diverged"
