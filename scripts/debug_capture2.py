"""2-rank debug: which operation invalidates the CUDA-graph capture of the training step when the gradients live in a
PeerGradBucket?  torchrun --nproc-per-node 2 scripts/debug_capture2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.utils._python_dispatch import TorchDispatchMode  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
dist.init_process_group("nccl", device_id=dev)
from dn_splatter_b200.cameras import Cameras  # noqa: E402
from dn_splatter_b200.dn_model import DNSplatterModelConfig  # noqa: E402
from dn_splatter_b200.graph_step import GraphedTrainStep  # noqa: E402
from dn_splatter_b200.losses import DepthLossType  # noqa: E402
from dn_splatter_b200.optim import FusedAdam  # noqa: E402
from dn_splatter_b200.synthetic import make_scene, ring_cameras  # noqa: E402

W, H, n = 640, 368, 200_000
cams = [Cameras(c["c2w"][None], c["fx"], c["fy"], c["cx"], c["cy"], W, H, metadata={"cam_idx": i}) for i, c in enumerate(ring_cameras(8, W, H))]
g = torch.Generator().manual_seed(3)
batch = {"image": (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8).to(dev), "mono_depth": (2 + 6 * torch.rand(H, W, 1, generator=g)).to(dev),
         "normal": (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8).to(dev)}


class Tracer(TorchDispatchMode):
    last = None

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        try:
            torch.cuda.is_current_stream_capturing()
        except Exception as exc:  # noqa: BLE001
            raise RuntimeError(f"capture invalid after torch op {func} (previous ok op: {Tracer.last}): {exc}") from exc
        Tracer.last = str(func)
        return out


def attempt(tag, peer, use_dense, tracer):
    cfg = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black", use_depth_loss=True, depth_lambda=0.2,
                                depth_loss_type=DepthLossType.EdgeAwareLogL1, ssim_lambda=0.2, sync_free=True)
    m = cfg.setup(device=dev)
    m.load_gaussians(make_scene(n, seed=5))
    m.step = 30000
    m.train()
    bucket = m.enable_flat_grads(peer=peer)
    if peer and not use_dense:
        bucket.dense = {}
    opt = FusedAdam.for_model(m)
    for s in range(4):
        bucket.zero_()
        o = m.get_outputs(cams[(s * world + rank) % 8])
        ld = m.get_loss_dict(o, dict(batch))
        (ld["main_loss"] + ld["scale_reg"]).backward()
        if peer:
            opt.step_reduce(bucket)
        else:
            bucket.all_reduce()
            opt.step()
    torch.cuda.synchronize()
    try:
        if tracer:
            with Tracer():
                gs = GraphedTrainStep(m, bucket, cams[rank], batch, n_slots=1, warmup=2)
        else:
            gs = GraphedTrainStep(m, bucket, cams[rank], batch, n_slots=1, warmup=2)
        gs(cams[rank], 0)
        torch.cuda.synchronize()
        res = "captured + replayed OK"
    except Exception as exc:  # noqa: BLE001
        res = f"FAILED: {type(exc).__name__}: {str(exc)[:400]}"
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
    print(f"[rank {rank}] {tag}: {res}", flush=True)
    dist.barrier()


os.environ["DNR_DEBUG_CAPTURE"] = "1"
for tag, peer, dense, tracer in (("nccl bucket", False, False, False), ("peer bucket, no dense sink", True, False, False),
                                 ("peer bucket, dense sink", True, True, False), ("peer bucket, dense sink, traced", True, True, True),
                                 ("peer bucket, no dense sink (again)", True, False, False)):
    attempt(tag, peer, dense, tracer)
dist.destroy_process_group()
