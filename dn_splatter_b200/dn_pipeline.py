"""DNSplatterPipeline / DNSplatterPipelineConfig with the reference's constructor contract
(/root/reference/dn_splatter/dn_pipeline.py:49-130): builds the datamanager, forwards seed points
(points3D_xyz / points3D_rgb / points3D_normals metadata) to the model, and — where the reference wraps the
model in DDP — installs the per-camera sharding + single flat all-reduce of parallel.py.
Eval loops, point-cloud metrics and render dumps (dn_pipeline.py:133-638) are I/O around the hot path and
out of scope."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Literal, Optional, Type

import torch
import torch.distributed as dist

from .dn_model import DNSplatterModel, DNSplatterModelConfig
from .parallel import FlatGradBucket

try:  # with nerfstudio the classes extend VanillaPipeline(Config), exactly as the reference's do (dn_pipeline.py:36-66)
    from nerfstudio.pipelines.base_pipeline import VanillaPipeline as _PipelineBase  # type: ignore
    from nerfstudio.pipelines.base_pipeline import VanillaPipelineConfig as _PipelineConfigBase  # type: ignore

    HAVE_NERFSTUDIO = True
except Exception:  # noqa: BLE001
    HAVE_NERFSTUDIO = False
    _PipelineBase = torch.nn.Module

    @dataclass
    class _PipelineConfigBase:
        pass


@dataclass
class DNSplatterPipelineConfig(_PipelineConfigBase):
    _target: Type = field(default_factory=lambda: DNSplatterPipeline)
    datamanager: Any = None  # object with .setup(device=, test_mode=, world_size=, local_rank=) or a ready datamanager
    model: DNSplatterModelConfig = field(default_factory=DNSplatterModelConfig)
    experiment_name: str = "experiment"
    skip_point_metrics: bool = True
    num_pd_points: int = 1_000_000
    save_train_images: bool = False

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class DNSplatterPipeline(_PipelineBase):
    """Datamanager contract: `next_train(step) -> (camera, batch)` with the batch keys of the reference's
    DNSplatterDataManager (image, mask, sensor_depth, mono_depth, normal, confidence; dn_datamanager.py:90-150),
    optional `train_dataparser_outputs.metadata`, `train_dataset` (len + optional scene_box/metadata)."""

    def __init__(self, config: DNSplatterPipelineConfig, device: str,
                 test_mode: Literal["test", "val", "inference"] = "val", world_size: int = 1, local_rank: int = 0,
                 grad_scaler=None):
        # like the reference (dn_pipeline.py:76): skip VanillaPipeline.__init__, which would build its own datamanager / DDP
        (super(_PipelineBase, self) if HAVE_NERFSTUDIO else super()).__init__()
        self.config, self.test_mode = config, test_mode
        dm = config.datamanager
        if hasattr(dm, "setup"):
            dm = dm.setup(device=device, test_mode=test_mode, world_size=world_size, local_rank=local_rank)
        self.datamanager = dm
        seed_pts = None
        meta = getattr(getattr(dm, "train_dataparser_outputs", None), "metadata", None) or {}
        if "points3D_xyz" in meta:
            seed_pts = (meta["points3D_xyz"], meta["points3D_rgb"])
            if "points3D_normals" in meta:
                seed_pts = seed_pts + (meta["points3D_normals"],)
        train_ds = getattr(dm, "train_dataset", None)
        assert train_ds is not None, "Missing input dataset"
        self._model = config.model.setup(num_train_data=len(train_ds), device=device, seed_points=seed_pts,
                                         metadata=getattr(train_ds, "metadata", None),
                                         scene_box=getattr(train_ds, "scene_box", None), grad_scaler=grad_scaler)
        self.world_size, self.local_rank = world_size, local_rank
        self.bucket: Optional[FlatGradBucket] = None
        if world_size > 1:  # reference :123-128 wraps in DDP + barrier
            self.bucket = FlatGradBucket(dict(self._model.gauss_params))
            if dist.is_initialized():
                dist.barrier()

    @property
    def model(self) -> DNSplatterModel:
        return self._model

    @property
    def device(self):
        return self._model.device

    def get_train_loss_dict(self, step: int):
        """VanillaPipeline.get_train_loss_dict [EXT]: one view per rank; call `reduce_gradients()` after backward."""
        camera, batch = self.datamanager.next_train(step)
        outputs = self._model(camera)
        loss_dict = self._model.get_loss_dict(outputs, batch, None)
        return outputs, loss_dict, {}

    def reduce_gradients(self):
        if self.bucket is not None:
            self.bucket.all_reduce()
