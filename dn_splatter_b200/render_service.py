"""Render-all-views forward service: what `gs-mesh`, `ns-eval` and `render_model.py` do with the model
— `for camera in views: model.get_outputs_for_camera(camera)` (reference export_mesh.py:360-367, 863-905, 965-1017;
dn_pipeline.py:194-214; scripts/render_model.py:47-69; utils/utils.py:331-443) — as one no-grad loop that keeps the
device busy: sync-free binning, camera by value, optional asynchronous copy of the selected maps to pinned host memory
(double-buffered) so the consumer (TSDF / Poisson / image writer) overlaps with rendering.
Round-1 status: eager launches, host logic covered by tests/test_render_service_cpu_proxy.py; throughput not yet measured."""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional, Sequence, Tuple

import torch
from torch import Tensor

DEFAULT_KEYS = ("rgb", "depth", "normal", "surface_normal", "accumulation")


class ViewRenderer:
    def __init__(self, model, keys: Sequence[str] = DEFAULT_KEYS, to_host: bool = False, n_host_buffers: int = 2):
        self.model, self.keys, self.to_host = model, tuple(keys), to_host
        self._host: list = [None] * n_host_buffers
        self._events: list = [None] * n_host_buffers

    @torch.no_grad()
    def render(self, cameras: Iterable) -> Iterator[Tuple[int, Dict[str, Tensor]]]:
        """Yields (view index, {key: map}) in order.  With to_host=True the maps are pinned host tensors whose copy has
        completed when they are yielded; the next view is already rendering while the caller consumes them.  The host
        buffers are reused round-robin: consume (or copy) a view's maps before asking for the next one."""
        m = self.model
        was_training = m.training
        m.eval()
        try:
            pending: Optional[Tuple[int, int]] = None
            for idx, cam in enumerate(cameras):
                out = m.get_outputs(cam)
                maps = {k: out[k] for k in self.keys if k in out}
                if not self.to_host:
                    yield idx, maps
                    continue
                slot = idx % len(self._host)
                if self._host[slot] is None:
                    self._host[slot] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() if torch.cuda.is_available()
                                        else torch.empty(v.shape, dtype=v.dtype) for k, v in maps.items()}
                for k, v in maps.items():
                    self._host[slot][k].copy_(v, non_blocking=True)
                ev = torch.cuda.Event() if torch.cuda.is_available() else None
                if ev is not None:
                    ev.record()
                self._events[slot] = ev
                if pending is not None:  # hand out the previous view while this one renders / copies
                    yield self._finish(*pending)
                pending = (idx, slot)
            if pending is not None:
                yield self._finish(*pending)
        finally:
            m.train(was_training)

    def _finish(self, idx: int, slot: int):
        ev = self._events[slot]
        if ev is not None:
            ev.synchronize()
        return idx, self._host[slot]
