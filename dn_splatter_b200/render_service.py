"""Render-all-views forward service: what `gs-mesh`, `ns-eval` and `render_model.py` do with the model
— `for camera in views: model.get_outputs_for_camera(camera)` (reference export_mesh.py:360-367, 863-905, 965-1017;
dn_pipeline.py:194-214; scripts/render_model.py:47-69; utils/utils.py:331-443) — as a service that keeps the device busy.

CUDA path (`graph=True`, the default on a GPU): the forward of one view (project -> bin/sort -> composite -> depth fill +
surface normal, ~20 launches and ~1 ms of Python) is captured ONCE per resolution as a CUDA graph in `n_slots` copies
that write into their own static output maps; a view is then one 148-byte camera upload + one graph launch.  With
`to_host=True` the selected maps of slot s are copied to pinned host buffers on a side stream while slot s+1 renders, and
a view is handed out when its copy has landed.  The intersection buffers have a fixed capacity inside a graph: every
replay's count is read back with the maps, and a view that needed more is re-rendered after re-capturing with a larger
capacity — the consumer never sees a truncated render.

`graph=False` is the plain loop (also what non-CUDA models, i.e. the CPU-proxy tests, use).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

DEFAULT_KEYS = ("rgb", "depth", "normal", "surface_normal", "accumulation")


class _ForwardGraphs:
    """`n_slots` captured copies of model.get_outputs for one resolution."""

    def __init__(self, model, camera, keys: Sequence[str], n_slots: int, capacity: int):
        from .graph_step import GraphedTrainStep  # camera block upload is shared

        self.model, self.keys, self.n_slots = model, tuple(keys), n_slots
        self.device = model.device
        self.size = (int(camera.width.flatten()[0]), int(camera.height.flatten()[0]))
        self.capacity = int(capacity)
        self._cam_dev = torch.zeros(37, device=self.device)
        self.cam = {"viewmat": self._cam_dev[:16].view(4, 4), "K": self._cam_dev[16:25].view(3, 3),
                    "c2w": self._cam_dev[25:37].view(3, 4), "capacity": self.capacity}
        self._camera = camera
        self._load_camera = GraphedTrainStep.load_camera.__get__(self)  # same pinned 148-byte upload
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.maps: List[Dict[str, Tensor]] = []
        self.counts: List[Tensor] = []
        self._capture()

    def _eager(self) -> Tuple[Dict[str, Tensor], Tensor]:
        m = self.model
        m.__dict__["_graph_cam"] = self.cam
        try:
            out = m.get_outputs(self._camera)
            return {k: out[k] for k in self.keys if k in out}, m.raster_out.info["n_isects_dev"]
        finally:
            m.__dict__["_graph_cam"] = None

    @torch.no_grad()
    def _capture(self) -> None:
        self.cam["capacity"] = self.capacity
        self._load_camera(self._camera)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graphs, self.maps, self.counts, pool = [], [], [], None
        for _ in range(self.n_slots):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                maps, count = self._eager()
            pool = g.pool()
            self.graphs.append(g)
            self.maps.append(maps)
            self.counts.append(count)

    def replay(self, camera, slot: int) -> None:
        self._load_camera(camera)
        self.graphs[slot].replay()


class ViewRenderer:
    def __init__(self, model, keys: Sequence[str] = DEFAULT_KEYS, to_host: bool = False, n_host_buffers: int = 2,
                 graph: Optional[bool] = None, n_slots: int = 2):
        self.model, self.keys, self.to_host = model, tuple(keys), to_host
        self._host: list = [None] * n_host_buffers
        self._events: list = [None] * n_host_buffers
        self.graph = (model.device.type == "cuda") if graph is None else bool(graph)
        self.n_slots = max(2, n_slots) if to_host else max(1, n_slots)
        self._graphs: Dict[Tuple[int, int], _ForwardGraphs] = {}
        self.recaptures = 0

    # ------------------------------------------------------------------ plain loop
    @torch.no_grad()
    def _render_eager(self, cameras: Iterable) -> Iterator[Tuple[int, Dict[str, Tensor]]]:
        m = self.model
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

    def _finish(self, idx: int, slot: int):
        ev = self._events[slot]
        if ev is not None:
            ev.synchronize()
        return idx, self._host[slot]

    # ------------------------------------------------------------------ captured forward
    def _graphs_for(self, cam) -> _ForwardGraphs:
        from .rasterize import suggested_capacity

        m = self.model
        key = (int(cam.width.flatten()[0]), int(cam.height.flatten()[0]))
        fg = self._graphs.get(key)
        if fg is None or fg.model.num_points != m.num_points:
            cfg = m.config
            with torch.no_grad():  # two synchronous views seed the capacity statistics for this size if there are none
                cap = suggested_capacity(m.num_points, key[0], key[1], cfg.predict_normals, cfg.exact_isect_lists, m.device.index,
                                         0 if cfg.exact_isect_lists else cfg.list_shift)
                if cap <= 0:
                    from .rasterize import round_capacity

                    m.get_outputs(cam)
                    cap = round_capacity(int(int(m.raster_out.info["n_isects_dev"]) * 1.3) + 4096)
            fg = self._graphs[key] = _ForwardGraphs(m, cam, self.keys, self.n_slots, cap)
        return fg

    @torch.no_grad()
    def _render_graphed(self, cameras: Iterable) -> Iterator[Tuple[int, Dict[str, Tensor]]]:
        cams = list(cameras)
        if not cams:
            return
        dev = self.model.device
        copy_stream = torch.cuda.Stream(device=dev)
        compute = torch.cuda.current_stream()
        n_slots = self.n_slots
        host = [None] * n_slots            # pinned {key: tensor, "_count": int64[1]} per slot
        rendered = [torch.cuda.Event() for _ in range(n_slots)]
        copied = [torch.cuda.Event() for _ in range(n_slots)]
        busy = [False] * n_slots
        inflight: List[Tuple[int, int, _ForwardGraphs]] = []  # (view index, slot, graphs) in submission order

        def submit(idx):
            cam = cams[idx]
            fg = self._graphs_for(cam)
            slot = idx % n_slots
            if busy[slot]:
                compute.wait_event(copied[slot])  # the copy that still reads this slot's static maps
            fg.replay(cam, slot)
            rendered[slot].record(compute)
            if host[slot] is None or any(host[slot][k].shape != v.shape for k, v in fg.maps[slot].items()):
                host[slot] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in fg.maps[slot].items()}
                host[slot]["_count"] = torch.zeros(1, dtype=torch.int64).pin_memory()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(rendered[slot])
                host[slot]["_count"].copy_(fg.counts[slot], non_blocking=True)
                if self.to_host:
                    for k, v in fg.maps[slot].items():
                        host[slot][k].copy_(v, non_blocking=True)
                copied[slot].record(copy_stream)
            busy[slot] = True
            inflight.append((idx, slot, fg))

        def collect():
            idx, slot, fg = inflight.pop(0)
            copied[slot].synchronize()
            need = int(host[slot]["_count"])
            if need > fg.capacity:  # truncated: grow, re-capture, render this view again (synchronously)
                from .rasterize import round_capacity

                for _, s2, _ in inflight:  # drain what is in flight on the old graphs first
                    copied[s2].synchronize()
                fg.capacity = round_capacity(int(need * 1.3) + 4096)
                fg._capture()
                self.recaptures += 1
                redo = [idx] + [i for i, _, _ in inflight]
                inflight.clear()
                for s2 in range(n_slots):
                    busy[s2] = False
                out = None
                for j in redo:  # serial re-render of the affected window keeps the order
                    submit(j)
                    i2, s2, fg2 = inflight.pop(0)
                    copied[s2].synchronize()
                    maps = ({k: v.clone() for k, v in host[s2].items() if k != "_count"} if self.to_host
                            else {k: v.clone() for k, v in fg2.maps[s2].items()})
                    if out is None:
                        out = [(i2, maps)]
                    else:
                        out.append((i2, maps))
                return out
            if self.to_host:
                return [(idx, {k: v for k, v in host[slot].items() if k != "_count"})]
            return [(idx, fg.maps[slot])]

        depth = n_slots - 1 if self.to_host else 0  # views in flight while the caller consumes one
        nxt = 0
        while nxt < len(cams) or inflight:
            while nxt < len(cams) and len(inflight) <= depth:
                submit(nxt)
                nxt += 1
            for item in collect():
                yield item

    @torch.no_grad()
    def render(self, cameras: Iterable) -> Iterator[Tuple[int, Dict[str, Tensor]]]:
        """Yields (view index, {key: map}) in order.  With to_host=True the maps are pinned host tensors whose copy has
        completed when they are yielded; the next view is already rendering while the caller consumes them.  The buffers
        (host buffers, or the graph's static device maps with to_host=False) are reused round-robin: consume (or copy) a
        view's maps before asking for the next one."""
        m = self.model
        was_training = m.training
        m.eval()
        try:
            it = self._render_graphed(cameras) if (self.graph and m.device.type == "cuda") else self._render_eager(cameras)
            for item in it:
                yield item
        finally:
            m.train(was_training)
