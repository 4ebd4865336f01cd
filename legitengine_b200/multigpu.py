"""Strip-sharded rendering of ONE frame on several GPUs of a box (BASELINE configs[3]).

One process per GPU (torchrun); every rank owns a `harness.Renderer` with full-size images and renders its row strip in
stages, exchanging halo rows with the ranks that own them between the stages (plans: legitengine_b200/sharding.py). Rows of an
image level are contiguous in the linear layout, so every transfer is one contiguous send/recv on a byte view of the image
memory; `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) moves them. torch computes nothing here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

from . import abi, harness, sharding


class _DevicePtr:
    """Adapter giving a raw device allocation the __cuda_array_interface__, so torch can view it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def image_bytes(desc: abi.LgcuImage) -> int:
    last = desc.imageMipCount - 1
    h = max(desc.height >> last, 1)
    return int(desc.levelOffset[last]) + int(desc.levelPitch[last]) * h


def run_transfers(plan: Sequence[sharding.Transfer], views: Dict[str, Tuple[object, abi.LgcuImage]], rank: int, dist, group=None) -> int:
    """Executes this rank's part of `plan` (sends of rows it owns, receives of rows it needs). `views[name]` is (flat uint8
    tensor over the image allocation, descriptor). Returns the bytes this rank received. Collective: every rank calls it with
    the same plan."""
    ops, received = [], 0
    for t in plan:
        if t.src != rank and t.dst != rank:
            continue
        tensor, desc = views[t.image]
        pitch = int(desc.levelPitch[t.level])
        begin = int(desc.levelOffset[t.level]) + t.row0 * pitch
        slab = tensor[begin: begin + (t.row1 - t.row0) * pitch]
        if t.src == rank:
            ops.append(dist.P2POp(dist.isend, slab, t.dst, group=group))
        else:
            ops.append(dist.P2POp(dist.irecv, slab, t.src, group=group))
            received += slab.numel()
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return received


class StripRenderer:
    """Renders rows `bounds[rank]` of every frame; after `render()` rank `root` holds the complete swapchain image."""

    EXCHANGED = sharding.CHAINS + sharding.BLURRED + ("swapchain",)

    def __init__(self, width: int, height: int, rank: int, world: int, dist, stream: int = 0, root: int = 0, present: bool = True):
        import torch

        self.width, self.height, self.rank, self.world, self.dist, self.root, self.present = width, height, rank, world, dist, root, present
        self.bounds = sharding.strip_bounds(height, world)
        self.rows = self.bounds[rank]
        self.renderer = harness.Renderer(width, height, stream=stream)
        self.plan_chains = sharding.plan_chains(self.bounds, width, height)
        self.plan_gather = sharding.plan_gather(self.bounds, width, height)
        self.plan_present = sharding.plan_present(self.bounds, height, root)
        self._views: Optional[Dict[str, Tuple[object, abi.LgcuImage]]] = None
        self._torch = torch
        self.received_bytes = 0
        self._graph = None

    def _ensure_views(self):
        if self._views is None:  # image memory exists after the first Execute
            self._views = {}
            for name in self.EXCHANGED:
                d = self.renderer.image_desc(name)
                self._views[name] = (self._torch.as_tensor(_DevicePtr(int(d.base), image_bytes(d)), device="cuda"), d)
        return self._views

    def upload_strip(self, fragments_host_ptr: int, pitch: int) -> None:
        if self.rows[1] > self.rows[0]:
            self.renderer.upload_fragments(fragments_host_ptr, pitch, rows=self.rows)

    def render(self, gi_flags: int = abi.GI_DEFAULT) -> None:
        """One frame. The caller's current torch stream must be the renderer's stream (kernels and NCCL then order by stream)."""
        r, rows = self.renderer, self.rows
        have = rows[1] > rows[0]
        if have:
            r.render_stages(harness.STAGE_FRONT, rows, gi_flags=gi_flags)
        views = self._ensure_views() if have or self._views is not None else self._allocate_idle()
        got = run_transfers(self.plan_chains, views, self.rank, self.dist)
        if have:
            r.render_stages(harness.STAGE_CHAINS, rows, gi_flags=gi_flags)
        got += run_transfers(self.plan_gather, views, self.rank, self.dist)
        if have:
            r.render_stages(harness.STAGE_GATHER | harness.STAGE_FINAL, rows, gi_flags=gi_flags)
        if self.present:
            got += run_transfers(self.plan_present, views, self.rank, self.dist)
        self.received_bytes = got

    def capture(self, gi_flags: int = abi.GI_DEFAULT) -> None:
        """Captures one whole sharded frame — the CUDA kernels of every stage AND the NCCL halo transfers between them — into a
        CUDA graph on the renderer's stream (which must be the current torch stream). At 8 GPUs a strip is well under a
        millisecond of GPU work, so per-frame launch + Python cost has to disappear for the strips to scale."""
        torch = self._torch
        self.render(gi_flags)  # allocations, NCCL connection set-up and the cached views must exist before capture
        self.render(gi_flags)
        torch.cuda.synchronize()
        self.dist.barrier()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=torch.cuda.current_stream()):
            self.render(gi_flags)

    def replay(self) -> None:
        self._graph.replay()

    def release_graph(self) -> None:
        self._graph = None

    def _allocate_idle(self):
        # a rank with an empty strip still has to allocate its images once to take part in the collectives' bookkeeping
        self.renderer.render_stages(harness.STAGE_FRONT, (0, 0))
        return self._ensure_views()

    def close(self):
        self._views = None
        self.renderer.close()
