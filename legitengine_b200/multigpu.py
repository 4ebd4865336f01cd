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

    def __init__(self, width: int, height: int, rank: int, world: int, dist, stream: int = 0, root: int = 0, present: bool = True, bounds=None):
        import torch

        self.width, self.height, self.rank, self.world, self.dist, self.root, self.present = width, height, rank, world, dist, root, present
        # equal-row strips unless the caller supplies cost-aware ones (sharding.rebalance_bounds)
        self.bounds = list(bounds) if bounds is not None else sharding.strip_bounds(height, world)
        if len(self.bounds) != world or any(y0 % sharding.GRANULE for y0, _ in self.bounds):
            raise ValueError("StripRenderer: one strip per rank, starting on multiples of %d rows" % sharding.GRANULE)
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


class P2PStripRenderer(StripRenderer):
    """Strip renderer whose halo rows move through our own kernels over NVLink peer memory (CUDA IPC mappings) instead of NCCL:
    per exchange ONE copy kernel pulls every slab this rank needs straight out of the owners' images (lgcu_copy_rows), ordered
    by flag words that the owners' streams write into this GPU's memory (lgcu_signal_flags / lgcu_wait_flags). The swapchain
    strips are PUSHED into the presenting rank's image as soon as a rank finishes. torch.distributed is used once, to swap the
    IPC handles. Per frame and rank (F = frame number, bumped on the device so the whole frame replays from a CUDA graph):

        bump F | root: signal FREE=F to pushers | wait ACK >= F-1 from my pullers
        front  | signal FRONT=F to chain pullers  | wait FRONT >= F from chain sources  | pull chain halos
        chains | signal CHAINS=F to gather pullers | wait CHAINS >= F from gather sources | pull gather halos | signal ACK=F to my sources
        gather + final | pushers: wait FREE >= F, push strip into root's swapchain, signal DELIVERED=F | root: wait DELIVERED >= F

    With `direct_present` (default) there is no push: a non-root rank's swapchain target IS the root's image (peer mapping put behind the
    renderer's external swapchain proxy), so the composite kernel's BGRA8 stores travel over NVLink while it runs; the rank waits
    FREE >= F before gather + final and signals DELIVERED=F after it.
    No wait depends on a later signal of the waiting GPU, streams are in order, so the protocol cannot deadlock."""

    FRONT, CHAINS, DELIVERED, ACK, FREE = range(5)

    def __init__(self, *args, direct_present: bool = True, fused_exchange: bool = True, **kwargs):
        stream = kwargs.get("stream") or (args[5] if len(args) > 5 else None)  # StripRenderer(width, height, rank, world, dist, stream, ...)
        if not stream:
            raise ValueError("P2PStripRenderer needs an explicit CUDA stream, passed as stream=... (the same one the harness renders on)")
        super().__init__(*args, **kwargs)
        self._lgcu = abi.load_lgcu()
        self._stream = C.c_void_p(stream)
        self._direct_present = direct_present
        # fused_exchange: every exchange step (signal -> wait -> pull -> acknowledge) is ONE kernel (lgcu_exchange) instead of up to four
        # launches; needs the composite without a copy (the separate push keeps the unfused sequence)
        self._fused = fused_exchange and direct_present
        self._ready = False
        self._peer_ptrs: List[int] = []

    # -- one-time set-up ----------------------------------------------------------------------------------------------------
    def _setup(self):
        torch, dist, rank, world = self._torch, self.dist, self.rank, self.world
        # image memory exists after a first Execute of every stage
        self.renderer.render_stages(harness.STAGE_ALL, self.rows if self.rows[1] > self.rows[0] else (0, 0))
        self.renderer.sync()
        self._flags = harness.device_alloc_zeroed(4096)
        mine = {name: harness.ipc_export_image(self.renderer, name) for name in self.EXCHANGED}
        mine["__flags__"] = harness.ipc_export_ptr(self._flags)
        everyone: List[Optional[dict]] = [None] * world
        dist.all_gather_object(everyone, mine)
        base: List[Dict[str, int]] = []
        for peer in range(world):
            if peer == rank:
                base.append({name: int(self.renderer.image_desc(name).base) for name in self.EXCHANGED} | {"__flags__": self._flags})
            else:
                opened = {}
                for name, h in everyone[peer].items():
                    if name == "swapchain" and peer != self.root:
                        continue  # only the presenting rank's swapchain image is ever written by a peer
                    try:
                        opened[name] = harness.ipc_open(h)
                    except RuntimeError as e:
                        raise RuntimeError(f"rank {rank}: cannot map '{name}' of rank {peer} ({e}); mapped so far from that rank: {sorted(opened)}") from e
                self._peer_ptrs += list(opened.values())
                base.append(opened)
        desc = {name: self.renderer.image_desc(name) for name in self.EXCHANGED}

        def slab(t: sharding.Transfer):
            d = desc[t.image]
            pitch = int(d.levelPitch[t.level])
            return int(d.levelOffset[t.level]) + t.row0 * pitch, (t.row1 - t.row0) * pitch

        def copies(plan, pull: bool):
            items = []
            for t in plan:
                if (t.dst if pull else t.src) != rank:
                    continue
                off, nbytes = slab(t)
                items.append(abi.RowCopy(base[t.src][t.image] + off, base[t.dst][t.image] + off, nbytes))
            return (abi.RowCopy * max(len(items), 1))(*items), len(items)

        def flag_list(stage: int, owners, writer: int):
            """addresses of flag word (stage, writer) in the memory of each rank in `owners`"""
            ptrs = [base[o]["__flags__"] + 4 * (stage * world + writer) for o in sorted(owners)]
            return (C.c_void_p * max(len(ptrs), 1))(*ptrs), len(ptrs)

        def local_flags(stage: int, writers):
            ptrs = [self._flags + 4 * (stage * world + w) for w in sorted(writers)]
            return (C.c_void_p * max(len(ptrs), 1))(*ptrs), len(ptrs)

        pc, pg = self.plan_chains, self.plan_gather
        pp = self.plan_present if self.present else []
        self._counter = C.c_void_p(self._flags + 4 * (5 * world))
        self._pull_chains, self._pull_gather, self._push = copies(pc, True), copies(pg, True), copies(pp, False)
        if self._direct_present and self.present and rank != self.root and self.rows[1] > self.rows[0]:
            # composite without a copy: this rank's final pass writes its strip straight into the presenting GPU's swapchain image
            self.renderer.set_external_swapchain(base[self.root]["swapchain"])
            self._external_swapchain = True
            self._push = ((abi.RowCopy * 1)(), 0)
        chain_pullers, chain_sources = {t.dst for t in pc if t.src == rank}, {t.src for t in pc if t.dst == rank}
        gather_pullers, gather_sources = {t.dst for t in pg if t.src == rank}, {t.src for t in pg if t.dst == rank}
        pushers = {t.src for t in pp}
        self._sig_front, self._wait_front = flag_list(self.FRONT, chain_pullers, rank), local_flags(self.FRONT, chain_sources)
        self._sig_chains, self._wait_chains = flag_list(self.CHAINS, gather_pullers, rank), local_flags(self.CHAINS, gather_sources)
        self._sig_ack, self._wait_ack = flag_list(self.ACK, chain_sources | gather_sources, rank), local_flags(self.ACK, chain_pullers | gather_pullers)
        self._is_pusher = rank in pushers
        self._sig_free = flag_list(self.FREE, pushers, rank) if rank == self.root else ((C.c_void_p * 1)(), 0)
        self._wait_free = local_flags(self.FREE, {self.root}) if self._is_pusher else ((C.c_void_p * 1)(), 0)
        self._sig_delivered = flag_list(self.DELIVERED, {self.root}, rank) if self._is_pusher else ((C.c_void_p * 1)(), 0)
        self._wait_delivered = local_flags(self.DELIVERED, pushers) if rank == self.root else ((C.c_void_p * 1)(), 0)
        self.received_bytes = sum(c.bytes for c in self._pull_chains[0][: self._pull_chains[1]]) + sum(c.bytes for c in self._pull_gather[0][: self._pull_gather[1]])
        if self._fused:
            if max(self._pull_chains[1], self._pull_gather[1]) > 64:
                self._fused = False  # more slabs than one exchange launch takes: keep the chunked copies
            else:
                done = self._flags + 4 * (5 * world + 1)  # a zero-initialised word after the frame counter

                def step(sig_before=None, wait=None, lag=0, copies=None, sig_after=None, bump=False):
                    none_f, none_c = ((C.c_void_p * 1)(), 0), ((abi.RowCopy * 1)(), 0)
                    sb, w, cp, sa = sig_before or none_f, wait or none_f, copies or none_c, sig_after or none_f
                    return abi.ExchangeDesc(sb[0], sb[1], w[0], w[1], lag, cp[0], cp[1], sa[0], sa[1], self._counter, C.c_void_p(done), 1 if bump else 0), (sb, w, cp, sa)

                def merged(a, b):
                    ptrs = list(a[0][: a[1]]) + list(b[0][: b[1]])
                    return (C.c_void_p * max(len(ptrs), 1))(*ptrs), len(ptrs)

                free_wait = self._wait_free if self._is_pusher else ((C.c_void_p * 1)(), 0)
                self._steps = {
                    "open": step(sig_before=self._sig_free, wait=self._wait_ack, lag=1, bump=True),
                    "front": step(sig_before=self._sig_front, wait=self._wait_front, copies=self._pull_chains),
                    # FREE is waited for here, together with the CHAINS flags: it has to precede gather + final, which writes the strip into
                    # the presenting GPU's image
                    "chains": step(sig_before=self._sig_chains, wait=merged(self._wait_chains, free_wait), copies=self._pull_gather, sig_after=self._sig_ack),
                    "close": step(sig_before=self._sig_delivered, wait=self._wait_delivered),
                }
        torch.cuda.synchronize()
        dist.barrier()
        self._ready = True

    # -- per-frame ------------------------------------------------------------------------------------------------------------
    def _signal(self, lst):
        if lst[1]:
            abi.check(self._lgcu.lgcu_signal_flags(lst[0], lst[1], self._counter, self._stream), "lgcu_signal_flags")

    def _wait(self, lst, lag=0):
        if lst[1]:
            abi.check(self._lgcu.lgcu_wait_flags(lst[0], lst[1], self._counter, lag, self._stream), "lgcu_wait_flags")

    def _copy(self, lst):
        if lst[1]:
            abi.check(self._lgcu.lgcu_copy_rows(lst[0], lst[1], self._stream), "lgcu_copy_rows")

    STAGE_MARKS = ("start", "front", "chain_halo", "chains", "gather_halo", "gather_final", "present")

    def render(self, gi_flags: int = abi.GI_DEFAULT, marks=None) -> None:
        """One frame. `marks` (optional): a list that receives one torch.cuda.Event per STAGE_MARKS entry, recorded on the current
        stream after that stage was enqueued (un-captured frames only) — the per-rank, per-stage profile of bench.py --shard strips."""
        if not self._ready:
            self._setup()
        r, rows = self.renderer, self.rows
        have = rows[1] > rows[0]

        def mark():
            if marks is not None:
                ev = self._torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)

        if self._fused:
            return self._render_fused(gi_flags, mark)
        abi.check(self._lgcu.lgcu_frame_counter_bump(self._counter, self._stream), "lgcu_frame_counter_bump")
        self._signal(self._sig_free)
        self._wait(self._wait_ack, lag=1)
        mark()
        if have:
            r.render_stages(harness.STAGE_FRONT, rows, gi_flags=gi_flags)
        mark()
        self._signal(self._sig_front)
        self._wait(self._wait_front)
        self._copy(self._pull_chains)
        mark()
        if have:
            r.render_stages(harness.STAGE_CHAINS, rows, gi_flags=gi_flags)
        mark()
        self._signal(self._sig_chains)
        self._wait(self._wait_chains)
        self._copy(self._pull_gather)
        self._signal(self._sig_ack)
        mark()
        direct = self._direct_present and self._is_pusher
        if direct:
            self._wait(self._wait_free)  # the presenting GPU is done with the previous frame's image
        if have:
            r.render_stages(harness.STAGE_GATHER | harness.STAGE_FINAL, rows, gi_flags=gi_flags)
        mark()
        if self._is_pusher:
            if not direct:
                self._wait(self._wait_free)
                self._copy(self._push)
            self._signal(self._sig_delivered)
        self._wait(self._wait_delivered)
        mark()

    def _exchange(self, name: str) -> None:
        desc, _keepalive = self._steps[name]
        abi.check(self._lgcu.lgcu_exchange(C.byref(desc), self._stream), "lgcu_exchange")

    def _render_fused(self, gi_flags: int, mark) -> None:
        """The frame with every exchange step as one launch: open | front | exchange | chains | exchange | gather + final | close."""
        r, rows = self.renderer, self.rows
        have = rows[1] > rows[0]
        self._exchange("open")    # bump F, root: signal FREE, wait ACK >= F-1
        mark()
        if have:
            r.render_stages(harness.STAGE_FRONT, rows, gi_flags=gi_flags)
        mark()
        self._exchange("front")   # signal FRONT, wait FRONT, pull chain halos
        mark()
        if have:
            r.render_stages(harness.STAGE_CHAINS, rows, gi_flags=gi_flags)
        mark()
        self._exchange("chains")  # signal CHAINS, wait CHAINS (+ FREE), pull gather halos, last CTA: signal ACK
        mark()
        if have:
            r.render_stages(harness.STAGE_GATHER | harness.STAGE_FINAL, rows, gi_flags=gi_flags)
        mark()
        self._exchange("close")   # pushers: signal DELIVERED; root: wait DELIVERED
        mark()

    def capture(self, gi_flags: int = abi.GI_DEFAULT) -> None:
        torch = self._torch
        self.render(gi_flags)
        self.render(gi_flags)
        torch.cuda.synchronize()
        self.dist.barrier()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=torch.cuda.current_stream()):
            self.render(gi_flags)

    def close(self):
        self._torch.cuda.synchronize()
        self.dist.barrier()  # nobody may still be reading this rank's memory
        if getattr(self, "_external_swapchain", False):
            self.renderer.set_external_swapchain(0)  # drop the view of the presenting GPU's image before its mapping goes
        for p in self._peer_ptrs:
            harness.ipc_close(p)
        self._peer_ptrs = []
        # every mapping of this rank's memory is closed before the memory is freed: a peer that still held the old mapping while this
        # rank re-allocated and re-exported the same block could not open the new handle (cudaIpcOpenMemHandle: invalid argument)
        self.dist.barrier()
        if self._ready:
            harness.device_free(self._flags)
        super().close()
