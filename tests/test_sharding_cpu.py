"""CPU tests of the multi-GPU strip sharding (no GPU): the partition / halo plans (legitengine_b200/sharding.py) and the
transfer executor (legitengine_b200/multigpu.run_transfers) under world_size 2 and 3 with the gloo backend.

The distributed test emulates the strip pipeline with the CPU ORACLE standing in for the CUDA stages (front -> exchange ->
chains -> exchange -> gather/final -> present) and checks that every rank's strip, and the composited swapchain on rank 0,
equal the whole-frame oracle bit for bit — i.e. that the halo plans carry every row the stages read."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from legitengine_b200 import abi, images, passes, scene, sharding


def test_strip_bounds_cover_and_align():
    for H in (4320, 2160, 1080, 141, 16, 5):
        for world in (1, 2, 3, 4, 8):
            b = sharding.strip_bounds(H, world)
            assert b[0][0] == 0 and b[-1][1] == H and len(b) == world
            for (a0, a1), (b0, b1) in zip(b, b[1:]):
                assert a1 == b0 and a0 <= a1
            for y0, y1 in b[:-1]:
                assert y0 % sharding.GRANULE == 0 and (y1 % sharding.GRANULE == 0 or y1 == H)
            sizes = [y1 - y0 for y0, y1 in b]
            assert max(sizes) - min(sizes) <= sharding.GRANULE
    assert sharding.strip_bounds(4320, 8)[1] == (528, 1072)


def test_level_rows_partition_every_level():
    H, world = 4320, 8
    b = sharding.strip_bounds(H, world)
    for l in range(0, sharding.FRONT_LEVELS + 1):
        rows = [sharding.level_rows(s, l, H) for s in b]
        assert rows[0][0] == 0 and rows[-1][1] == H >> l
        for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
            assert a1 == b0  # strip-local levels tile the level without overlap


def test_plans_only_move_rows_the_source_owns_and_sizes_are_small():
    W, H, world = 7680, 4320, 8
    b = sharding.strip_bounds(H, world)
    for plan in (sharding.plan_chains(b, W, H), sharding.plan_gather(b, W, H), sharding.plan_present(b, H)):
        assert plan
        for t in plan:
            o0, o1 = sharding.level_rows(b[t.src], t.level, H)
            assert o0 <= t.row0 < t.row1 <= o1 and t.src != t.dst
    pitch = {(n, l): images.make_layout(abi.FORMAT_R16G16B16A16_SFLOAT, W, H, 10)[0].levelPitch[l] for n in sharding.CHAINS + sharding.BLURRED for l in range(10)}
    per_rank = sharding.transfer_bytes([t for t in sharding.plan_gather(b, W, H) if t.dst == 3], pitch)
    assert per_rank < 8e6  # ~7.6 MB of halo per interior GPU at 8K (SURVEY.md §8e estimated ~6.4 MB)
    chains3 = sharding.transfer_bytes([t for t in sharding.plan_chains(b, W, H) if t.dst == 3], pitch)
    assert chains3 < 4e6   # blur halos + the whole of level 4 (480 x 270 texels x 2 chains)
    # a single rank needs nothing
    assert sharding.plan_chains(sharding.strip_bounds(H, 1), W, H) == [] and sharding.plan_gather(sharding.strip_bounds(H, 1), W, H) == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _v(img, base=0, n=None):
    return C.byref(img.view(base, n))


def _oracle_stage(be, fi, p, inp, stage, rows, W, Hh):
    """The CUDA stages of multigpu.StripRenderer, restated with oracle passes (same row semantics)."""
    r = C.byref(abi.LgcuRows(*rows))
    levels = passes.mip_levels_built(W, Hh)
    grid = min(sharding.FRONT_LEVELS, levels - 1)
    chains = ((fi.directLight, fi.blurredDirectLight), (fi.depthMoments, fi.blurredDepthMoments))

    def blur(src, dst, l, rr):
        w, h = images.mip_size(W, Hh, l)
        bp = abi.BlurLayerBuilderData((C.c_int32 * 4)(w, h, 0, 0), 0 if l == 0 else 2)
        be.blur_level(C.byref(bp), _v(src, l, 1), _v(dst, l, 1), rr)

    if stage == "front":
        passes.run_pass_list(be, fi, p, inp, rows=rows, stop_after="light")
        for src, dst in chains:
            blur(src, dst, 0, r)
            for l in range(1, grid + 1):
                be.mip_level(C.byref(p.mip), _v(src, l - 1, 1), _v(src, l, 1), r)
    elif stage == "chains":
        for src, dst in chains:
            for l in range(1, grid + 1):
                blur(src, dst, l, r)
            for l in range(grid + 1, levels):  # the tail: whole levels on every rank
                be.mip_level(C.byref(p.mip), _v(src, l - 1, 1), _v(src, l, 1), None)
            for l in range(grid + 1, levels):
                blur(src, dst, l, None)
    elif stage == "gather":
        be.gi_gather(C.byref(p.indirect), _v(fi.blurredDirectLight), _v(fi.blurredDepthMoments), _v(fi.normal), _v(fi.depthStencil), _v(fi.indirectLight), 0, r)
        be.denoise(C.byref(p.denoiser), _v(fi.indirectLight), _v(fi.normal), _v(fi.depthMoments), _v(fi.denoisedIndirectLight), r)
        be.final_gather(C.byref(p.final), _v(fi.directLight), _v(fi.blurredDirectLight), _v(fi.albedo), _v(fi.denoisedIndirectLight), _v(fi.swapchain), r)


def _worker(rank, world, port, W, Hh, result_dir, uneven=False):
    import torch
    import torch.distributed as dist

    from legitengine_b200 import multigpu
    from oracle import loader

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        be = loader.port()
        be.set_num_threads(2)
        sc = scene.make_scene(31, W, Hh, n_boxes=24, shadow_size=128)
        p = passes.make_params(W, Hh, sc.matrices, 0)
        whole = passes.FrameImages(W, Hh, images.HostImage, shadow_size=128)
        passes.run_pass_list(be, whole, p, passes.upload_inputs(whole, sc))

        fi = passes.FrameImages(W, Hh, images.HostImage, shadow_size=128)  # poison-filled: rows that never arrive stay 0xCD
        inp = passes.upload_inputs(fi, sc)
        bounds = sharding.strip_bounds(Hh, world)
        if uneven:  # cost-aware boundaries: the plans must still deliver every halo row
            bounds = sharding.rebalance_bounds(bounds, [1.0 + 1.5 * r for r in range(world)], Hh)
            assert bounds != sharding.strip_bounds(Hh, world)
        rows = bounds[rank]
        views = {n: (torch.from_numpy(getattr(fi, n).buf), getattr(fi, n).desc) for n in multigpu.StripRenderer.EXCHANGED}
        _oracle_stage(be, fi, p, inp, "front", rows, W, Hh)
        multigpu.run_transfers(sharding.plan_chains(bounds, W, Hh), views, rank, dist)
        _oracle_stage(be, fi, p, inp, "chains", rows, W, Hh)
        multigpu.run_transfers(sharding.plan_gather(bounds, W, Hh), views, rank, dist)
        _oracle_stage(be, fi, p, inp, "gather", rows, W, Hh)
        multigpu.run_transfers(sharding.plan_present(bounds, Hh), views, rank, dist)

        errors = []
        for name in ("directLight", "blurredDirectLight", "depthMoments", "blurredDepthMoments", "indirectLight", "swapchain"):
            a, b = getattr(fi, name), getattr(whole, name)
            y0, y1 = rows
            if not np.array_equal(a.level_bytes(0)[y0:y1], b.level_bytes(0)[y0:y1]):
                errors.append(f"rank {rank}: {name} strip differs from the whole frame")
        for l in range(sharding.FRONT_LEVELS + 1, passes.mip_levels_built(W, Hh)):
            if not fi.blurredDirectLight.levels_equal(whole.blurredDirectLight, l):
                errors.append(f"rank {rank}: coarse level {l} differs")
        if rank == 0 and not fi.swapchain.levels_equal(whole.swapchain, 0):
            errors.append("rank 0: composited swapchain differs from the whole frame")
        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as f:
            f.write("\n".join(errors) if errors else "ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,size,uneven", [(2, (96, 80), False), (3, (160, 112), False), (3, (160, 112), True)])
def test_strip_pipeline_gloo(world, size, uneven, tmp_path):
    import torch.multiprocessing as mp

    W, Hh = size
    mp.spawn(_worker, args=(world, _free_port(), W, Hh, str(tmp_path), uneven), nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / f"rank{rank}.txt").read_text() == "ok"


def test_rebalance_bounds_equalises_measured_cost():
    """Cost-aware strips: contiguous cover, granule-aligned, at least one granule each, and the predicted max cost drops to
    within one granule's worth of the mean."""
    H, world = 4320, 8
    b = sharding.strip_bounds(H, world)
    costs = [0.885, 1.003, 1.061, 1.098, 1.103, 0.903, 0.810, 0.713]  # profiles/r01f: gather+final ms per rank, 8K on 8 GPUs
    nb = sharding.rebalance_bounds(b, costs, H)
    assert nb[0][0] == 0 and nb[-1][1] == H and all(nb[i][1] == nb[i + 1][0] for i in range(world - 1))
    assert all(y0 % sharding.GRANULE == 0 and y1 > y0 for y0, y1 in nb)

    def predicted(bounds):
        dens = np.zeros(H)
        for (y0, y1), c in zip(b, costs):
            dens[y0:y1] = c / (y1 - y0)
        return [dens[y0:y1].sum() for y0, y1 in bounds]

    before, after = predicted(b), predicted(nb)
    mean = sum(costs) / world
    assert abs(sum(after) - sum(costs)) < 1e-9
    assert max(after) < max(before) and max(after) <= mean + 1.5 * sharding.GRANULE * max(costs) / 528
    # fixed points and degenerate inputs
    assert sharding.rebalance_bounds(b, [1.0] * world, H) == b
    assert sharding.rebalance_bounds([(0, H)], [3.0], H) == [(0, H)]
    tiny = sharding.strip_bounds(48, 8)  # 3 granules for 8 ranks: left alone
    assert sharding.rebalance_bounds(tiny, [1.0] * 8, 48) == tiny
    skew = sharding.rebalance_bounds(sharding.strip_bounds(256, 4), [100.0, 1.0, 1.0, 1.0], 256)
    assert all(y1 - y0 >= sharding.GRANULE for y0, y1 in skew) and skew[0][1] < 64


def test_cost_profile_is_learnt_across_rebalancing_rounds():
    """sharding.refine_cost_density + bounds_from_density (bench.py's re-balancing): per-strip measurements of DIFFERENT partitions refine
    one per-block cost profile; the partitions converge on equal cost, stay on the 16-row grid, cover the frame, keep every rank non-empty,
    and a profile with a sharp expensive band (the horizon rows of the synthetic frame) ends within a few per cent of balance."""
    import math

    H, world = 4320, 8
    blocks = H // sharding.GRANULE
    true = [1.0 + 0.4 * math.sin(b / 25.0) + (1.5 if 128 <= b < 140 else 0.0) for b in range(blocks)]

    def measure(bounds):
        return [sum(true[y0 // 16:(y1 + 15) // 16]) for y0, y1 in bounds]

    bounds, density, spreads = sharding.strip_bounds(H, world), None, []
    for _ in range(6):
        costs = measure(bounds)
        spreads.append(max(costs) / (sum(costs) / world))
        density = sharding.refine_cost_density(density, bounds, costs, H)
        assert len(density) == blocks and all(d > 0 for d in density)
        bounds = sharding.bounds_from_density(density, world, H)
        assert bounds[0][0] == 0 and bounds[-1][1] == H
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:])) and all(y1 > y0 and y0 % 16 == 0 for y0, y1 in bounds)
    # (with a 16-row granule and a sharp band the last rounds hop between two neighbouring partitions: bench.py keeps the best one measured)
    assert spreads[0] > 1.15 and min(spreads) < 1.04 and max(spreads[1:]) < 1.06, spreads
    # the profile keeps what an earlier partition taught: after the rounds its shape correlates with the true one inside a strip too
    y0, y1 = bounds[2]
    seg = slice(y0 // 16, y1 // 16)
    est, tru = density[seg], true[seg]
    me, mt = sum(est) / len(est), sum(tru) / len(tru)
    cov = sum((a - me) * (b - mt) for a, b in zip(est, tru))
    assert cov >= 0.0
    # degenerate inputs: one rank, more ranks than blocks, a rank with an empty strip and no cost
    assert sharding.bounds_from_density([1.0] * 4, 1, 64) == [(0, 64)]
    assert sharding.bounds_from_density([1.0] * 2, 4, 32) == sharding.strip_bounds(32, 4)
    d = sharding.refine_cost_density(None, [(0, 32), (32, 32)], [2.0, 0.0], 32)
    assert d == [1.0, 1.0]
