"""Row-strip sharding of one SSVGI frame across the GPUs of a box: who owns which rows of which image level, and which
rows have to travel between ranks before each stage. Pure host logic, no device code, so it
is covered by the CPU tests (world_size-2 gloo) as well as by the 2-GPU test.

The frame is cut into horizontal strips on multiples of `GRANULE` = 16 base rows (the tile height of the frame-front kernel,
so mip levels 1..4 of a strip are built from the strip alone). Full-frame coordinates are kept everywhere: pattern index,
uv, clamp-to-edge and the blur window all use the full image size, every kernel just takes its row range (`lgcu_rows`).

Stages and what each needs from other ranks (derived from the shaders' dependency radii):

  front   K1, K2, level-0 blur copy, mips 1..4           per-pixel / tile-local: nothing
  chains  blur(-2..+1) of levels 1..4 on own rows        2 rows above, 1 below of chain levels 1..4  (blurLayerBuilder.frag:22-24)
          mips 5..9 + their blur, built whole            every row of chain level 4                   (= all-gather of level 4)
  gather  march samples at pixel distance r use LOD      blurred levels 0..4: REACH rows above and below the strip, in
          log2(0.785 (r-1)) - 2  => level l is read      level-l rows (indirectLighting.frag:217, 231-240); levels >= 5 are
          within ~10.2 * 2^l base px, +1 for bilinear     already whole on every rank
  final   K6 (r=0) + K7                                  per-pixel: nothing
  present swapchain strips -> rank 0                     every swapchain row on rank 0

A plan is a list of `Transfer(image, level, row0, row1, src, dst)`: rows [row0, row1) of `level` of `image`, owned by rank
`src`, needed by rank `dst`. Rows of a level are contiguous in the linear image layout, so a transfer is one contiguous copy.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

GRANULE = 16        # strip boundaries are multiples of this many base rows (frame-front tile height, lgcu.h LGCU_FRONT_MIP_LEVELS = 4)
FRONT_LEVELS = 4    # mip levels built strip-locally by the frame-front kernel
BLUR_ABOVE, BLUR_BELOW = 2, 1  # blur window -2..+1 (blurLayerBuilder.frag:22-24 with radius 2)
GATHER_REACH = 16   # level-l rows beyond the strip the march can touch (10.2 + 1 bilinear + 1 trilinear partner, rounded up)
MIPS = 10           # MippedProxy level count (src/Render/Common/MipBuilder.h:21)

CHAINS = ("directLight", "depthMoments")
BLURRED = ("blurredDirectLight", "blurredDepthMoments")


@dataclass(frozen=True)
class Transfer:
    image: str
    level: int
    row0: int
    row1: int
    src: int
    dst: int


def strip_bounds(height: int, world: int, granule: int = GRANULE) -> List[Tuple[int, int]]:
    """Base-row range [y0, y1) of every rank: contiguous, covering [0, height), boundaries on multiples of `granule`,
    sizes differing by at most one granule. Ranks beyond the number of granules get empty strips."""
    if world < 1 or height < 1:
        raise ValueError("strip_bounds: world and height must be positive")
    blocks = (height + granule - 1) // granule
    bounds = []
    for r in range(world):
        b0, b1 = blocks * r // world, blocks * (r + 1) // world
        bounds.append((min(b0 * granule, height), min(b1 * granule, height)))
    return bounds


def rebalance_bounds(bounds: Sequence[Tuple[int, int]], costs: Sequence[float], height: int, granule: int = GRANULE) -> List[Tuple[int, int]]:
    """Cost-aware strip boundaries. `costs[r]` is the time rank r spent on its own strip `bounds[r]` (its kernels only, not the
    time it waited for halos). The GI gather's cost per row depends on what the rows show (≈ ±25 % between strips of the synthetic
    8K scene), so equal row counts leave the slowest strip on the critical path of every frame. The measured cost is spread
    evenly over the rows of each strip (piecewise-constant cost density), and the new boundaries cut the cumulative cost into
    `world` equal parts, snapped to multiples of `granule` rows; every rank keeps at least one granule. Deterministic: all ranks
    compute the same bounds from the same all-gathered costs. Iterate (measure -> rebalance) two or three times to converge."""
    world = len(bounds)
    if world != len(costs) or world < 1:
        raise ValueError("rebalance_bounds: one cost per strip")
    blocks = (height + granule - 1) // granule
    if world == 1 or blocks <= world or not all(c > 0.0 for (y0, y1), c in zip(bounds, costs) if y1 > y0):
        return list(bounds)
    # cost of every granule block under the piecewise-constant density
    block_cost = [0.0] * blocks
    for (y0, y1), c in zip(bounds, costs):
        if y1 <= y0:
            continue
        per_row = c / (y1 - y0)
        for b in range(y0 // granule, (y1 + granule - 1) // granule):
            lo, hi = max(b * granule, y0), min((b + 1) * granule, y1, height)
            if hi > lo:
                block_cost[b] += per_row * (hi - lo)
    total = sum(block_cost)
    cuts, acc, b = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while b < blocks and acc + block_cost[b] * 0.5 < target:  # nearest block boundary to the target
            acc += block_cost[b]
            b += 1
        cut = min(max(b, cuts[-1] + 1), blocks - (world - r))  # at least one block per rank, and room for the ranks after it
        cuts.append(cut)
        if cut != b:  # re-sync the accumulator with the boundary actually chosen
            acc, b = sum(block_cost[:cut]), cut
    cuts.append(blocks)
    return [(min(cuts[r] * granule, height), min(cuts[r + 1] * granule, height)) for r in range(world)]


def refine_cost_density(density: Sequence[float] | None, bounds: Sequence[Tuple[int, int]], costs: Sequence[float], height: int,
                        granule: int = GRANULE) -> List[float]:
    """One step of learning the frame's cost per granule block from per-strip measurements: inside every strip the current estimate
    (uniform at first) is scaled so that the strip's total equals its measured cost. The shape an earlier partition taught survives
    inside the strips of the next one, so a few measure -> cut rounds with DIFFERENT boundaries converge on the real profile instead of
    re-flattening it every round (which is what `rebalance_bounds` alone does and why it stalls at a +-6 % spread on 8 strips)."""
    blocks = (height + granule - 1) // granule
    d = [1.0] * blocks if density is None else list(density)
    for (y0, y1), c in zip(bounds, costs):
        if y1 <= y0 or not c > 0.0:
            continue
        idx = range(y0 // granule, (y1 + granule - 1) // granule)
        cur = sum(d[b] for b in idx)
        if cur > 0.0:
            k = c / cur
            for b in idx:
                d[b] *= k
    return d


def bounds_from_density(density: Sequence[float], world: int, height: int, granule: int = GRANULE) -> List[Tuple[int, int]]:
    """Strip boundaries that cut the cumulative block cost into `world` equal parts (block granularity, at least one block per rank)."""
    blocks = len(density)
    if world == 1 or blocks <= world:
        return strip_bounds(height, world)
    total = sum(density)
    cuts, acc, b = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while b < blocks and acc + density[b] * 0.5 < target:
            acc += density[b]
            b += 1
        cut = min(max(b, cuts[-1] + 1), blocks - (world - r))
        cuts.append(cut)
        if cut != b:
            acc, b = sum(density[:cut]), cut
    cuts.append(blocks)
    return [(min(cuts[r] * granule, height), min(cuts[r + 1] * granule, height)) for r in range(world)]


def level_height(height: int, level: int) -> int:
    return height >> level


def level_rows(strip: Tuple[int, int], level: int, height: int) -> Tuple[int, int]:
    """Rows of `level` that the strip's kernels produce: [y0 >> l, ceil(y1 / 2^l)) clipped to the level (lgcu.h, lgcu_rows)."""
    y0, y1 = strip
    h = level_height(height, level)
    a, b = y0 >> level, (y1 + (1 << level) - 1) >> level
    return min(a, h), min(b, h)


def _owned(bounds: Sequence[Tuple[int, int]], level: int, height: int) -> List[Tuple[int, int]]:
    return [level_rows(s, level, height) for s in bounds]


def _fetch(image: str, level: int, need: Tuple[int, int], dst: int, owned: List[Tuple[int, int]]) -> List[Transfer]:
    """Transfers that bring rows `need` of (image, level) to rank `dst` from whoever owns them."""
    out = []
    a, b = need
    for src, (o0, o1) in enumerate(owned):
        if src == dst:
            continue
        lo, hi = max(a, o0), min(b, o1)
        if lo < hi:
            out.append(Transfer(image, level, lo, hi, src, dst))
    return out


def built_levels(width: int, height: int, mips: int = MIPS) -> int:
    n = 1
    for l in range(1, mips):
        if (width >> l) == 0 or (height >> l) == 0:
            break
        n += 1
    return n


def plan_chains(bounds: Sequence[Tuple[int, int]], width: int, height: int) -> List[Transfer]:
    """Before the chains stage: blur halos of chain levels 1..4 and the all-gather of the last strip-local level."""
    levels = built_levels(width, height)
    grid_levels = min(FRONT_LEVELS, levels - 1)
    plan: List[Transfer] = []
    for l in range(1, grid_levels + 1):
        owned = _owned(bounds, l, height)
        h = level_height(height, l)
        whole = l == grid_levels and levels > grid_levels + 1  # the tail CTA rebuilds levels above it from the whole level
        for dst, (o0, o1) in enumerate(owned):
            if o0 >= o1:
                continue
            need = (0, h) if whole else (max(o0 - BLUR_ABOVE, 0), min(o1 + BLUR_BELOW, h))
            for image in CHAINS:
                plan += _fetch(image, l, need, dst, owned)
    return plan


def plan_gather(bounds: Sequence[Tuple[int, int]], width: int, height: int) -> List[Transfer]:
    """Before the gather stage: REACH rows of the blurred levels 0..4 on either side of the strip."""
    levels = built_levels(width, height)
    grid_levels = min(FRONT_LEVELS, levels - 1)
    plan: List[Transfer] = []
    for l in range(0, grid_levels + 1):
        owned = _owned(bounds, l, height)
        h = level_height(height, l)
        for dst, (o0, o1) in enumerate(owned):
            if o0 >= o1:
                continue
            need = (max(o0 - GATHER_REACH, 0), min(o1 + GATHER_REACH, h))
            for image in BLURRED:
                plan += _fetch(image, l, need, dst, owned)
    return plan


def plan_present(bounds: Sequence[Tuple[int, int]], height: int, root: int = 0) -> List[Transfer]:
    """After the final stage: every swapchain strip to the presenting rank."""
    owned = _owned(bounds, 0, height)
    return _fetch("swapchain", 0, (0, height), root, owned)


def transfer_bytes(plan: Sequence[Transfer], pitches: Dict[Tuple[str, int], int]) -> int:
    return sum((t.row1 - t.row0) * pitches[(t.image, t.level)] for t in plan)
