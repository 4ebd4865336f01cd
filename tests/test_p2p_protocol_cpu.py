"""Model check of the peer-to-peer frame protocol of the strip-sharded renderer (legitengine_b200/multigpu.py: P2PStripRenderer,
csrc/k_p2p.cu) — CPU only, no CUDA.

Every rank's CUDA stream is a sequential program of stage kernels, flag signals, flag waits and peer copies; the GPUs run these
programs concurrently with no other ordering between them. The test replays the programs of all ranks under many random
interleavings (each operation split into a begin and an end event, so operations of different ranks overlap) and checks, for
several frames in a row as under CUDA-graph replay:

  * no deadlock: every schedule runs to completion;
  * read-after-write: a pull or a push only ever reads rows of the frame it belongs to;
  * write-after-read / write-after-write: nobody overwrites rows a peer is still reading, or writes rows someone else is writing.

Who signals whom is derived from the same transfer plans the renderer uses (legitengine_b200/sharding.py). A mutated protocol
(the ACK wait removed) must be caught, which shows the checker can see the hazard the ACK exists for."""
import random

import pytest

from legitengine_b200 import sharding

FRONT, CHAINS, DELIVERED, ACK, FREE = range(5)


def _neighbours(bounds, width, height, root=0):
    pc, pg, pp = sharding.plan_chains(bounds, width, height), sharding.plan_gather(bounds, width, height), sharding.plan_present(bounds, height, root)
    world = len(bounds)
    info = []
    for r in range(world):
        info.append({
            "chain_pullers": {t.dst for t in pc if t.src == r}, "chain_sources": {t.src for t in pc if t.dst == r},
            "gather_pullers": {t.dst for t in pg if t.src == r}, "gather_sources": {t.src for t in pg if t.dst == r},
            "pusher": r in {t.src for t in pp}, "pushers": {t.src for t in pp},
        })
    return info


def _program(rank, info, root, frames, ack_wait=True, direct_present=False):
    """The operation list of one rank's stream for `frames` frames, in the order P2PStripRenderer.render() enqueues it.
    ops: ("bump",) | ("signal", stage, targets) | ("wait", stage, writers, lag) | ("work", name, reads, writes)
    resources are (owner rank, name); reads carry the frame the data must belong to."""
    me = info[rank]
    ops = []
    for _ in range(frames):
        ops.append(("bump",))
        if rank == root:
            ops.append(("signal", FREE, me["pushers"]))
        if ack_wait:
            ops.append(("wait", ACK, me["chain_pullers"] | me["gather_pullers"], 1))
        ops.append(("work", "front", [], [(rank, "chain"), (rank, "blurred0")]))
        ops.append(("signal", FRONT, me["chain_pullers"]))
        ops.append(("wait", FRONT, me["chain_sources"], 0))
        ops.append(("work", "pull_chains", [(s, "chain") for s in me["chain_sources"]], [(rank, "chain_halo")]))
        ops.append(("work", "chains", [(rank, "chain"), (rank, "chain_halo")], [(rank, "blurred")]))
        ops.append(("signal", CHAINS, me["gather_pullers"]))
        ops.append(("wait", CHAINS, me["gather_sources"], 0))
        if direct_present and me["pusher"]:  # fused exchange: FREE is waited for together with the CHAINS flags, before the pull
            ops.append(("wait", FREE, {root}, 0))
        ops.append(("work", "pull_gather", [(s, "blurred") for s in me["gather_sources"]] + [(s, "blurred0") for s in me["gather_sources"]], [(rank, "blurred_halo")]))
        ops.append(("signal", ACK, me["chain_sources"] | me["gather_sources"]))
        if direct_present and me["pusher"]:  # the final pass writes the strip straight into the root's swapchain image
            ops.append(("work", "gather_final", [(rank, "blurred"), (rank, "blurred0"), (rank, "blurred_halo")], [(root, f"swap_from_{rank}")]))
            ops.append(("signal", DELIVERED, {root}))
        else:
            ops.append(("work", "gather_final", [(rank, "blurred"), (rank, "blurred0"), (rank, "blurred_halo")], [(rank, "swap")]))
            if me["pusher"]:
                ops.append(("wait", FREE, {root}, 0))
                ops.append(("work", "push", [(rank, "swap")], [(root, f"swap_from_{rank}")]))
                ops.append(("signal", DELIVERED, {root}))
        if rank == root:
            ops.append(("wait", DELIVERED, me["pushers"], 0))
            ops.append(("work", "present", [(root, "swap")] + [(root, f"swap_from_{p}") for p in me["pushers"]], []))
    return ops


class Hazard(AssertionError):
    pass


def _simulate(world, bounds, width, height, frames, seed, ack_wait=True, root=0, direct_present=False):
    rng = random.Random(seed)
    info = _neighbours(bounds, width, height, root)
    programs = [_program(r, info, root, frames, ack_wait, direct_present) for r in range(world)]
    pc = [0] * world            # next op of every rank
    in_flight = [None] * world  # the "work" op a rank has begun and not yet ended
    frame = [0] * world         # device-side frame counter of every rank
    flags = {}                  # (owner rank, stage, writer rank) -> value
    version = {}                # resource -> frame of its contents
    readers, writer = {}, {}    # resource -> set of ranks reading / rank writing

    def blocked(r):
        op = programs[r][pc[r]]
        if op[0] == "wait":
            _, stage, writers, lag = op
            return any(flags.get((r, stage, w), 0) < frame[r] - lag for w in writers)
        return False

    steps = 0
    while any(pc[r] < len(programs[r]) or in_flight[r] for r in range(world)):
        runnable = [r for r in range(world) if in_flight[r] or (pc[r] < len(programs[r]) and not blocked(r))]
        if not runnable:
            raise Hazard(f"deadlock: {[(r, programs[r][pc[r]]) for r in range(world) if pc[r] < len(programs[r])]}")
        r = rng.choice(runnable)
        steps += 1
        if in_flight[r]:  # end of a work op
            _, name, reads, writes = in_flight[r]
            for res in reads:
                readers[res].discard(r)
            for res in writes:
                writer[res] = None
                version[res] = frame[r]
            in_flight[r] = None
            continue
        op = programs[r][pc[r]]
        pc[r] += 1
        if op[0] == "bump":
            frame[r] += 1
        elif op[0] == "signal":
            for target in op[2]:
                flags[(target, op[1], r)] = frame[r]
        elif op[0] == "work":
            _, name, reads, writes = op
            for res in reads:
                if writer.get(res) is not None:
                    raise Hazard(f"rank {r} {name} frame {frame[r]} reads {res} while rank {writer[res]} writes it")
                if res[0] != r and version.get(res) != frame[r]:
                    raise Hazard(f"rank {r} {name} frame {frame[r]} reads {res} of frame {version.get(res)}")
                readers.setdefault(res, set()).add(r)
            for res in writes:
                if readers.get(res):
                    raise Hazard(f"rank {r} {name} frame {frame[r]} overwrites {res} while ranks {sorted(readers[res])} read it")
                if writer.get(res) is not None:
                    raise Hazard(f"rank {r} {name} writes {res} while rank {writer[res]} writes it")
                writer[res] = r
            in_flight[r] = op
    return steps


@pytest.mark.parametrize("direct_present", [False, True])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_frame_protocol_has_no_deadlock_and_no_data_hazard(world, direct_present):
    """Both composites: the strip pushed by a copy after the final pass, and (default) the final pass writing it straight into the
    presenting GPU's swapchain image."""
    width, height = 7680, 4320
    even = sharding.strip_bounds(height, world)
    uneven = sharding.rebalance_bounds(even, [1.0 + 0.3 * ((r * 5) % 4) for r in range(world)], height)
    for bounds in (even, uneven):
        for seed in range(60):
            assert _simulate(world, bounds, width, height, frames=4, seed=seed, direct_present=direct_present) > 0


def test_small_frame_where_strips_reach_beyond_their_neighbours():
    """A low frame on many ranks: halos come from ranks that are not adjacent, and the coarse level is gathered from everyone."""
    width, height, world = 640, 160, 8
    bounds = sharding.strip_bounds(height, world)
    info = _neighbours(bounds, width, height)
    assert any(len(i["gather_sources"]) > 2 for i in info)
    for seed in range(60):
        _simulate(world, bounds, width, height, frames=3, seed=seed)


def test_checker_catches_the_hazard_the_ack_exists_for():
    """Without the ACK wait a fast rank starts the next frame's front stage while a slow neighbour still pulls its rows."""
    width, height, world = 7680, 4320, 4
    bounds = sharding.strip_bounds(height, world)
    caught = 0
    for seed in range(200):
        try:
            _simulate(world, bounds, width, height, frames=3, seed=seed, ack_wait=False)
        except Hazard:
            caught += 1
    assert caught > 0
