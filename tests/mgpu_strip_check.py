"""2+ GPU check of the strip-sharded frame, launched by tests/test_multigpu_gpu.py (or by hand) as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_strip_check.py W H
Every rank renders the whole frame alone (reference) and its strip of the sharded frame; the strip images and, on rank 0, the
composited swapchain must equal the whole-frame result bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist

    from legitengine_b200 import abi, harness, multigpu, scene

    W, H = int(sys.argv[1]), int(sys.argv[2])
    transport = sys.argv[3] if len(sys.argv) > 3 else "nccl"
    uneven = len(sys.argv) > 4 and sys.argv[4] == "uneven"  # cost-aware (non-uniform) strip boundaries
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scene.make_scene(41, W, H)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        whole = harness.Renderer(W, H, stream=stream.cuda_stream)
        whole.upload_scene(sc)
        whole.render_frame(harness.MODE_FUSED, 0, abi.GI_DEFAULT)
        whole.sync()
        cls = multigpu.P2PStripRenderer if transport == "p2p" else multigpu.StripRenderer
        bounds = None
        if uneven:  # what sharding.rebalance_bounds produces for a frame whose lower strips are the expensive ones
            from legitengine_b200 import sharding

            bounds = sharding.rebalance_bounds(sharding.strip_bounds(H, world), [1.0 + 0.6 * r for r in range(world)], H)
            assert bounds != sharding.strip_bounds(H, world)
        sr = cls(W, H, rank, world, dist, stream=stream.cuda_stream, bounds=bounds)
        # only this rank's strip of the fragments is uploaded: the rest of the fragment buffer stays unwritten
        sr.renderer.upload_objects(sc.objects.ctypes.data, len(sc.objects))
        sr.renderer.upload_light_depth(np.ascontiguousarray(sc.shadow_map).ctypes.data, sc.shadow_map.shape[0])
        sr.upload_strip(sc.fragments.ctypes.data, sc.fragments.strides[0])
        for _ in range(3):  # several frames: warm allocations / cached views, and the frame-to-frame ack protocol of the p2p transport
            sr.render()
        if transport == "p2p":  # and once more from a CUDA graph (stages + flag kernels + peer copies)
            sr.capture()
            sr.replay()
            sr.replay()
        torch.cuda.synchronize()
        y0, y1 = sr.rows
        errors = []
        for name in ("directLight", "blurredDirectLight", "blurredDepthMoments", "indirectLight", "swapchain"):
            a, b = sr.renderer.download_image(name), whole.download_image(name)
            if not np.array_equal(a.level_bytes(0)[y0:y1], b.level_bytes(0)[y0:y1]):
                errors.append(f"{name} strip rows [{y0},{y1}) differ")
        a, b = sr.renderer.download_image("blurredDirectLight"), whole.download_image("blurredDirectLight")
        for l in range(5, 10):
            if (W >> l) and (H >> l) and not a.levels_equal(b, l):
                errors.append(f"blurredDirectLight level {l} differs")
        if rank == 0 and not sr.renderer.download_image("swapchain").levels_equal(whole.download_image("swapchain"), 0):
            errors.append("composited swapchain differs")
        flag = torch.tensor([len(errors)], device="cuda")
        dist.all_reduce(flag)
        print(f"rank {rank}/{world} rows [{y0},{y1}) received {sr.received_bytes} B/frame: {'OK' if not errors else errors}", flush=True)
        sr.close()
        whole.close()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
