#!/usr/bin/env python
"""Per-pass GPU times of the fused frame (mean of N profiled frames), or of the gather stage on a row strip, as one JSON line:
    python scripts/pass_times.py [W H [row0 row1]]
Used for A/B runs of switch-guarded kernel changes (LGCU_FAST_SRGB, LGCU_FRONT_BLOCKS, LGCU_RASTER_PATH: the switches that exist are
echoed into the line). Round 1 also A/B-ed gather variants through switches that have since been removed; the "variant" / "slices" /
"order" / "specialised" keys of profiles/r01*.jsonl refer to those (profiles/README.md says what each one was)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from legitengine_b200 import abi, harness, scene  # noqa: E402

GI = abi.GI_STRICT if os.environ.get("LGCU_PASS_TIMES_STRICT") else abi.GI_DEFAULT  # shader-order gather kernel (ncu: algorithmic FP32 count)
RADIUS = int(os.environ.get("LGCU_PASS_TIMES_RADIUS", "0"))  # denoiser radius (0 or 2)
W, H = (3840, 2160) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
ROWS = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) >= 5 else None  # time the gather stage on a row strip only
m = scene.frame_matrices(W, H)
frags = scene.scene_fragments(0xC0FFEE, W, H, m)
objects = scene.scene_objects(0xC0FFEE)
shadow = scene.scene_shadow_map(0xC0FFEE, m)
r = harness.Renderer(W, H)
r.upload_fragments(frags.ctypes.data, frags.strides[0])
r.upload_objects(objects.ctypes.data, len(objects))
r.upload_light_depth(shadow.ctypes.data, 1024)
r.sync()
for _ in range(3):
    r.render_frame(harness.MODE_FUSED, RADIUS, GI)
r.sync()
acc, n = {}, 10
if ROWS is None:
    for _ in range(n):
        r.render_frame(harness.MODE_FUSED, RADIUS, GI, profile=True)
        r.sync()
        for name, ms in r.profile():
            acc[name] = acc.get(name, 0.0) + ms / n
else:
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream()
    r.close()
    r = harness.Renderer(W, H, stream=stream.cuda_stream)
    r.upload_fragments(frags.ctypes.data, frags.strides[0])
    r.upload_objects(objects.ctypes.data, len(objects))
    r.upload_light_depth(shadow.ctypes.data, 1024)
    r.render_frame(harness.MODE_FUSED, RADIUS, GI)
    r.sync()
    with torch.cuda.stream(stream):
        for i in range(n + 3):
            if i == 3:
                ev0.record(stream)
            r.render_stages(harness.STAGE_GATHER, rows=ROWS)
        ev1.record(stream)
    r.sync()
    acc["GatherStage(rows %d..%d)" % ROWS] = ev0.elapsed_time(ev1) / n
print(json.dumps({"gather_lock": os.environ.get("LGCU_GATHER_LOCK", "default"), "fast_srgb": os.environ.get("LGCU_FAST_SRGB", "default"), "front_blocks": os.environ.get("LGCU_FRONT_BLOCKS", "default"), "size": [W, H], "pass_ms": {k: round(v, 4) for k, v in acc.items()}}))
