#!/usr/bin/env python
"""BASELINE.json configs[0] and configs[1] on one B200, one JSON line each (bench.py covers configs[2..4]):

  bundled512 : the reference's bundled sample scene (SponzaScene.json, 364 157 triangles; oracle/_ref/bundled_sponza_mesh.npz, packed by
               oracle/make_bundled_mesh.py where /root/reference exists) rendered 512x512 from its vertex / index buffers: ShadowPass +
               GBufferRasterPass on the device, then the fused frame. value = scene resident (CUDA-graph replay), e2e = scene uploaded from
               pinned host memory and swapchain downloaded every step. CPU column: the reference's SPIR-V passes on the same 512x512 frame.
  k1k2_1080p : 1920x1080 G-buffer resolve + shadow-mapped direct lighting through the C ABI: the two passes separately, their fusion
               (lgcu_gbuffer_direct_light) and the frame-front kernel the fused frame actually runs (K1 + K2 + level-0 blur copies + mips
               1..4); three image sets are cycled so that every launch reads and writes memory that is not in L2 (3 x 141 MB > 126 MB).
Timing: CUDA events on the launching stream, warm-up, then N launches.   python scripts/bench_configs.py [bundled512] [k1k2_1080p]"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from legitengine_b200 import abi, harness, images, passes, scene  # noqa: E402

PEAK_GBS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timed(stream, fn, n, warm=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def bundled512(steps=100):
    path = ROOT / "oracle" / "_ref" / "bundled_sponza_mesh.npz"
    if not path.exists():
        print(json.dumps({"config": "bundled512", "unavailable": "oracle/_ref/bundled_sponza_mesh.npz not generated (needs /root/reference at build time)"}))
        return
    mesh = scene.load_packed_mesh(path)
    W = H = 512
    pinned = {}
    for name in ("vertices", "indices", "draws", "objects"):
        a = getattr(mesh, name)
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype)
        v[...] = a
        pinned[name] = (t, v)
    mesh_pinned = scene.Mesh(*(pinned[n][1] for n in ("vertices", "indices", "draws", "objects")))
    swap = torch.empty((H, W * 4), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.Stream()
    r = harness.Renderer(W, H, stream=stream.cuda_stream)
    r.upload_mesh(mesh_pinned)
    r.render_frame(harness.MODE_FUSED, 0, abi.GI_DEFAULT)
    r.sync()
    pass_ms = {}
    for _ in range(5):
        r.render_frame(harness.MODE_FUSED, 0, abi.GI_DEFAULT, profile=True)
        r.sync()
        for n, ms in r.profile():
            pass_ms[n] = pass_ms.get(n, 0.0) + ms / 5
    r.capture_frame(harness.MODE_FUSED, 0, abi.GI_DEFAULT)
    ms = timed(stream, r.replay_frame, steps)

    def e2e():
        r.upload_mesh(mesh_pinned)
        r.replay_frame()
        r.download_swapchain(swap.data_ptr(), W * 4)

    e2e_ms = timed(stream, e2e, steps)
    # CPU column: the reference's own SPIR-V passes on the fragments the oracle's rasteriser produces (checker-side code, timed after the GPU legs)
    from oracle import frames as OF
    from oracle import loader

    sc, p, ref = OF.oracle_frame_from_mesh(mesh, W, H)
    be = loader.ref() if loader.have_ref() else loader.port()
    be.set_num_threads(len(__import__("os").sched_getaffinity(0)))
    fi = passes.FrameImages(W, H, images.HostImage)
    inp = passes.upload_inputs(fi, sc)
    passes.run_pass_list(be, fi, p, inp)
    t0 = time.perf_counter()
    for _ in range(3):
        passes.run_pass_list(be, fi, p, inp)
    cpu_s = (time.perf_counter() - t0) / 3
    npx = W * H
    print(json.dumps({
        "config": "BASELINE configs[0]: 512x512 render of the bundled sample scene (SponzaScene.json, %d triangles, %d draws) from its vertex / index buffers" % (mesh.triangle_count, len(mesh.draws)),
        "metric": "full_gi_frame_mpix_per_s", "value": npx / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": ms, "frames_per_s": 1e3 / ms, "steps": steps,
        "e2e": {"value": npx / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(sum(pinned[n][0].numel() for n in pinned)), "d2h_bytes_per_step": W * H * 4},
        "pass_ms": {k: round(v, 4) for k, v in pass_ms.items()},
        "cpu_baseline": {"value": npx / cpu_s / 1e6, "unit": "Mpix/s", "ms_per_frame": cpu_s * 1e3, "kind": be.kind, "cores": int(be.num_threads()),
                         "sample": "the reference's SPIR-V fragment passes K1..K7 on the same 512x512 frame (the reference rasterises with Vulkan; not timed here)"},
    }), flush=True)
    r.close()


def k1k2_1080p(steps=300):
    W, H = 1920, 1080
    npx = W * H
    sc = scene.make_scene(0xC0FFEE, W, H)
    p = passes.make_params(W, H, sc.matrices, 0)
    cu = passes.CudaPasses()
    lib = cu.lib
    v = lambda img, base=0, n=None: C.byref(img.view(base, n))
    sets = []
    for _ in range(3):
        dev = passes.FrameImages(W, H, images.DeviceImage, kwargs={"device": "cuda:0"})
        sets.append((dev, passes.upload_inputs(dev, sc)))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    k = [0]

    def nxt():
        k[0] += 1
        return sets[k[0] % 3]

    def resolve():
        dev, inp = nxt()
        cu.gbuffer_resolve(C.byref(p.gbuffer), inp["objects_ptr"], inp["n_objects"], inp["fragments_ptr"], inp["pitch"], C.byref(p.clear), v(dev.albedo), v(dev.emissive), v(dev.normal),
                           v(dev.depthMoments, 0, 1), v(dev.depthStencil), None)

    def light():
        dev, inp = nxt()
        cu.direct_light(C.byref(p.light), v(dev.albedo), v(dev.emissive), v(dev.normal), v(dev.depthStencil), v(dev.shadowMap), v(dev.directLight, 0, 1), None)

    def fused():
        dev, inp = nxt()
        cu.gbuffer_direct_light(C.byref(p.gbuffer), C.byref(p.light), inp["objects_ptr"], inp["n_objects"], inp["fragments_ptr"], inp["pitch"], C.byref(p.clear), v(dev.albedo), v(dev.emissive),
                                v(dev.normal), v(dev.depthMoments, 0, 1), v(dev.depthStencil), v(dev.shadowMap), v(dev.directLight, 0, 1), None)

    def front():
        dev, inp = nxt()
        cu.frame_front(C.byref(p.gbuffer), C.byref(p.light), inp["objects_ptr"], inp["n_objects"], inp["fragments_ptr"], inp["pitch"], C.byref(p.clear), v(dev.albedo), v(dev.emissive), v(dev.normal),
                       v(dev.depthMoments), v(dev.depthStencil), v(dev.shadowMap), v(dev.directLight), v(dev.blurredDirectLight), v(dev.blurredDepthMoments), None)

    for fn in (resolve, light, fused, front):
        fn()
    torch.cuda.synchronize()
    t = {name: timed(stream, fn, steps) for name, fn in (("K1 gbuffer_resolve", resolve), ("K2 direct_light", light), ("K1+K2 gbuffer_direct_light", fused), ("frame_front (K1+K2+blur0+mips1..4)", front))}
    alg = 68.0 * npx + 36.0 * npx + 4 * 1024 * 1024  # pass-granular algorithmic bytes of K1 + K2
    fused_bytes = (32.0 + 36.0 + 8.0) * npx + 4 * 1024 * 1024  # what one fused pass must move: fragments in, five G-buffer images + directLight out
    sep = t["K1 gbuffer_resolve"] + t["K2 direct_light"]
    print(json.dumps({
        "config": "BASELINE configs[1]: 1920x1080 G-buffer resolve + shadow-mapped direct lighting on 1 B200",
        "ms": {k_: round(v_, 5) for k_, v_ in t.items()},
        "algorithmic_bytes_k1_k2": alg, "roofline_us_at_measured_peak": alg / (PEAK_GBS * 1e9) * 1e6, "peak_gbs": PEAK_GBS,
        "separate_passes": {"ms": sep, "achieved_gbs": alg / (sep * 1e-3) / 1e9, "frac": alg / (sep * 1e-3) / 1e9 / PEAK_GBS},
        "fused": {"ms": t["K1+K2 gbuffer_direct_light"], "achieved_gbs_pass_granular": alg / (t["K1+K2 gbuffer_direct_light"] * 1e-3) / 1e9,
                  "frac_pass_granular": alg / (t["K1+K2 gbuffer_direct_light"] * 1e-3) / 1e9 / PEAK_GBS, "bytes_a_fused_pass_moves": fused_bytes,
                  "frac_fused_bytes": fused_bytes / (t["K1+K2 gbuffer_direct_light"] * 1e-3) / 1e9 / PEAK_GBS},
        "value_mpix_per_s_fused": npx / (t["K1+K2 gbuffer_direct_light"] * 1e-3) / 1e6,
        "l2": "three image sets cycled (3 x 141 MB > 126 MB L2)",
    }), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["bundled512", "k1k2_1080p"]
    if "k1k2_1080p" in which:
        k1k2_1080p()
    if "bundled512" in which:
        bundled512()
