#!/usr/bin/env python
"""Generates the committed golden fixtures of the SSVGI path FROM THE REFERENCE ARM (oracle/_ref/libref_spirv.so: the
reference's own shipped SPIR-V passes run through its vendored SPIRV-Cross C++ backend and GLM, built from
/root/reference by oracle/Makefile). Run in the development container only (the reference tree does not exist on the
GPU box):

    python tests/golden/make_golden.py

Each fixture is one small frame: the INPUTS of the path (fragment buffer, per-draw-call table, light depth map, the
four frame matrices) and every image the reference passes produce from them, all levels, as raw storage bytes.
The reference ships no golden vectors of its own for this path, so these — outputs of the reference's
arithmetic itself — are the pins: tests check the C port, and on the GPU the CUDA passes, against them.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from legitengine_b200 import abi, images, passes, scene  # noqa: E402
from oracle import loader  # noqa: E402

# name: (seed, W, H, boxes, denoise radius, shadow-map size)
FIXTURES = {
    "ssvgi_64x36_s3_r0": (3, 64, 36, 24, 0, 128),
    "ssvgi_50x29_s5_r2": (5, 50, 29, 16, 2, 64),   # ragged size: odd trailing rows/columns dropped by the mip chain
    "ssvgi_96x64_s9_r0": (9, 96, 64, 64, 0, 256),
}


def aux_fixture() -> None:
    """The passes either side of the hot path through the reference's own deinterleave.frag.spv,
    interleave.frag.spv, mipLevelBuilder.frag.spv (Depth branch) and debugRenderer.frag.spv: inputs and outputs as raw storage bytes."""
    import ctypes as C

    from tests import aux_helpers as A

    ref = loader.ref()
    out = {}
    for k, (fmt, W, H, gx, gy) in enumerate([(abi.FORMAT_R16G16B16A16_SFLOAT, 50, 29, 4, 4), (abi.FORMAT_R32G32_SFLOAT, 48, 32, 3, 2),
                                             (abi.FORMAT_R32G32B32A32_SFLOAT, 17, 9, 4, 4)]):
        src = A.random_image(fmt, W, H, seed=100 + k)
        p = A.interleave_params(W, H, gx, gy)
        out[f"interleave{k}.meta"] = np.array([fmt, W, H, gx, gy], dtype=np.int64)
        out[f"interleave{k}.src"] = np.ascontiguousarray(src.level_bytes(0))
        out[f"interleave{k}.deinterleaved"] = np.ascontiguousarray(A.run_pass(ref.deinterleave, p, src).level_bytes(0))
        out[f"interleave{k}.interleaved"] = np.ascontiguousarray(A.run_pass(ref.interleave, p, src).level_bytes(0))
    for k, (fmt, W, H) in enumerate([(abi.FORMAT_R16G16B16A16_SFLOAT, 33, 17), (abi.FORMAT_R32G32_SFLOAT, 64, 36)]):
        src = A.random_depth_range_image(fmt, W, H, seed=200 + k)
        img = images.HostImage(fmt, W, H, 2)
        img.level_bytes(0)[...] = src.level_bytes(0)
        mp = abi.MipLevelBuilderData(1.0)
        assert ref.mip_level(C.byref(mp), C.byref(img.view(0, 1)), C.byref(img.view(1, 1)), None) == 0
        out[f"depthmip{k}.meta"] = np.array([fmt, W, H], dtype=np.int64)
        out[f"depthmip{k}.src"] = np.ascontiguousarray(img.level_bytes(0))
        out[f"depthmip{k}.level1"] = np.ascontiguousarray(img.level_bytes(1))
    W, H = 96, 54
    target = A.random_image(abi.FORMAT_B8G8R8A8_SRGB, W, H, seed=300)
    out["overlay.meta"] = np.array([W, H, 4], dtype=np.int64)
    out["overlay.target_before"] = np.ascontiguousarray(target.level_bytes(0)).copy()
    for k, quad in enumerate(A.debug_tiles(4)):
        src = A.random_image(abi.FORMAT_R16G16B16A16_SFLOAT, W, H, seed=310 + k, lo=0.0, hi=1.5)
        out[f"overlay.src{k}"] = np.ascontiguousarray(src.level_bytes(0))
        out[f"overlay.quad{k}"] = np.array(list(quad.minmax), dtype=np.float32)
        assert ref.debug_overlay(C.byref(quad), C.byref(src.view()), C.byref(target.view()), None) == 0
    out["overlay.target_after"] = np.ascontiguousarray(target.level_bytes(0))
    path = HERE / "aux_passes.npz"
    np.savez_compressed(path, **out)
    print(f"{path.name}: {path.stat().st_size / 1024:.0f} KiB")


def bundled_scene_fixture(W: int = 192, H: int = 108, shadow: int = 256) -> None:
    """BASELINE configs[0], scaled down to fixture size: the reference's bundled sample scene (bin/data/Scenes/SponzaScene.json +
    OBJ meshes, loaded where they lie by tests/bundled_scene.py), rasterised by the oracle's rasteriser (rule R) from the
    application's default camera / light, then through the reference's own SPIR-V passes. Same layout as the ssvgi_* fixtures."""
    import ctypes as C

    from legitengine_b200 import raster
    from tests import bundled_scene as B

    mesh = B.load_bundled_scene()
    m = scene.frame_matrices(W, H)
    port, ms = loader.port(), raster.host_mesh_desc(mesh)
    frags = np.zeros((H, W), dtype=abi.FRAGMENT_DTYPE)
    g = abi.GBufferBuilderData(abi.mat4(m.view), abi.mat4(m.proj), 0.0, 0.0)
    assert port.raster_gbuffer(C.byref(g), C.byref(ms), W, H, frags.ctypes.data, frags.strides[0], None) == 0
    depth = np.zeros((shadow, shadow), dtype=np.float32)
    sp = abi.ShadowmapBuilderData(abi.mat4(m.light_view), abi.mat4(m.light_proj))
    assert port.raster_shadow_map(C.byref(sp), C.byref(ms), shadow, depth.ctypes.data, depth.strides[0]) == 0
    sc = scene.Scene(W, H, 0, m, frags, mesh.objects, depth)
    p = passes.make_params(W, H, m, 0)
    fi = passes.FrameImages(W, H, images.HostImage, shadow_size=shadow)
    passes.run_pass_list(loader.ref(), fi, p, passes.upload_inputs(fi, sc))
    out = {
        "meta": np.array([0, W, H, 0, 0, shadow], dtype=np.int64),
        "triangles": np.array([mesh.triangle_count], dtype=np.int64),
        "fragments": frags.view(np.uint8).reshape(H, W * 32).copy(),
        "objects": mesh.objects.view(np.uint8).reshape(-1).copy(),
        "shadow_map": depth.copy(),
        "view": m.view, "proj": m.proj, "light_view": m.light_view, "light_proj": m.light_proj,
    }
    for iname, img in fi.items():
        if iname == "shadowMap":
            continue
        for l in range(img.mips):
            w, h = img.level_size(l)
            if w > 0 and h > 0:
                out[f"img.{iname}.{l}"] = np.ascontiguousarray(img.level_bytes(l))
    path = HERE / f"bundled_sponza_{W}x{H}.npz"
    np.savez_compressed(path, **out)
    print(f"{path.name}: {path.stat().st_size / 1024:.0f} KiB")


def main() -> None:
    aux_fixture()
    bundled_scene_fixture()
    ref = loader.ref()
    for name, (seed, W, H, boxes, radius, shadow) in FIXTURES.items():
        sc = scene.make_scene(seed, W, H, n_boxes=boxes, shadow_size=shadow)
        p = passes.make_params(W, H, sc.matrices, radius)
        fi = passes.FrameImages(W, H, images.HostImage, shadow_size=shadow)
        passes.run_pass_list(ref, fi, p, passes.upload_inputs(fi, sc))
        out = {
            "meta": np.array([seed, W, H, boxes, radius, shadow], dtype=np.int64),
            "fragments": sc.fragments.view(np.uint8).reshape(H, W * 32).copy(),
            "objects": sc.objects.view(np.uint8).reshape(-1).copy(),
            "shadow_map": sc.shadow_map.copy(),
            "view": sc.matrices.view, "proj": sc.matrices.proj, "light_view": sc.matrices.light_view, "light_proj": sc.matrices.light_proj,
        }
        for iname, img in fi.items():
            if iname == "shadowMap":
                continue
            for l in range(img.mips):
                w, h = img.level_size(l)
                if w > 0 and h > 0:
                    out[f"img.{iname}.{l}"] = np.ascontiguousarray(img.level_bytes(l))
        path = HERE / f"{name}.npz"
        np.savez_compressed(path, **out)
        print(f"{path.name}: {path.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
