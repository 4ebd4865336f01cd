"""Shared helpers for the parity tests: oracle frames, device upload/download, comparison metrics."""
from __future__ import annotations

import ctypes as C
import functools
from typing import Dict, Tuple

import numpy as np

from legitengine_b200 import abi, images, passes, scene
from oracle import loader

F16_EPS = 2.0 ** -10  # one fp16 ulp, relative (upper bound)


@functools.lru_cache(maxsize=8)
def oracle_frame(seed: int, width: int, height: int, denoise_radius: int = 0, n_boxes: int = 64, indirect_format: int = abi.FORMAT_R16G16B16A16_SFLOAT,
                 shadow_size: int = 1024):
    """Scene + the port oracle's images for the whole frame (cached: several tests share one frame)."""
    sc = scene.make_scene(seed, width, height, n_boxes=n_boxes, shadow_size=shadow_size)
    p = passes.make_params(width, height, sc.matrices, denoise_radius)
    fi = passes.FrameImages(width, height, images.HostImage, indirect_format=indirect_format, shadow_size=shadow_size)
    inp = passes.upload_inputs(fi, sc)
    passes.run_pass_list(loader.port(), fi, p, inp)
    return sc, p, fi


def device_frame_like(host: passes.FrameImages, copy=()) -> passes.FrameImages:
    """Device images with the same declaration as `host`; images named in `copy` are uploaded, the rest poisoned."""
    import torch

    dev = passes.FrameImages(host.width, host.height, images.DeviceImage, kwargs={"device": "cuda:0"}, indirect_format=host.indirect_format,
                             shadow_size=host.shadow_size)
    for name in copy:
        getattr(dev, name).tensor.copy_(torch.from_numpy(getattr(host, name).buf))
    return dev


def psnr(a: np.ndarray, b: np.ndarray, peak: float | None = None) -> float:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    mse = float(np.mean((a - b) ** 2))
    if mse == 0.0:
        return float("inf")
    if peak is None:
        peak = max(float(np.max(np.abs(b))), 1e-12)
    return 10.0 * np.log10(peak * peak / mse)


def compare_level(cuda_img: images.HostImage, ref_img: images.HostImage, level: int = 0) -> Dict[str, float]:
    a, b = cuda_img.level_f32(level), ref_img.level_f32(level)
    raw_equal = np.all(cuda_img.level_bytes(level) == ref_img.level_bytes(level), axis=2)
    diff = np.abs(a - b)
    finite = np.isfinite(a) & np.isfinite(b)
    diff = np.where(finite, diff, np.where(np.isnan(a) == np.isnan(b), 0.0, np.inf))
    tol = np.maximum(1e-3, F16_EPS * np.abs(b))
    return {
        "texels": raw_equal.size,
        "mismatched_texels": int((~raw_equal).sum()),
        "max_abs": float(diff.max()) if diff.size else 0.0,
        "outside_tol": int((diff > tol).any(axis=2).sum()),
        "psnr": psnr(a, b),
    }


def assert_bit_exact(cuda_img, ref_img, level=0, what=""):
    r = compare_level(cuda_img, ref_img, level)
    assert r["mismatched_texels"] == 0, f"{what} level {level}: {r}"


def assert_close(cuda_img, ref_img, level=0, what="", max_outside_frac=1e-4, min_psnr=60.0):
    """North-star tolerance: max-abs 1e-3 (or one fp16 ulp of the value where that is larger, for fp16-stored HDR radiance)
    on all but `max_outside_frac` of the texels, and PSNR >= 60 dB."""
    r = compare_level(cuda_img, ref_img, level)
    assert r["outside_tol"] <= max_outside_frac * r["texels"], f"{what} level {level}: {r}"
    assert r["psnr"] >= min_psnr, f"{what} level {level}: {r}"
    return r


# ---- committed golden fixtures (tests/golden/*.npz, generated from the reference arm by tests/golden/make_golden.py) ----
GOLDEN_DIR = __import__("pathlib").Path(__file__).resolve().parent / "golden"
GOLDEN_NAMES = sorted(p.stem for p in GOLDEN_DIR.glob("ssvgi_*.npz"))


BUNDLED_GOLDEN = "bundled_sponza_192x108"  # the reference's bundled sample scene (BASELINE configs[0]) at fixture size


def load_golden(name: str):
    """-> (scene, params, FrameImages on the host holding the reference arm's outputs, denoise radius)."""
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    seed, W, Hh, boxes, radius, shadow = (int(v) for v in z["meta"])
    m = scene.FrameMatrices(z["view"].copy(), z["proj"].copy(), z["light_view"].copy(), z["light_proj"].copy())
    frags = np.ascontiguousarray(z["fragments"]).view(abi.FRAGMENT_DTYPE).reshape(Hh, W)
    objs = np.ascontiguousarray(z["objects"]).view(abi.DRAW_CALL_DTYPE)
    sc = scene.Scene(W, Hh, seed, m, frags, objs, z["shadow_map"].copy())
    p = passes.make_params(W, Hh, m, radius)
    fi = passes.FrameImages(W, Hh, images.HostImage, shadow_size=shadow)
    fi.shadowMap.set_level(0, sc.shadow_map[..., None])
    for key in z.files:
        if key.startswith("img."):
            _, iname, level = key.split(".")
            getattr(fi, iname).level_bytes(int(level))[...] = z[key]
    return sc, p, fi, radius


# ---- mesh scenes: the oracle's rasteriser (oracle/raster_oracle.c) in front of the oracle's fragment passes ----
@functools.lru_cache(maxsize=4)
def oracle_mesh_frame(seed: int, width: int, height: int, camera_key: Tuple = None, n_boxes: int = 64, denoise_radius: int = 0):
    """Mesh + the scene it rasterises to (oracle rasteriser: fragments and shadow map) + the port oracle's frame images."""
    from legitengine_b200 import raster

    camera = dict(pos=camera_key[0], vert=camera_key[1], hor=camera_key[2]) if camera_key else None
    m = scene.frame_matrices(width, height, camera=camera)
    mesh = scene.scene_mesh(seed, n_boxes)
    port, ms = loader.port(), raster.host_mesh_desc(mesh)
    frags = np.zeros((height, width), dtype=abi.FRAGMENT_DTYPE)
    g = abi.GBufferBuilderData(abi.mat4(m.view), abi.mat4(m.proj), 0.0, 0.0)
    assert port.raster_gbuffer(C.byref(g), C.byref(ms), width, height, frags.ctypes.data, frags.strides[0], None) == 0
    shadow = np.zeros((scene.SHADOW_MAP_SIZE, scene.SHADOW_MAP_SIZE), dtype=np.float32)
    sp = abi.ShadowmapBuilderData(abi.mat4(m.light_view), abi.mat4(m.light_proj))
    assert port.raster_shadow_map(C.byref(sp), C.byref(ms), scene.SHADOW_MAP_SIZE, shadow.ctypes.data, shadow.strides[0]) == 0
    sc = scene.Scene(width, height, seed, m, frags, mesh.objects, shadow)
    p = passes.make_params(width, height, m, denoise_radius)
    fi = passes.FrameImages(width, height, images.HostImage)
    passes.run_pass_list(port, fi, p, passes.upload_inputs(fi, sc))
    return mesh, camera, sc, p, fi


# ---- full-size frames: whole-frame oracle for the cheap passes, row strips for the expensive ones ----
def oracle_frame_on_strips(seed: int, width: int, height: int, strips):
    """Oracle images of a big frame in seconds: K1..K4 (streaming passes) over the whole frame, K5..K7 (gather, denoise, final) only
    on the given row strips — the oracle's passes take the same lgcu_rows as the CUDA ones. Rows outside the strips stay poisoned."""
    sc = scene.make_scene(seed, width, height)
    p = passes.make_params(width, height, sc.matrices, 0)
    fi = passes.FrameImages(width, height, images.HostImage)
    inp = passes.upload_inputs(fi, sc)
    be = loader.port()
    passes.run_pass_list(be, fi, p, inp, stop_after="blur")
    v = lambda img, base=0, n=None: C.byref(img.view(base, n))
    for rows in strips:
        r = C.byref(abi.LgcuRows(*rows))
        assert be.gi_gather(C.byref(p.indirect), v(fi.blurredDirectLight), v(fi.blurredDepthMoments), v(fi.normal), v(fi.depthStencil), v(fi.indirectLight), 0, r) == 0
        assert be.denoise(C.byref(p.denoiser), v(fi.indirectLight), v(fi.normal), v(fi.depthMoments), v(fi.denoisedIndirectLight), r) == 0
        assert be.final_gather(C.byref(p.final), v(fi.directLight), v(fi.blurredDirectLight), v(fi.albedo), v(fi.denoisedIndirectLight), v(fi.swapchain), r) == 0
    return sc, p, fi


def rows_close(got: images.HostImage, want: images.HostImage, y0: int, y1: int, what: str, max_outside_frac: float = 1e-3, min_psnr: float = 60.0, level: int = 0):
    """assert_close restricted to rows [y0, y1) of one level (north-star tolerance: 1e-3 or one fp16 ulp of the value, PSNR >= 60 dB)."""
    a, b = got.level_f32(level)[y0:y1], want.level_f32(level)[y0:y1]
    tol = np.maximum(1e-3, F16_EPS * np.abs(b))
    outside = float((np.abs(a - b) > tol).any(axis=2).mean())
    assert outside <= max_outside_frac, f"{what} rows [{y0},{y1}): {outside:.2e} of the texels outside the tolerance"
    # PSNR against a peak of at least 1.0: a dark strip must not turn the 60 dB bar into something stricter than max-abs 1e-3
    peak = max(1.0, float(np.max(np.abs(b))))
    assert psnr(a, b, peak) >= min_psnr, f"{what} rows [{y0},{y1}): PSNR {psnr(a, b, peak):.1f} dB"


def check_big_frame(get_image, ref: passes.FrameImages, strips, width: int, height: int, whole_frame_images=True):
    """Parity of a full-size frame: `get_image(name)` returns the device frame's HostImage. G-buffer and the depth-moment chains are
    bit-exact over the whole frame (every level); the light chains within tolerance; gather / denoise / swapchain on `strips`."""
    levels = passes.mip_levels_built(width, height)
    if whole_frame_images:
        for name in ("normal", "depthStencil"):
            assert_bit_exact(get_image(name), getattr(ref, name), 0, name)
        for name in ("albedo", "emissive"):  # pow(colour, 2.2) per object: libm vs device pow, then fp16 rounding (as in test_gbuffer_resolve)
            r = compare_level(get_image(name), getattr(ref, name), 0)
            assert r["outside_tol"] == 0 and r["mismatched_texels"] <= 0.02 * r["texels"], (name, r)
        for name in ("depthMoments", "blurredDepthMoments"):
            got = get_image(name)
            for l in range(levels):
                assert_bit_exact(got, getattr(ref, name), l, name)
        for name in ("directLight", "blurredDirectLight"):
            got = get_image(name)
            for l in range(levels):
                w, h = images.mip_size(width, height, l)
                if w * h >= 4096:  # one texel of a tiny level is already more than the allowed fraction
                    rows_close(got, getattr(ref, name), 0, h, f"{name} level {l}", level=l)
    indirect, denoised, swap = get_image("indirectLight"), get_image("denoisedIndirectLight"), get_image("swapchain")
    for y0, y1 in strips:
        rows_close(indirect, ref.indirectLight, y0, y1, "indirectLight")
        rows_close(denoised, ref.denoisedIndirectLight, y0, y1, "denoisedIndirectLight")
        a = swap.level_raw(0)[y0:y1].astype(np.int32)
        b = ref.swapchain.level_raw(0)[y0:y1].astype(np.int32)
        assert (np.abs(a - b) > 1).mean() < 1e-3, f"swapchain rows [{y0},{y1})"
