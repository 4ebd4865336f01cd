"""Helpers of the §8f rank 3 / 4 pass tests (interleaved rendering, Depth-filter mips, debug overlay): seeded images, parameter
blocks, and the tile layout loop of DebugRenderer::RenderImageViews restated in numpy fp32."""
from __future__ import annotations

import ctypes as C

import numpy as np

from legitengine_b200 import abi, images

CHANNELS = {abi.FORMAT_R16G16B16A16_SFLOAT: 4, abi.FORMAT_R32G32_SFLOAT: 2, abi.FORMAT_R32G32B32A32_SFLOAT: 4, abi.FORMAT_B8G8R8A8_SRGB: 4, abi.FORMAT_D32_SFLOAT: 1}


def random_image(fmt: int, width: int, height: int, seed: int, mips: int = 1, lo: float = -4.0, hi: float = 60.0) -> images.HostImage:
    """Finite pseudo-random texels in level 0 (no NaN: a NaN's payload is not preserved through the shader's vec4)."""
    rng = np.random.default_rng(seed)
    img = images.HostImage(fmt, width, height, mips)
    if fmt == abi.FORMAT_B8G8R8A8_SRGB:
        img.level_bytes(0)[...] = rng.integers(0, 256, size=img.level_bytes(0).shape, dtype=np.uint8)
    else:
        img.set_level(0, rng.uniform(lo, hi, size=(height, width, 4)).astype(np.float32)[..., : CHANNELS[fmt]])
    return img


def random_depth_range_image(fmt: int, width: int, height: int, seed: int) -> images.HostImage:
    """(min depth, max depth >= min, density, unused) texels, the layout the Depth filter reduces."""
    rng = np.random.default_rng(seed)
    lo = rng.uniform(0.1, 50.0, size=(height, width)).astype(np.float32)
    v = np.stack([lo, lo + rng.uniform(0.01, 5.0, size=lo.shape).astype(np.float32), rng.uniform(0.0, 1.0, size=lo.shape).astype(np.float32),
                  np.zeros_like(lo)], axis=-1)
    img = images.HostImage(fmt, width, height, 1)
    img.set_level(0, v[..., : CHANNELS[fmt]])
    return img


def interleave_params(width: int, height: int, gx: int, gy: int) -> abi.InterleaveData:
    return abi.InterleaveData((C.c_int32 * 4)(gx, gy, 0, 0), (C.c_int32 * 4)(width, height, 0, 0))


def run_pass(fn, params, src: images.HostImage, rows=None) -> images.HostImage:
    dst = images.HostImage(src.format, src.desc.width, src.desc.height, 1)
    st = fn(C.byref(params), C.byref(src.view()), C.byref(dst.view()), C.byref(rows) if rows is not None else None)
    assert st == 0, st
    return dst


def equal_nan_aware(a: images.HostImage, b: images.HostImage, level: int) -> bool:
    """Bit-identical texels, except that two NaNs compare equal (0/0 has no defined payload)."""
    fa, fb = a.level_f32(level), b.level_f32(level)
    same = np.all(a.level_bytes(level) == b.level_bytes(level), axis=2)
    both_nan_only = np.all((fa == fb) | (np.isnan(fa) & np.isnan(fb)), axis=2)
    return bool(np.all(same | both_nan_only))


from legitengine_b200.passes import debug_tiles  # noqa: E402,F401  (host-side tile layout of DebugRenderer.h:27-57)


# ---- committed golden fixture of these passes (tests/golden/aux_passes.npz: outputs of the reference's own SPIR-V, make_golden.py) ----
def load_aux_golden():
    from pathlib import Path

    return np.load(Path(__file__).resolve().parent / "golden" / "aux_passes.npz")


def image_from_bytes(fmt: int, raw: np.ndarray, mips: int = 1) -> images.HostImage:
    h, w = raw.shape[:2]
    img = images.HostImage(fmt, w, h, mips)
    img.level_bytes(0)[...] = raw
    return img


def golden_cases(z):
    """Yields (name, run(backend) -> HostImage-or-bytes comparison) for every case of the aux fixture; `backend` has the pass
    callables of include/lgcu.h minus the stream (an oracle, or an adapter over the CUDA library working on device copies)."""
    k = 0
    while f"interleave{k}.meta" in z:
        fmt, W, H, gx, gy = (int(v) for v in z[f"interleave{k}.meta"])
        yield ("interleave", k, fmt, (W, H, gx, gy))
        k += 1
    k = 0
    while f"depthmip{k}.meta" in z:
        fmt, W, H = (int(v) for v in z[f"depthmip{k}.meta"])
        yield ("depthmip", k, fmt, (W, H))
        k += 1
    yield ("overlay", 0, abi.FORMAT_B8G8R8A8_SRGB, tuple(int(v) for v in z["overlay.meta"]))
