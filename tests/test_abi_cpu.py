"""CPU tests of the drop-in boundary (no GPU, no compute): the C-ABI libraries load, export every symbol that
include/*.h declares, agree with the ctypes mirror on structure layout, and reject malformed calls with the documented
status codes before anything is launched."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from legitengine_b200 import abi, harness, images, passes

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: str, prefix: str):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_[a-z0-9_]+)\s*\(", text)))


def test_lgcu_exports_every_declared_symbol():
    lib = abi.load_lgcu()
    names = _declared("lgcu.h", "lgcu")
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"liblgcu.so does not export {n}"
    assert lib.lgcu_abi_version() == 1


def test_harness_exports_every_declared_symbol():
    lib = harness.load_harness()
    names = _declared("lgcu_harness.h", "lgh")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"liblegit_cuda.so does not export {n}"


def test_ubo_struct_sizes_match_the_reference():
    # sizes of the reference's tightly packed UBO structs
    assert C.sizeof(abi.GBufferBuilderData) == 136
    assert C.sizeof(abi.DrawCallData) == 96
    assert C.sizeof(abi.DirectLightingData) == 260
    assert C.sizeof(abi.MipLevelBuilderData) == 4
    assert C.sizeof(abi.BlurLayerBuilderData) == 20
    assert C.sizeof(abi.IndirectLightingData) == 144
    assert C.sizeof(abi.DenoiserData) == 148
    assert C.sizeof(abi.FinalGathererData) == 128
    assert abi.FRAGMENT_DTYPE.itemsize == 32
    assert C.sizeof(abi.LgcuImage) == 8 + 8 * 4 + 16 * 8 + 16 * 4


@pytest.mark.parametrize("fmt", [abi.FORMAT_R16G16B16A16_SFLOAT, abi.FORMAT_R32G32_SFLOAT, abi.FORMAT_D32_SFLOAT, abi.FORMAT_B8G8R8A8_SRGB])
@pytest.mark.parametrize("size", [(512, 512), (1920, 1080), (3840, 2160), (7680, 4320), (250, 141), (1, 1)])
def test_image_layout_matches_c_and_floor_halving(fmt, size):
    lib = abi.load_lgcu()
    W, Hh = size
    for mips in (1, 10):
        d = abi.LgcuImage()
        total = lib.lgcu_image_layout(C.byref(d), fmt, W, Hh, mips)
        py, py_total = images.make_layout(fmt, W, Hh, mips)
        assert total == py_total and total > 0
        assert list(d.levelOffset) == list(py.levelOffset) and list(d.levelPitch) == list(py.levelPitch)
        ts = lib.lgcu_format_texel_size(fmt)
        assert ts == abi.TEXEL_SIZE[fmt]
        for l in range(mips):
            w, h = images.mip_size(W, Hh, l)
            assert (w, h) == (W // (1 << l), Hh // (1 << l))  # LV/RenderGraph.h:373-374
            assert d.levelPitch[l] >= max(w, 1) * ts and d.levelPitch[l] % 128 == 0 and d.levelOffset[l] % 256 == 0


def test_mip_dims_known_answers():
    # 1080p chain
    dims = [images.mip_size(1920, 1080, l) for l in range(10)]
    assert dims == [(1920, 1080), (960, 540), (480, 270), (240, 135), (120, 67), (60, 33), (30, 16), (15, 8), (7, 4), (3, 2)]
    assert images.mip_size(3840, 2160, 9) == (7, 4) and images.mip_size(7680, 4320, 9) == (15, 8)
    assert passes.mip_levels_built(512, 512) == 10 and passes.mip_levels_built(250, 141) == 8 and passes.mip_levels_built(64, 36) == 6


def test_invalid_calls_are_rejected_without_a_device():
    lib = abi.load_lgcu()
    img = abi.LgcuImage()
    lib.lgcu_image_layout(C.byref(img), abi.FORMAT_R16G16B16A16_SFLOAT, 64, 64, 1)  # base stays NULL
    mp = abi.MipLevelBuilderData(0.0)
    assert lib.lgcu_mip_level(C.byref(mp), None, None, None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_mip_level(C.byref(mp), C.byref(img), C.byref(img), None, None) == abi.LGCU_ERR_INVALID_ARGUMENT  # null base
    assert b"null image" in lib.lgcu_last_error()
    depth_filter = abi.MipLevelBuilderData(1.0)  # FilterTypes::Depth is implemented: the same argument checks apply
    assert lib.lgcu_mip_level(C.byref(depth_filter), C.byref(img), C.byref(img), None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    ip = abi.InterleaveData((C.c_int32 * 4)(4, 4, 0, 0), (C.c_int32 * 4)(64, 64, 0, 0))
    assert lib.lgcu_deinterleave(C.byref(ip), C.byref(img), C.byref(img), None, None) == abi.LGCU_ERR_INVALID_ARGUMENT  # null base
    assert lib.lgcu_interleave(None, None, None, None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_debug_overlay(None, None, None, None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    d32 = abi.LgcuImage()
    lib.lgcu_image_layout(C.byref(d32), abi.FORMAT_D32_SFLOAT, 64, 64, 1)
    assert lib.lgcu_mip_level(C.byref(mp), C.byref(d32), C.byref(d32), None, None) == abi.LGCU_ERR_UNSUPPORTED_FORMAT
    bp = abi.BlurLayerBuilderData((C.c_int32 * 4)(64, 64, 0, 0), 2)
    assert lib.lgcu_blur_level(C.byref(bp), C.byref(img), C.byref(d32), None, None) == abi.LGCU_ERR_UNSUPPORTED_FORMAT
    assert lib.lgcu_mip_blur_chain(None, None, 2, None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_gi_gather(None, None, None, None, None, None, 0, None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_denoise(None, None, None, None, None, None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_format_texel_size(12345) == 0


def test_product_package_does_not_touch_the_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    pkg = ROOT / "legitengine_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cpp")):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "liboracle" not in text and "libref_spirv" not in text and "ssvgi_oracle" not in text, path


def test_missing_cuda_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(abi, "_lgcu", None)
    monkeypatch.setattr(abi, "LIB_DIR", tmp_path)
    with pytest.raises(abi.LibraryMissing):
        abi.load_lgcu()


def test_every_entry_point_is_documented():
    """Every function include/lgcu.h declares is accounted for in INTEGRATION.md or DESIGN.md (which reference interface it replaces,
    or why it has none)."""
    hdr = (ROOT / "include" / "lgcu.h").read_text()
    names = sorted(set(re.findall(r"\b(lgcu_[a-z0-9_]+)\s*\(", hdr)))
    docs = (ROOT / "INTEGRATION.md").read_text() + (ROOT / "DESIGN.md").read_text()
    missing = [n for n in names if n not in docs]
    assert len(names) >= 30 and not missing, missing
