"""CPU tests of the passes either side of the hot path in the oracle: interleaved rendering
(InterleaveBuilder.h / interleave.frag / deinterleave.frag), the Depth filter of the mip builder (mipLevelBuilder.frag:29-42) and one
quad of the debug overlay (DebugRenderer.h). The plain-C restatement is pinned bit for bit against the reference arm (the
reference's own SPIR-V for these passes) and against the integer contracts of the index maps."""
import ctypes as C

import numpy as np
import pytest

from legitengine_b200 import abi, images
from oracle import loader
from tests import aux_helpers as A

needs_ref = pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")

SIZES = [(64, 48, 4, 4), (250, 141, 4, 4), (131, 77, 3, 2), (17, 9, 4, 4), (96, 64, 8, 8), (33, 21, 1, 1), (40, 30, 40, 30)]
FORMATS = [abi.FORMAT_R16G16B16A16_SFLOAT, abi.FORMAT_R32G32_SFLOAT, abi.FORMAT_R32G32B32A32_SFLOAT]


@pytest.mark.parametrize("case", SIZES)
@pytest.mark.parametrize("fmt", FORMATS)
def test_interleave_maps_follow_the_shader_formulas(case, fmt):
    """Bit-exact index contract: the oracle's passes move exactly the texels the GLSL formulas name (numpy restatement)."""
    W, Hh, gx, gy = case
    src = A.random_image(fmt, W, Hh, seed=W * 7 + gx)
    raw = src.level_bytes(0)
    p = A.interleave_params(W, Hh, gx, gy)
    x, y = np.meshgrid(np.arange(W), np.arange(Hh))
    dvx, dvy = W // gx, Hh // gy
    d = A.run_pass(loader.port().deinterleave, p, src)
    assert np.array_equal(d.level_bytes(0), raw[(y % dvy) * gy + y // dvy, (x % dvx) * gx + x // dvx])  # deinterleave.frag:17-22
    i = A.run_pass(loader.port().interleave, p, src)
    assert np.array_equal(i.level_bytes(0), raw[(y % gy) * dvy + y // gy, (x % gx) * dvx + x // gx])  # interleave.frag:16-21


@pytest.mark.parametrize("case", [(64, 48, 4, 4), (256, 144, 4, 4), (96, 60, 3, 5)])
def test_interleave_inverts_deinterleave_on_divisible_viewports(case):
    W, Hh, gx, gy = case
    src = A.random_image(abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, seed=3)
    p = A.interleave_params(W, Hh, gx, gy)
    port = loader.port()
    back = A.run_pass(port.interleave, p, A.run_pass(port.deinterleave, p, src))
    assert back.levels_equal(src, 0)


@needs_ref
@pytest.mark.parametrize("case", SIZES)
@pytest.mark.parametrize("fmt", FORMATS)
def test_interleave_port_equals_reference_arm(case, fmt):
    W, Hh, gx, gy = case
    src = A.random_image(fmt, W, Hh, seed=W + gy)
    p = A.interleave_params(W, Hh, gx, gy)
    for name in ("deinterleave", "interleave"):
        a = A.run_pass(getattr(loader.port(), name), p, src)
        b = A.run_pass(getattr(loader.ref(), name), p, src)
        assert a.levels_equal(b, 0), name
    rows = abi.LgcuRows(5, Hh - 3)  # row strips only touch their rows
    part = A.run_pass(loader.port().deinterleave, p, src, rows=rows)
    whole = A.run_pass(loader.port().deinterleave, p, src)
    assert np.array_equal(part.level_bytes(0)[5:Hh - 3], whole.level_bytes(0)[5:Hh - 3])
    assert np.all(part.level_bytes(0)[:5] == 0xCD) and np.all(part.level_bytes(0)[Hh - 3:] == 0xCD)


def test_interleave_rejects_bad_arguments():
    src = A.random_image(abi.FORMAT_R16G16B16A16_SFLOAT, 32, 16, seed=1)
    dst = images.HostImage(abi.FORMAT_R16G16B16A16_SFLOAT, 32, 16)
    for gx, gy, vw, vh in ((0, 4, 32, 16), (4, 4, 31, 16), (64, 4, 32, 16)):
        p = abi.InterleaveData((C.c_int32 * 4)(gx, gy, 0, 0), (C.c_int32 * 4)(vw, vh, 0, 0))
        assert loader.port().deinterleave(C.byref(p), C.byref(src.view()), C.byref(dst.view()), None) == abi.LGCU_ERR_INVALID_ARGUMENT


@needs_ref
@pytest.mark.parametrize("fmt", [abi.FORMAT_R16G16B16A16_SFLOAT, abi.FORMAT_R32G32_SFLOAT])
@pytest.mark.parametrize("size", [(64, 48), (250, 141), (33, 17)])
def test_depth_filter_mip_port_equals_reference_arm(size, fmt):
    """mipLevelBuilder.frag:29-42 (FilterTypes::Depth): min / max / mass per 2x2 block, through the reference's own SPIR-V."""
    W, Hh = size
    src = A.random_depth_range_image(fmt, W, Hh, seed=W)
    p = abi.MipLevelBuilderData(1.0)
    outs = []
    for be in (loader.port(), loader.ref()):
        dst = images.HostImage(fmt, W, Hh, 2)
        dst.level_bytes(0)[...] = src.level_bytes(0)
        assert be.mip_level(C.byref(p), C.byref(dst.view(0, 1)), C.byref(dst.view(1, 1)), None) == 0
        outs.append(dst)
    assert A.equal_nan_aware(outs[0], outs[1], 1)
    lvl = outs[0].level_f32(1)
    s = src.level_f32(0)[: (Hh // 2) * 2, : (W // 2) * 2]
    blocks = s.reshape(Hh // 2, 2, W // 2, 2, -1)
    assert np.array_equal(lvl[..., 0], blocks[..., 0].min(axis=(1, 3)))  # min of .x
    assert np.array_equal(lvl[..., 1], blocks[..., 1].max(axis=(1, 3)))  # max of .y


@needs_ref
@pytest.mark.parametrize("target_fmt", [abi.FORMAT_B8G8R8A8_SRGB, abi.FORMAT_R16G16B16A16_SFLOAT])
@pytest.mark.parametrize("size", [(320, 180), (250, 141), (64, 48)])
def test_debug_overlay_port_equals_reference_arm(size, target_fmt):
    W, Hh = size
    for k, quad in enumerate(A.debug_tiles(4)):
        src = A.random_image(abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, seed=k, lo=0.0, hi=1.5)
        outs = []
        for be in (loader.port(), loader.ref()):
            target = A.random_image(target_fmt, W, Hh, seed=99)
            assert be.debug_overlay(C.byref(quad), C.byref(src.view()), C.byref(target.view()), None) == 0
            outs.append(target)
        assert outs[0].levels_equal(outs[1], 0)
        # the quad covers about tile-size pixels and nothing else changes (loadOp eLoad)
        before = A.random_image(target_fmt, W, Hh, seed=99).level_bytes(0)
        changed = (outs[0].level_bytes(0) != before).any(axis=2)
        ys, xs = np.nonzero(changed)
        mm = quad.minmax
        assert xs.min() >= int(mm[0] * W) - 1 and xs.max() <= int(mm[2] * W) + 1 and ys.min() >= int(mm[1] * Hh) - 1 and ys.max() <= int(mm[3] * Hh) + 1
        assert changed.sum() >= 0.8 * (mm[2] - mm[0]) * W * (mm[3] - mm[1]) * Hh - 4


def test_debug_tile_layout_matches_the_reference_loop():
    """DebugRenderer.h:27-57: tiles of 0.1 with 0.02 padding, left to right, wrapping when the next tile would cross 1.0."""
    tiles = [tuple(round(v, 5) for v in t.minmax) for t in A.debug_tiles(10)]
    assert tiles[0] == (0.02, 0.02, 0.12, 0.12) and tiles[1] == (0.14, 0.02, 0.24, 0.12) and tiles[3] == (0.38, 0.02, 0.48, 0.12)
    assert tiles[8][0] == 0.02 and tiles[8][1] == 0.14  # 8 tiles fit in a row (0.02 + 8 * 0.12 = 0.98, the 9th would end at 1.08)


def test_port_reproduces_reference_golden_fixture_of_the_aux_passes():
    """tests/golden/aux_passes.npz holds what the reference's own SPIR-V produced (tests/golden/make_golden.py); the port must
    reproduce it bit for bit — also where /root/reference and oracle/_ref are absent."""
    z = A.load_aux_golden()
    port = loader.port()
    for kind, k, fmt, dims in A.golden_cases(z):
        if kind == "interleave":
            W, Hh, gx, gy = dims
            src = A.image_from_bytes(fmt, z[f"interleave{k}.src"])
            p = A.interleave_params(W, Hh, gx, gy)
            assert np.array_equal(A.run_pass(port.deinterleave, p, src).level_bytes(0), z[f"interleave{k}.deinterleaved"]), (kind, k)
            assert np.array_equal(A.run_pass(port.interleave, p, src).level_bytes(0), z[f"interleave{k}.interleaved"]), (kind, k)
        elif kind == "depthmip":
            img = A.image_from_bytes(fmt, z[f"depthmip{k}.src"], mips=2)
            mp = abi.MipLevelBuilderData(1.0)
            assert port.mip_level(C.byref(mp), C.byref(img.view(0, 1)), C.byref(img.view(1, 1)), None) == 0
            want = images.HostImage(fmt, dims[0], dims[1], 2)
            want.level_bytes(1)[...] = z[f"depthmip{k}.level1"]
            assert A.equal_nan_aware(img, want, 1), (kind, k)
        else:
            W, Hh, tiles = dims
            target = A.image_from_bytes(fmt, z["overlay.target_before"])
            for t in range(tiles):
                quad = abi.DebugQuadData((C.c_float * 4)(*[float(v) for v in z[f"overlay.quad{t}"]]))
                src = A.image_from_bytes(abi.FORMAT_R16G16B16A16_SFLOAT, z[f"overlay.src{t}"])
                assert port.debug_overlay(C.byref(quad), C.byref(src.view()), C.byref(target.view()), None) == 0
            assert np.array_equal(target.level_bytes(0), z["overlay.target_after"])
