"""GPU parity of the passes either side of the hot path (csrc/k_aux.cu and the Depth branch of the mip
kernel) through the C ABI against the CPU oracle: (de)interleave bit-exact incl. ragged viewports, row strips and BASELINE's 4K
size (round trip + oracle-free index check), Depth-filter mips bit-exact, debug overlay bit-exact on float targets and within one
sRGB8 step on the swapchain format."""
import ctypes as C

import numpy as np
import pytest

from legitengine_b200 import abi, images, passes
from oracle import loader
from tests import aux_helpers as A

pytestmark = pytest.mark.gpu

SIZES = [(64, 48, 4, 4), (250, 141, 4, 4), (131, 77, 3, 2), (17, 9, 4, 4), (96, 64, 8, 8), (33, 21, 1, 1), (40, 30, 40, 30), (1920, 1080, 4, 4)]
FORMATS = [abi.FORMAT_R16G16B16A16_SFLOAT, abi.FORMAT_R32G32_SFLOAT, abi.FORMAT_R32G32B32A32_SFLOAT]


@pytest.fixture(scope="module")
def cu():
    return passes.CudaPasses()


def _cuda_pass(fn, params, src_host, rows=None, extra_src=None):
    import torch

    src = images.DeviceImage.from_host(src_host)
    dst = images.DeviceImage(src_host.format, src_host.desc.width, src_host.desc.height, 1)
    fn(C.byref(params), C.byref(src.view()), C.byref(dst.view()), C.byref(rows) if rows is not None else None)
    torch.cuda.synchronize()
    return dst.to_host()


@pytest.mark.parametrize("case", SIZES)
@pytest.mark.parametrize("fmt", FORMATS)
def test_interleave_passes_bit_exact(cu, case, fmt):
    W, Hh, gx, gy = case
    src = A.random_image(fmt, W, Hh, seed=W * 3 + gy)
    p = A.interleave_params(W, Hh, gx, gy)
    for name in ("deinterleave", "interleave"):
        want = A.run_pass(getattr(loader.port(), name), p, src)
        got = _cuda_pass(getattr(cu, name), p, src)
        assert got.levels_equal(want, 0), (name, case, fmt)


@pytest.mark.parametrize("name", ["deinterleave", "interleave"])
def test_interleave_row_strips(cu, name):
    W, Hh = 250, 141
    src = A.random_image(abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, seed=5)
    p = A.interleave_params(W, Hh, 4, 4)
    want = A.run_pass(getattr(loader.port(), name), p, src).level_bytes(0)
    got = _cuda_pass(getattr(cu, name), p, src, rows=abi.LgcuRows(37, 101)).level_bytes(0)
    assert np.array_equal(got[37:101], want[37:101])
    assert np.all(got[:37] == 0xCD) and np.all(got[101:] == 0xCD)


def test_interleave_4k_round_trip_and_index_map(cu):
    """BASELINE's 3840x2160 through size-independent properties: interleave(deinterleave(x)) == x, and every de-interleaved texel
    is the one the shader formula names (numpy index map, no oracle run)."""
    import torch

    W, Hh, g = 3840, 2160, 4
    src = A.random_image(abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, seed=11)
    p = A.interleave_params(W, Hh, g, g)
    a = images.DeviceImage.from_host(src)
    b = images.DeviceImage(src.format, W, Hh, 1)
    c = images.DeviceImage(src.format, W, Hh, 1)
    cu.deinterleave(C.byref(p), C.byref(a.view()), C.byref(b.view()), None)
    cu.interleave(C.byref(p), C.byref(b.view()), C.byref(c.view()), None)
    torch.cuda.synchronize()
    assert c.to_host().levels_equal(src, 0)
    x, y = np.meshgrid(np.arange(W), np.arange(Hh))
    raw = src.level_bytes(0)
    assert np.array_equal(b.to_host().level_bytes(0), raw[(y % (Hh // g)) * g + y // (Hh // g), (x % (W // g)) * g + x // (W // g)])


def test_interleave_rejects_bad_arguments():
    lib = abi.load_lgcu()
    src = images.DeviceImage(abi.FORMAT_R16G16B16A16_SFLOAT, 32, 16, 1)
    dst = images.DeviceImage(abi.FORMAT_R16G16B16A16_SFLOAT, 32, 16, 1)
    other = images.DeviceImage(abi.FORMAT_R32G32_SFLOAT, 32, 16, 1)
    for gx, gy, vw, vh in ((0, 4, 32, 16), (4, 4, 31, 16), (64, 4, 32, 16)):
        p = abi.InterleaveData((C.c_int32 * 4)(gx, gy, 0, 0), (C.c_int32 * 4)(vw, vh, 0, 0))
        assert lib.lgcu_deinterleave(C.byref(p), C.byref(src.view()), C.byref(dst.view()), None, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    p = A.interleave_params(32, 16, 4, 4)
    assert lib.lgcu_interleave(C.byref(p), C.byref(src.view()), C.byref(other.view()), None, None) == abi.LGCU_ERR_UNSUPPORTED_FORMAT


@pytest.mark.parametrize("fmt", [abi.FORMAT_R16G16B16A16_SFLOAT, abi.FORMAT_R32G32_SFLOAT])
@pytest.mark.parametrize("size", [(64, 48), (250, 141), (33, 17), (1920, 1080)])
def test_depth_filter_mip_bit_exact(cu, size, fmt):
    import torch

    W, Hh = size
    src = A.random_depth_range_image(fmt, W, Hh, seed=W + 1)
    p = abi.MipLevelBuilderData(1.0)
    want = images.HostImage(fmt, W, Hh, 2)
    want.level_bytes(0)[...] = src.level_bytes(0)
    assert loader.port().mip_level(C.byref(p), C.byref(want.view(0, 1)), C.byref(want.view(1, 1)), None) == 0
    dev = images.DeviceImage.from_host(want)
    dev.tensor[dev.desc.levelOffset[1]:] = 0xCD
    cu.mip_level(C.byref(p), C.byref(dev.view(0, 1)), C.byref(dev.view(1, 1)), None)
    torch.cuda.synchronize()
    assert A.equal_nan_aware(dev.to_host(), want, 1)


@pytest.mark.parametrize("target_fmt", [abi.FORMAT_B8G8R8A8_SRGB, abi.FORMAT_R16G16B16A16_SFLOAT, abi.FORMAT_R32G32B32A32_SFLOAT])
@pytest.mark.parametrize("size", [(320, 180), (250, 141), (1920, 1080)])
def test_debug_overlay(cu, size, target_fmt):
    import torch

    W, Hh = size
    target_host = A.random_image(target_fmt, W, Hh, seed=99)
    want = A.random_image(target_fmt, W, Hh, seed=99)
    got_dev = images.DeviceImage.from_host(target_host)
    for k, quad in enumerate(A.debug_tiles(4)):  # the four thumbnails of SSVGIRenderer.h:344-350, drawn one after the other
        src = A.random_image(abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, seed=k, lo=0.0, hi=1.5)
        assert loader.port().debug_overlay(C.byref(quad), C.byref(src.view()), C.byref(want.view()), None) == 0
        src_dev = images.DeviceImage.from_host(src)
        cu.debug_overlay(C.byref(quad), C.byref(src_dev.view()), C.byref(got_dev.view()), None)
        torch.cuda.synchronize()
    got = got_dev.to_host()
    if target_fmt == abi.FORMAT_B8G8R8A8_SRGB:
        d = np.abs(got.level_bytes(0).astype(np.int32) - want.level_bytes(0).astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3  # the sRGB encode's powf may differ in the last bit between libm and CUDA
    else:
        assert got.levels_equal(want, 0)
    below = int(0.2 * Hh)  # the four tiles end at y = 0.12: everything below keeps the target's contents (loadOp eLoad)
    assert np.array_equal(got.level_bytes(0)[below:], target_host.level_bytes(0)[below:])


def test_cuda_reproduces_reference_golden_fixture_of_the_aux_passes(cu):
    """The CUDA passes against tests/golden/aux_passes.npz (outputs of the reference's own SPIR-V): bit-exact; the sRGB8 overlay
    target within one code."""
    import torch

    z = A.load_aux_golden()
    for kind, k, fmt, dims in A.golden_cases(z):
        if kind == "interleave":
            W, Hh, gx, gy = dims
            src = A.image_from_bytes(fmt, z[f"interleave{k}.src"])
            p = A.interleave_params(W, Hh, gx, gy)
            assert np.array_equal(_cuda_pass(cu.deinterleave, p, src).level_bytes(0), z[f"interleave{k}.deinterleaved"]), (kind, k)
            assert np.array_equal(_cuda_pass(cu.interleave, p, src).level_bytes(0), z[f"interleave{k}.interleaved"]), (kind, k)
        elif kind == "depthmip":
            dev = images.DeviceImage.from_host(A.image_from_bytes(fmt, z[f"depthmip{k}.src"], mips=2))
            mp = abi.MipLevelBuilderData(1.0)
            cu.mip_level(C.byref(mp), C.byref(dev.view(0, 1)), C.byref(dev.view(1, 1)), None)
            torch.cuda.synchronize()
            want = images.HostImage(fmt, dims[0], dims[1], 2)
            want.level_bytes(1)[...] = z[f"depthmip{k}.level1"]
            assert A.equal_nan_aware(dev.to_host(), want, 1), (kind, k)
        else:
            W, Hh, tiles = dims
            target = images.DeviceImage.from_host(A.image_from_bytes(fmt, z["overlay.target_before"]))
            for t in range(tiles):
                quad = abi.DebugQuadData((C.c_float * 4)(*[float(v) for v in z[f"overlay.quad{t}"]]))
                src = images.DeviceImage.from_host(A.image_from_bytes(abi.FORMAT_R16G16B16A16_SFLOAT, z[f"overlay.src{t}"]))
                cu.debug_overlay(C.byref(quad), C.byref(src.view()), C.byref(target.view()), None)
                torch.cuda.synchronize()
            d = np.abs(target.to_host().level_bytes(0).astype(np.int32) - z["overlay.target_after"].astype(np.int32))
            assert d.max() <= 1 and (d > 0).mean() < 1e-3
