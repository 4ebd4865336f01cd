"""GPU parity on the reference's bundled sample scene (BASELINE configs[0]) at fixture size: the CUDA passes, from the committed
fixture's inputs (tests/golden/bundled_sponza_192x108.npz: fragment buffer of SponzaScene.json rasterised by the oracle, object
table, shadow map), against the images the REFERENCE's own SPIR-V passes produced from them. Real geometry (466k triangles, thin
features, depth discontinuities everywhere) instead of the synthetic boxes. (Named zz so that it runs after the other GPU tests.)"""
import ctypes as C

import numpy as np
import pytest

from legitengine_b200 import passes
from tests import helpers as H
from tests.test_cuda_parity import _gather, _sync, _v

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    return passes.CudaPasses()


@pytest.mark.parametrize("mode", ["strict", "packed"])
def test_bundled_scene_frame_vs_reference_fixture(cu, mode):
    sc, p, ref, _ = H.load_golden(H.BUNDLED_GOLDEN)
    dev = H.device_frame_like(ref)
    inp = passes.upload_inputs(dev, sc)
    passes.run_pass_list(cu, dev, p, inp, stop_after="blur")
    _gather(cu, p, dev, mode)
    _sync()
    levels = passes.mip_levels_built(sc.width, sc.height)
    for iname in ("normal", "depthMoments", "depthStencil"):
        H.assert_bit_exact(getattr(dev, iname).to_host(), getattr(ref, iname), 0, iname)
    moments, blurred = dev.depthMoments.to_host(), dev.blurredDepthMoments.to_host()
    for l in range(levels):  # exact-order fp32 chains: bit-exact on every level
        H.assert_bit_exact(moments, ref.depthMoments, l, "depthMoments")
        H.assert_bit_exact(blurred, ref.blurredDepthMoments, l, "blurredDepthMoments")
    for iname in ("directLight", "blurredDirectLight"):
        host = getattr(dev, iname).to_host()
        for l in range(levels):
            w, h = sc.width >> l, sc.height >> l
            if w * h >= 1024:
                H.assert_close(host, getattr(ref, iname), l, iname, max_outside_frac=5e-3)
    H.assert_close(dev.indirectLight.to_host(), ref.indirectLight, 0, "indirectLight", max_outside_frac=5e-3 if mode == "strict" else 1e-2)
    cu.denoise(C.byref(p.denoiser), _v(dev.indirectLight), _v(dev.normal), _v(dev.depthMoments), _v(dev.denoisedIndirectLight), None)
    cu.final_gather(C.byref(p.final), _v(dev.directLight), _v(dev.blurredDirectLight), _v(dev.albedo), _v(dev.denoisedIndirectLight), _v(dev.swapchain), None)
    _sync()
    a = dev.swapchain.to_host().level_raw(0).astype(np.int32)
    b = ref.swapchain.level_raw(0).astype(np.int32)
    assert (np.abs(a - b) > 1).mean() < 1e-2

