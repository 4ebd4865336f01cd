"""GPU parity on the reference's bundled sample scene (BASELINE configs[0]) at fixture size: the CUDA passes, from the committed
fixture's inputs (tests/golden/bundled_sponza_192x108.npz: fragment buffer of SponzaScene.json rasterised by the oracle, object
table, shadow map), against the images the REFERENCE's own SPIR-V passes produced from them. Real geometry (466k triangles, thin
features, depth discontinuities everywhere) instead of the synthetic boxes. (Named zz so that it runs after the other GPU tests.)"""
import pytest

from legitengine_b200 import passes
from tests import helpers as H
from tests.test_cuda_parity import _denoise_final, _gather, _sync

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
    _denoise_final(cu, p, dev)
    _sync()
    # stage by stage against the oracle's passes on the device's own inputs (one bar), G-buffer and depth-moment chains bit-exact
    # against the fixture = the reference's own SPIR-V output
    H.stagewise_check(lambda n: getattr(dev, n).to_host(), p, ref, exact=(mode != "strict"), what=f"bundled Sponza {mode}")
    # and the gather on the fixture's own pyramids straight against the reference's indirectLight
    dev2 = H.device_frame_like(ref, copy=("blurredDirectLight", "blurredDepthMoments", "normal", "depthStencil"))
    _gather(cu, p, dev2, mode)
    _sync()
    H.assert_images_radiance(dev2.indirectLight.to_host(), ref.indirectLight, f"bundled Sponza {mode}: gather on the fixture's pyramids",
                             exact=None if mode == "strict" else H.exact_gather(p, ref, (0, ref.height)))
