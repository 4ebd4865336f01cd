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


def test_bundled_scene_512_from_the_mesh():
    """BASELINE configs[0] at its own size: the bundled scene (SponzaScene.json: 364 157 triangles in 5 draws, packed by
    oracle/make_bundled_mesh.py where /root/reference exists) rendered 512x512 from its vertex / index buffers — ShadowPass and
    GBufferRasterPass on the device (the large-scene path of k_raster.cu), then the fused frame — against the oracle's rasteriser
    feeding the oracle's passes: shadow map and G-buffer bit-exact, every later stage at the one bar on its own inputs."""
    from legitengine_b200 import abi, harness
    from oracle import frames as OF
    from oracle import make_bundled_mesh as MB

    mesh = MB.load()
    if mesh is None:
        pytest.skip("oracle/_ref/bundled_sponza_mesh.npz not generated (needs /root/reference at build time)")
    W = Hh = 512
    sc, p, ref = OF.oracle_frame_from_mesh(mesh, W, Hh)
    r = harness.Renderer(W, Hh)
    r.upload_mesh(mesh)
    r.render_frame(harness.MODE_FUSED, 0, abi.GI_DEFAULT, profile=True)
    r.sync()
    names = [n for n, _ in r.profile()]
    assert names[0] == "ShadowPass" and names[1] == "GBufferRasterPass", names
    for n in ("shadowMap", "albedo", "emissive", "normal", "depthStencil"):
        H.assert_bit_exact(r.download_image(n), getattr(ref, n), 0, n)
    H.stagewise_check(r.download_image, p, ref, what="bundled Sponza 512x512 from the mesh")
    r.close()
