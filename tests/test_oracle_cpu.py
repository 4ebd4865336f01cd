"""CPU tests of the oracle (no GPU): the plain-C restatement (oracle/ssvgi_oracle.c) is pinned

  * bit for bit against the committed golden fixtures, which are outputs of the REFERENCE's own SPIR-V passes
    (tests/golden/make_golden.py), and
  * bit for bit against that reference arm itself (oracle/_ref/libref_spirv.so) on fresh seeded frames, when it is built
    (it is built wherever /root/reference exists and travels to the GPU box as a prebuilt .so),

and against the known-answer values SURVEY.md Appendix D derives from the reference's formulas.
"""
import ctypes as C

import numpy as np
import pytest

from legitengine_b200 import abi, images, passes, scene
from oracle import loader
from tests import helpers as H


def _all_levels(fi):
    for name, img in fi.items():
        for l in range(img.mips):
            w, h = img.level_size(l)
            if w > 0 and h > 0:
                yield name, img, l


def _run(be, sc, p, W, Hh, shadow):
    fi = passes.FrameImages(W, Hh, images.HostImage, shadow_size=shadow)
    passes.run_pass_list(be, fi, p, passes.upload_inputs(fi, sc))
    return fi


@pytest.mark.parametrize("name", H.GOLDEN_NAMES)
def test_port_reproduces_reference_golden_fixture(name):
    sc, p, want, _ = H.load_golden(name)
    got = _run(loader.port(), sc, p, sc.width, sc.height, want.shadow_size)
    levels = passes.mip_levels_built(sc.width, sc.height)
    for iname, img, l in _all_levels(want):
        if iname == "shadowMap":
            continue
        if l >= levels and iname in ("directLight", "depthMoments"):
            continue  # levels MipBuilder never builds (zero-sized): poison on both sides, nothing to compare
        assert getattr(got, iname).levels_equal(img, l), f"{name}: {iname} level {l} differs from the reference fixture"


def test_golden_fixtures_exist():
    assert len(H.GOLDEN_NAMES) >= 3


@pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", [(11, 64, 48, 0), (12, 250, 141, 0), (13, 131, 77, 2), (14, 320, 180, 0), (15, 17, 9, 2)])
def test_port_equals_reference_arm(case):
    seed, W, Hh, radius = case
    sc = scene.make_scene(seed, W, Hh, n_boxes=32, shadow_size=256)
    p = passes.make_params(W, Hh, sc.matrices, radius)
    a = _run(loader.port(), sc, p, W, Hh, 256)
    b = _run(loader.ref(), sc, p, W, Hh, 256)
    for iname, img, l in _all_levels(a):
        assert img.levels_equal(getattr(b, iname), l), f"{iname} level {l}: port != reference SPIR-V arm"


@pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built")
def test_port_row_strips_equal_whole_frame():
    """lgcu_rows contract on the oracle side: computing the frame strip by strip gives the same images."""
    W, Hh = 96, 64
    sc = scene.make_scene(21, W, Hh, n_boxes=16, shadow_size=128)
    p = passes.make_params(W, Hh, sc.matrices, 0)
    whole = _run(loader.port(), sc, p, W, Hh, 128)
    fi = passes.FrameImages(W, Hh, images.HostImage, shadow_size=128)
    inp = passes.upload_inputs(fi, sc)
    be = loader.port()
    # stage by stage, every stage over two strips (a stage needs the previous stage complete, like the multi-GPU path)
    for stop in ("gbuffer", "light", "mips", "blur", "gather", None):
        for rows in ((0, 32), (32, 64)):
            passes.run_pass_list(be, fi, p, inp, rows=rows, stop_after=stop)
    for iname, img, l in _all_levels(whole):
        assert img.levels_equal(getattr(fi, iname), l), f"{iname} level {l}"


# ---------------------------------------------------------------------------------------------- known answers (Appendix D)
def test_frame_matrices_known_answers():
    m = scene.frame_matrices(512, 512)
    view = m.view.reshape(4, 4)  # rows of this array are glm columns
    np.testing.assert_allclose(view[:3, :3], np.eye(3), atol=1e-7)
    np.testing.assert_allclose(view[3], [0.0, -0.5, 2.0, 1.0], atol=1e-7)
    proj = m.proj.reshape(4, 4)
    np.testing.assert_allclose(proj[0], [1.83048773, 0, 0, 0], rtol=1e-7)
    np.testing.assert_allclose(proj[1], [0, -1.83048773, 0, 0], rtol=1e-7)
    np.testing.assert_allclose(proj[2], [0, 0, 1.00002003, 1], rtol=1e-7)
    np.testing.assert_allclose(proj[3], [0, 0, -0.0200002007, 0], rtol=1e-6)
    lv = m.light_view.reshape(4, 4)
    np.testing.assert_allclose(lv[1], [0, 4.63287033e-05, -1, 0], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(lv[3], [0, -0.000231643513, 5, 1], rtol=1e-4, atol=1e-9)
    lp = m.light_proj.reshape(4, 4)
    np.testing.assert_allclose(lp[0][0], 2.36522222, rtol=1e-7)
    np.testing.assert_allclose(lp[2][2], 1.002002, rtol=1e-6)
    np.testing.assert_allclose(lp[3][2], -0.2002002, rtol=1e-6)
    # world (0, 0.5, 1) -> ndc z 0.993353307 through proj * view
    pv = proj.T @ view.T
    clip = pv @ np.array([0.0, 0.5, 1.0, 1.0])
    assert abs(clip[2] / clip[3] - 0.993353307) < 1e-6


@pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built")
def test_frame_matrices_equal_reference_glm():
    """host frame maths (synth_scene.cpp / Camera.h) == the same expressions evaluated by the reference's vendored glm."""
    ref = loader.ref().lib
    f4 = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    for (W, Hh) in [(512, 512), (1920, 1080), (3840, 2160), (7680, 4320)]:
        m = scene.frame_matrices(W, Hh)
        cp = np.asarray(scene.DEFAULT_CAMERA["pos"], dtype=np.float32)
        lp = np.asarray(scene.DEFAULT_LIGHT["pos"], dtype=np.float32)
        out = [np.zeros(16, dtype=np.float32) for _ in range(4)]
        ref.ref_frame_matrices(f4(cp), 0.0, 0.0, f4(lp), scene.DEFAULT_LIGHT["vert"], 0.0, W, Hh, *[f4(o) for o in out])
        for got, want in zip((m.view, m.proj, m.light_view, m.light_proj), out):
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _gather_single_pixel_setup(W, Hh):
    """Uniform light (1, .5, .25), camera-facing normal, flat depth: Appendix D smoke value."""
    m = scene.frame_matrices(W, Hh)
    p = passes.make_params(W, Hh, m, 0)
    fi = passes.FrameImages(W, Hh, images.HostImage, indirect_format=abi.FORMAT_R32G32B32A32_SFLOAT, shadow_size=16)
    levels = passes.MIPS
    for l in range(levels):
        w, h = fi.blurredDirectLight.level_size(l)
        if w <= 0 or h <= 0:
            continue
        fi.blurredDirectLight.set_level(l, np.broadcast_to(np.array([1.0, 0.5, 0.25, 1.0], np.float32), (h, w, 4)))
        fi.blurredDepthMoments.set_level(l, np.broadcast_to(np.array([3.0, 9.0], np.float32), (h, w, 2)))
    fi.normal.set_level(0, np.broadcast_to(np.array([0.0, 0.0, -1.0, 1.0], np.float32), (Hh, W, 4)))
    # ndc depth of a point 3 units along the view axis (z_view = 3): clip = proj * (0,0,3,1)
    proj = m.proj.reshape(4, 4)
    z = (proj[2][2] * 3.0 + proj[3][2]) / (proj[2][3] * 3.0)
    fi.depthStencil.set_level(0, np.full((Hh, W, 1), z, np.float32))
    return p, fi


def test_gather_uniform_light_known_answer():
    W = Hh = 64
    p, fi = _gather_single_pixel_setup(W, Hh)
    v = lambda img, b=0, n=None: C.byref(img.view(b, n))
    loader.port().gi_gather(C.byref(p.indirect), v(fi.blurredDirectLight), v(fi.blurredDepthMoments), v(fi.normal), v(fi.depthStencil),
                            v(fi.indirectLight), 0, None)
    out = fi.indirectLight.level_f32(0)
    # the gather is affine in the light: out_c = A + B * L_c with A, B independent of the channel (indirectLighting.frag:209, 261-262)
    assert np.all(np.isfinite(out))
    b1 = (out[..., 0] - out[..., 1]) / (1.0 - 0.5)
    b2 = (out[..., 1] - out[..., 2]) / (0.5 - 0.25)
    assert np.all(np.abs(b1 - b2) <= 1e-4 * np.abs(b1) + 1e-6) and np.all(b1 > 0.0)
    # where the constant-distance shell covers the whole hemisphere the indirect light is the incident light (Appendix D smoke value)
    np.testing.assert_allclose(out[10, 10, :3], [1.0, 0.5, 0.25], rtol=0.02)
    assert np.all(out[..., 3] == 1.0)


def test_hammersley_table_and_march_schedule():
    """Appendix D: idx -> (ang, lin) table; march offsets / LODs for lin = 0 at W = 1920, 3840, 7680, 512."""
    def bitrev(i):
        b = ((i << 16) | (i >> 16)) & 0xFFFFFFFF
        b = ((b & 0x55555555) << 1) | ((b & 0xAAAAAAAA) >> 1)
        b = ((b & 0x33333333) << 2) | ((b & 0xCCCCCCCC) >> 2)
        b = ((b & 0x0F0F0F0F) << 4) | ((b & 0xF0F0F0F0) >> 4)
        b = ((b & 0x00FF00FF) << 8) | ((b & 0xFF00FF00) >> 8)
        return b
    want = {1: (.0625, .5), 2: (.125, .25), 3: (.1875, .75), 7: (.4375, .875), 8: (.5, .0625), 13: (.8125, .6875), 15: (.9375, .9375)}
    for idx, (ang, lin) in want.items():
        assert idx / 16.0 == ang and np.float32(bitrev(idx)) / np.float32(4294967296.0) == np.float32(lin)
    sched = {1920: [(1.00, None), (4.02, -0.76), (11.77, 1.08), (31.70, 2.59), (82.94, 4.01), (214.66, 5.39), (553.27, 6.76), (1423.77, 8.13)],
             3840: [(1.00, None), (7.03, 0.24), (22.54, 2.08), (62.40, 3.59), (164.87, 5.01), (428.31, 6.39), (1105.54, 7.76), (2846.54, 9.13)],
             512: [(1.00, None), (1.80, -2.66), (3.87, -0.83), (9.19, 0.68), (22.85, 2.10), (57.97, 3.48), (148.27, 4.85), (380.41, 6.22)]}
    f = np.float32
    for W, steps in sched.items():
        near = f(W) / f(1000.0)
        for k, (off_want, lod_want) in enumerate(steps):
            off = f(f(near * f(np.power(f(2.57075), f(k)))) + f(1.0)) - near
            assert abs(float(off) - off_want) < 0.006 * max(1.0, off_want), (W, k, off)
            if lod_want is not None:
                lod = np.log(max(0.0, 1.57075 * (float(off) - 1.0) * 0.5)) / 0.693147182 - 2.0
                assert abs(lod - lod_want) < 0.01, (W, k, lod)


def test_stagewise_checker_on_the_oracle_itself():
    """The chained-frame parity helper of the -m gpu tests (tests/helpers.py: stagewise_check = every stage against the oracle's pass
    on that stage's own inputs, whole frame or row strips, with the binary64 gather alongside), exercised here with the oracle's own
    frame standing in for the device's: everything must agree exactly, and a corrupted stage must be caught."""
    W, Hh = 320, 180
    sc, p, whole = H.oracle_frame(11, W, Hh)
    H.stagewise_check(lambda name: getattr(whole, name), p, whole)
    strips = ((0, 16), (80, 96), (164, 180))
    H.stagewise_check(lambda name: getattr(whole, name), p, whole, strips=strips, exact=True)
    import copy
    broken = copy.copy(whole)
    broken.indirectLight = images.HostImage(whole.indirectLight.format, W, Hh, 1)
    broken.indirectLight.buf[...] = whole.indirectLight.buf
    lv = broken.indirectLight.level_raw(0).copy()
    lv[80:96, ::7, 0] += np.float16(0.01)
    broken.indirectLight.set_level(0, lv)
    with pytest.raises(AssertionError):
        H.stagewise_check(lambda name: getattr(broken, name), p, whole, strips=strips)


def test_reference_rounding_noise_floor_of_the_gather():
    """oracle/gather_noise_probe.c: the gather in binary64 on the fp32 oracle's inputs, with the shader's discrete decisions. At a
    small viewport the fp32 shader sits within the bar of the exact value of its own formula (so the plain bar is meaningful there);
    the probe and the port agree, which also pins the probe's restatement of the formula."""
    W, Hh = 250, 141
    sc, p, ref = H.oracle_frame(12, W, Hh)
    exact, cut = H.exact_gather(p, ref, (0, Hh))
    keep = ~cut
    rep = H.radiance_report(ref.indirectLight.level_f32(0)[..., :3][keep], exact[keep], "fp32 oracle vs binary64 probe 250x141")
    assert rep["outside"] <= H.OUTLIER_BAR and rep["psnr"] >= 60.0 and cut.mean() <= 2e-3, (rep, cut.mean())


def test_gather_work_counters():
    """orc_gi_gather's work counters (flag bit 31): march samples per pixel follow from the iteration-count formula alone
    (indirectLighting.frag:202-214), so they can be cross-checked in numpy; hits are bounded by the samples."""
    W, Hh = 250, 141
    sc, p, ref = H.oracle_frame(12, W, Hh)
    port = loader.port()
    port.lib.orc_gi_gather_counters.argtypes = [C.POINTER(C.c_ulonglong)]
    out_img = images.HostImage(abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, 1)
    v = lambda img: C.byref(img.view())
    assert port.gi_gather(C.byref(p.indirect), v(ref.blurredDirectLight), v(ref.blurredDepthMoments), v(ref.normal), v(ref.depthStencil), v(out_img), 0x80000000, None) == 0
    assert out_img.levels_equal(ref.indirectLight, 0)  # counting does not change the result
    c = (C.c_ulonglong * 3)()
    port.lib.orc_gi_gather_counters(c)
    pixels, samples, hits = (int(x) for x in c)
    assert pixels == W * Hh and 0 < hits < samples
    # numpy restatement of BoxRayCast + iterationsCount in fp32
    f = np.float32
    x, y = np.meshgrid(np.arange(W, dtype=np.float32) + f(0.5), np.arange(Hh, dtype=np.float32) + f(0.5))
    idx = (x.astype(np.int32) % 4) + (y.astype(np.int32) % 4) * 4
    total = 0
    for d in range(4):
        ang = (f(1.57075) * (idx.astype(np.float32) / f(16.0))) + f(1.57075) * f(d)
        dx, dy = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
        with np.errstate(divide="ignore"):
            ivx, ivy = f(1.0) / dx, f(1.0) / dy
        t1, t2, t3, t4 = (f(0.0) - x) * ivx, (f(W) - x) * ivx, (f(0.0) - y) * ivy, (f(Hh) - y) * ivy
        path = np.abs(np.minimum(np.maximum(t1, t2), np.maximum(t3, t4))).astype(np.float32)
        near = f(W) / f(1000.0)
        its = (np.log(path / near).astype(np.float32) / f(0.944197714328765869140625)).astype(np.int32) + 1
        total += int(np.maximum(its, 0).sum())
    assert abs(total - samples) <= 0.002 * samples  # cosf/sinf/logf of numpy vs libm may move a boundary pixel by one step


def test_port_reproduces_the_bundled_scene_fixture():
    """tests/golden/bundled_sponza_192x108.npz: the reference's bundled sample scene (BASELINE configs[0]: SponzaScene.json, 466k
    triangles) rasterised by the oracle's rasteriser and shaded by the reference's own SPIR-V passes. The port must reproduce every
    image and level bit for bit from the fixture's inputs."""
    sc, p, want, _ = H.load_golden(H.BUNDLED_GOLDEN)
    got = _run(loader.port(), sc, p, sc.width, sc.height, want.shadow_size)
    levels = passes.mip_levels_built(sc.width, sc.height)
    for iname, img, l in _all_levels(want):
        if iname == "shadowMap" or (l >= levels and iname in ("directLight", "depthMoments")):
            continue
        assert getattr(got, iname).levels_equal(img, l), f"{iname} level {l} differs from the bundled-scene fixture"
    # it really is the bundled scene: both emissive hornbugs and the atrium are in view, and every pixel is covered
    ids = sc.fragments["objectId"]
    assert set(np.unique(ids).tolist()) == {0, 1, 4} and (ids != abi.LGCU_NO_OBJECT).all()


@pytest.mark.skipif(not __import__("tests.bundled_scene", fromlist=["available"]).available(), reason="needs the reference's bundled scene under /root/reference")
def test_bundled_scene_fixture_regenerates_from_the_reference_assets():
    """Loader (tests/bundled_scene.py: Scene.h / Mesh.h semantics) + oracle rasteriser reproduce the fixture's fragment buffer, object
    table and shadow map from the OBJ / JSON files where they lie."""
    from legitengine_b200 import raster
    from tests import bundled_scene as B

    z = np.load(H.GOLDEN_DIR / f"{H.BUNDLED_GOLDEN}.npz")
    _, W, Hh, _, _, shadow = (int(v) for v in z["meta"])
    mesh = B.load_bundled_scene()
    assert mesh.triangle_count == int(z["triangles"][0]) and len(mesh.draws) == 5
    assert np.array_equal(mesh.objects.view(np.uint8).reshape(-1), z["objects"])
    m = scene.frame_matrices(W, Hh)
    port, ms = loader.port(), raster.host_mesh_desc(mesh)
    frags = np.zeros((Hh, W), dtype=abi.FRAGMENT_DTYPE)
    g = abi.GBufferBuilderData(abi.mat4(m.view), abi.mat4(m.proj), 0.0, 0.0)
    assert port.raster_gbuffer(C.byref(g), C.byref(ms), W, Hh, frags.ctypes.data, frags.strides[0], None) == 0
    assert np.array_equal(frags.view(np.uint8).reshape(Hh, W * 32), z["fragments"])
    depth = np.zeros((shadow, shadow), dtype=np.float32)
    sp = abi.ShadowmapBuilderData(abi.mat4(m.light_view), abi.mat4(m.light_proj))
    assert port.raster_shadow_map(C.byref(sp), C.byref(ms), shadow, depth.ctypes.data, depth.strides[0]) == 0
    assert np.array_equal(depth.view(np.uint32), z["shadow_map"].view(np.uint32))
