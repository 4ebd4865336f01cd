"""CPU tests of the rasterisation front end's oracle (oracle/raster_oracle.c) — run without a GPU.

 * vertex stage pinned bit-for-bit against the reference's shipped vertex SPIR-V (oracle/_ref, when built);
 * rule R validated against an INDEPENDENT implementation: the analytic ray caster of the synthetic scene
   (legitengine_b200/host/synth_scene.cpp) must see the same object in (almost) every pixel, at the same depth / position / normal;
 * draw-order and clipping properties of the rule.
"""
import ctypes as C

import numpy as np
import pytest

from legitengine_b200 import abi, raster, scene
from oracle import loader


def _gparams(m):
    return abi.GBufferBuilderData(abi.mat4(m.view), abi.mat4(m.proj), 0.0, 0.0)


def _oracle_fragments(mesh, m, W, H, rows=None):
    frags = np.zeros((H, W), dtype=abi.FRAGMENT_DTYPE)
    g, ms = _gparams(m), raster.host_mesh_desc(mesh)
    r = C.byref(abi.LgcuRows(*rows)) if rows else None
    assert loader.port().raster_gbuffer(C.byref(g), C.byref(ms), W, H, frags.ctypes.data, frags.strides[0], r) == 0
    return frags


@pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("stage", ["gbuffer", "shadowmap"])
def test_vertex_stage_matches_reference_spirv(stage):
    """gBufferBuilder.vert.spv / shadowmapBuilder.vert.spv through the reference's SPIRV-Cross + GLM vs the C restatement."""
    rng = np.random.default_rng(5)
    n = 4096
    verts = np.zeros(n, dtype=abi.VERTEX_DTYPE)
    verts["pos"] = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    verts["normal"] = rng.normal(size=(n, 3)).astype(np.float32)
    m = scene.frame_matrices(1920, 1080, camera=dict(pos=(0.3, 1.5, -4.0), vert=0.2, hor=-0.4))
    obj = np.zeros(1, dtype=abi.DRAW_CALL_DTYPE)
    model = np.eye(4, dtype=np.float32)
    model[:3, :3] = rng.normal(size=(3, 3)).astype(np.float32)
    model[3, :3] = (0.5, -1.25, 2.0)  # column-major storage: row 3 of the array is the translation column
    obj["modelMatrix"][0] = model.reshape(16)
    view, proj = (m.view, m.proj) if stage == "gbuffer" else (m.light_view, m.light_proj)
    ubo = abi.GBufferBuilderData(abi.mat4(view), abi.mat4(proj), 0.0, 0.0)  # ShadowmapBuilderData is its first 128 bytes
    want = np.zeros((n, 10), dtype=np.float32)
    fn = loader.ref().lib.ref_gbuffer_vertex_stage if stage == "gbuffer" else loader.ref().lib.ref_shadowmap_vertex_stage
    assert fn(obj.ctypes.data, C.addressof(ubo), verts.ctypes.data, n, want.ctypes.data) == 0
    got = np.zeros((n, 10), dtype=np.float32)
    assert loader.port().vertex_stage(obj["modelMatrix"][0].ctypes.data, view.ctypes.data, proj.ctypes.data, verts.ctypes.data, n, got.ctypes.data) == 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("size,seed", [((640, 360), 7), ((512, 512), 3), ((960, 540), 11)])
def test_rule_r_agrees_with_the_ray_caster(size, seed):
    W, H = size
    m = scene.frame_matrices(W, H)
    mesh = scene.scene_mesh(seed)
    frags = _oracle_fragments(mesh, m, W, H)
    ref = scene.scene_fragments(seed, W, H, m)
    same = frags["objectId"] == ref["objectId"]
    assert same.mean() >= 0.9995, f"object ids differ on {(~same).sum()} pixels"  # silhouette pixels may fall either way
    cov = same & (ref["objectId"] != abi.LGCU_NO_OBJECT)
    assert cov.mean() > 0.5
    assert np.abs(frags["ndcDepth"][cov] - ref["ndcDepth"][cov]).max() <= 1e-6
    assert np.abs(frags["worldPos"][cov] - ref["worldPos"][cov]).max() <= 2e-5
    assert np.abs(frags["worldNormal"][cov] - ref["worldNormal"][cov]).max() <= 1e-6
    unc = frags["objectId"] == abi.LGCU_NO_OBJECT
    assert np.all(frags["ndcDepth"][unc] == 1.0)


def test_shadow_map_agrees_with_the_ray_caster():
    m = scene.frame_matrices(512, 512)
    mesh = scene.scene_mesh(7)
    sm = np.zeros((1024, 1024), dtype=np.float32)
    p = abi.ShadowmapBuilderData(abi.mat4(m.light_view), abi.mat4(m.light_proj))
    ms = raster.host_mesh_desc(mesh)
    assert loader.port().raster_shadow_map(C.byref(p), C.byref(ms), 1024, sm.ctypes.data, sm.strides[0]) == 0
    want = scene.scene_shadow_map(7, m)
    d = np.abs(sm - want)
    assert (d > 1e-5).mean() <= 1e-3 and np.median(d) <= 1e-6


def test_rows_are_independent():
    """A row strip produces exactly the rows of the whole frame (the property the strip-sharded renderer relies on)."""
    W, H = 320, 192
    m = scene.frame_matrices(W, H)
    mesh = scene.scene_mesh(2, n_boxes=16)
    whole = _oracle_fragments(mesh, m, W, H)
    strip = _oracle_fragments(mesh, m, W, H, rows=(64, 128))
    assert np.array_equal(strip[64:128].view(np.uint8), whole[64:128].view(np.uint8))


def _two_quads(z_first, z_second):
    """Two screen-filling quads at view depths z_first / z_second drawn in that order (objects 0 and 1)."""
    verts = np.zeros(8, dtype=abi.VERTEX_DTYPE)
    for q, z in enumerate((z_first, z_second)):
        verts["pos"][4 * q:4 * q + 4] = [(-50, -50, z), (50, -50, z), (50, 50, z), (-50, 50, z)]
        verts["normal"][4 * q:4 * q + 4] = (0, 0, -1)
    idx = np.array([0, 1, 2, 0, 2, 3, 4, 5, 6, 4, 6, 7], dtype=np.uint32)
    draws = np.zeros(2, dtype=abi.DRAW_DTYPE)
    draws["firstIndex"], draws["indexCount"], draws["objectId"], draws["firstTriangle"] = (0, 6), 6, (0, 1), (0, 2)
    objs = np.zeros(2, dtype=abi.DRAW_CALL_DTYPE)
    objs["modelMatrix"][:] = np.eye(4, dtype=np.float32).reshape(16)
    return scene.Mesh(verts, idx, draws, objs)


def test_equal_depth_keeps_the_first_draw_and_nearer_wins():
    W, H = 64, 48
    m = scene.frame_matrices(W, H)
    same = _oracle_fragments(_two_quads(3.0, 3.0), m, W, H)
    assert np.all(same["objectId"] == 0)  # depth test LESS: the second, equal-depth draw fails
    nearer = _oracle_fragments(_two_quads(3.0, 1.0), m, W, H)
    assert np.all(nearer["objectId"] == 1)
    farther = _oracle_fragments(_two_quads(1.0, 3.0), m, W, H)
    assert np.all(farther["objectId"] == 0)


def test_geometry_behind_the_near_plane_or_beyond_the_far_plane_is_clipped():
    W, H = 64, 48
    m = scene.frame_matrices(W, H)  # camera at z = -2 looking down +z, near 0.01, far 1000
    behind = _oracle_fragments(_two_quads(-5.0, -2.005), m, W, H)
    assert np.all(behind["objectId"] == abi.LGCU_NO_OBJECT)
    beyond = _oracle_fragments(_two_quads(1200.0, 5000.0), m, W, H)
    assert np.all(beyond["objectId"] == abi.LGCU_NO_OBJECT)


def test_every_pixel_of_a_shared_edge_is_covered_exactly_once():
    """Top-left rule: two triangles sharing the diagonal of a quad cover each pixel once (no cracks, no double hits) — checked by
    drawing the quad's triangles as two objects in both orders and requiring complementary ownership."""
    W, H = 96, 64
    m = scene.frame_matrices(W, H)
    verts = np.zeros(4, dtype=abi.VERTEX_DTYPE)
    verts["pos"] = [(-0.83, -0.31, 1.0), (0.91, -0.47, 1.0), (0.77, 1.23, 1.0), (-0.69, 1.11, 1.0)]
    objs = np.zeros(2, dtype=abi.DRAW_CALL_DTYPE)
    objs["modelMatrix"][:] = np.eye(4, dtype=np.float32).reshape(16)

    def draw(order):
        idx = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32).reshape(2, 3)[list(order)].reshape(-1)
        draws = np.zeros(2, dtype=abi.DRAW_DTYPE)
        draws["firstIndex"], draws["indexCount"], draws["objectId"], draws["firstTriangle"] = (0, 3), 3, order, (0, 1)
        return _oracle_fragments(scene.Mesh(verts, idx, draws, objs), m, W, H)["objectId"]

    a, b = draw((0, 1)), draw((1, 0))
    assert np.array_equal(a, b)  # ownership does not depend on draw order => no pixel is covered by both triangles
    covered = a != abi.LGCU_NO_OBJECT
    assert covered.sum() > 200 and set(np.unique(a[covered])) == {0, 1}
