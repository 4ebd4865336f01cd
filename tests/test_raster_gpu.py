"""GPU parity of the rasterisation front end (csrc/k_raster.cu through lgcu_raster_*) against oracle/raster_oracle.c: bit-exact
fragments (object id, depth, interpolated position / normal) and shadow maps, at test sizes and at BASELINE's 4K."""
import ctypes as C

import numpy as np
import pytest

from legitengine_b200 import abi, images, raster, scene
from oracle import loader

pytestmark = pytest.mark.gpu

CAMERAS = {
    "default": None,
    "tilted": dict(pos=(1.5, 2.5, -5.0), vert=0.35, hor=-0.3),
    "inside": dict(pos=(-3.0, 0.8, 2.0), vert=-0.1, hor=1.9),
}


def _oracle_fragments(mesh, m, W, H, rows=None):
    frags = np.zeros((H, W), dtype=abi.FRAGMENT_DTYPE)
    g, ms = abi.GBufferBuilderData(abi.mat4(m.view), abi.mat4(m.proj), 0.0, 0.0), raster.host_mesh_desc(mesh)
    r = C.byref(abi.LgcuRows(*rows)) if rows else None
    assert loader.port().raster_gbuffer(C.byref(g), C.byref(ms), W, H, frags.ctypes.data, frags.strides[0], r) == 0
    return frags


def _cuda_fragments(mesh, m, W, H, rows=None, poison=0xCD):
    import torch

    dm = raster.DeviceMesh(mesh)
    frags = torch.full((H, W * 32), poison, dtype=torch.uint8, device="cuda:0")
    dm.raster_gbuffer(m.view, m.proj, W, H, frags, rows=rows)
    torch.cuda.synchronize()
    return frags.cpu().numpy().view(abi.FRAGMENT_DTYPE).reshape(H, W)


def _assert_fragments_equal(got, want):
    for field in ("objectId", "ndcDepth", "worldPos", "worldNormal"):
        a, b = got[field].view(np.uint32), want[field].view(np.uint32)
        bad = a != b
        assert not bad.any(), f"{field}: {int(bad.sum())} mismatching words, first at {np.argwhere(bad)[:4].tolist()}"


@pytest.mark.parametrize("size", [(256, 144), (640, 360), (1000, 563), (1920, 1080)])
@pytest.mark.parametrize("camera", ["default", "tilted", "inside"])
def test_gbuffer_raster_is_bit_exact(size, camera):
    W, H = size
    m = scene.frame_matrices(W, H, camera=CAMERAS[camera])
    mesh = scene.scene_mesh(5 + W % 7)
    _assert_fragments_equal(_cuda_fragments(mesh, m, W, H), _oracle_fragments(mesh, m, W, H))


def test_gbuffer_raster_4k_bit_exact():
    W, H = 3840, 2160
    m = scene.frame_matrices(W, H)
    mesh = scene.scene_mesh(0xC0FFEE)
    got = _cuda_fragments(mesh, m, W, H)
    _assert_fragments_equal(got, _oracle_fragments(mesh, m, W, H))
    # and the independent check: the analytic ray caster sees the same objects
    ref = scene.scene_fragments(0xC0FFEE, W, H, m)
    assert (got["objectId"] == ref["objectId"]).mean() >= 0.9999


def test_gbuffer_raster_row_strip_only_touches_its_rows():
    W, H = 640, 384
    m = scene.frame_matrices(W, H)
    mesh = scene.scene_mesh(3)
    got = _cuda_fragments(mesh, m, W, H, rows=(128, 256))
    want = _oracle_fragments(mesh, m, W, H)
    _assert_fragments_equal(got[128:256], want[128:256])
    assert np.all(got[:128].view(np.uint8) == 0xCD) and np.all(got[256:].view(np.uint8) == 0xCD)


@pytest.mark.parametrize("size", [256, 1024])
def test_shadow_map_raster_is_bit_exact(size):
    import torch

    m = scene.frame_matrices(512, 512)
    mesh = scene.scene_mesh(9)
    want = np.zeros((size, size), dtype=np.float32)
    p, ms = abi.ShadowmapBuilderData(abi.mat4(m.light_view), abi.mat4(m.light_proj)), raster.host_mesh_desc(mesh)
    assert loader.port().raster_shadow_map(C.byref(p), C.byref(ms), size, want.ctypes.data, want.strides[0]) == 0
    img = images.DeviceImage(abi.FORMAT_D32_SFLOAT, size, size)
    raster.DeviceMesh(mesh).raster_shadow_map(m.light_view, m.light_proj, img)
    torch.cuda.synchronize()
    got = img.to_host().level_f32(0)[..., 0]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_many_small_triangles_bit_exact():
    """A tessellated scene (every quad split into a 24 x 24 grid: 74k triangles, most smaller than a pixel tile) exercises the
    one-warp-per-triangle path and the visibility atomics."""
    W, H = 960, 540
    m = scene.frame_matrices(W, H)
    base = scene.scene_mesh(4, n_boxes=12)
    n = 24
    quads = base.vertices.reshape(-1, 4)
    u = (np.arange(n + 1, dtype=np.float64) / n)
    verts = np.zeros((len(quads), n + 1, n + 1), dtype=abi.VERTEX_DTYPE)
    p0, p1, p2, p3 = (quads["pos"][:, k].astype(np.float64) for k in range(4))
    uu, vv = u[None, :, None, None], u[None, None, :, None]
    pos = (p0[:, None, None] * (1 - uu) + p1[:, None, None] * uu) * (1 - vv) + (p3[:, None, None] * (1 - uu) + p2[:, None, None] * uu) * vv
    verts["pos"] = pos.astype(np.float32)
    verts["normal"] = quads["normal"][:, 0][:, None, None]
    cell = np.arange(n)
    i0 = (cell[:, None] * (n + 1) + cell[None, :]).reshape(-1)
    tri = np.stack([i0, i0 + (n + 1), i0 + (n + 1) + 1, i0, i0 + (n + 1) + 1, i0 + 1], axis=1).reshape(-1)
    idx = (tri[None, :] + (np.arange(len(quads)) * (n + 1) ** 2)[:, None]).reshape(-1).astype(np.uint32)
    draws = base.draws.copy()
    scale = (n * n * 6) // 6
    draws["firstIndex"] = base.draws["firstIndex"] * scale
    draws["indexCount"] = base.draws["indexCount"] * scale
    draws["firstTriangle"] = base.draws["firstTriangle"] * scale
    mesh = scene.Mesh(verts.reshape(-1), idx, draws, base.objects)
    assert mesh.triangle_count == len(quads) * n * n * 2
    _assert_fragments_equal(_cuda_fragments(mesh, m, W, H), _oracle_fragments(mesh, m, W, H))
