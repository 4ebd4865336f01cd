"""Oracle-side frame builders — TEST INFRASTRUCTURE (used by tests/ and by __graft_entry__.smoke(), never by the product path)."""
from __future__ import annotations

import ctypes as C
import functools
from typing import Tuple

import numpy as np

from legitengine_b200 import abi, images, passes, scene
from oracle import loader


# ---- mesh scenes: the oracle's rasteriser (oracle/raster_oracle.c) in front of the oracle's fragment passes ----
@functools.lru_cache(maxsize=4)
def oracle_mesh_frame(seed: int, width: int, height: int, camera_key: Tuple = None, n_boxes: int = 64, denoise_radius: int = 0):
    """Mesh + the scene it rasterises to (oracle rasteriser: fragments and shadow map) + the port oracle's frame images."""
    from legitengine_b200 import raster

    camera = dict(pos=camera_key[0], vert=camera_key[1], hor=camera_key[2]) if camera_key else None
    mesh = scene.scene_mesh(seed, n_boxes)
    return (mesh, camera) + oracle_frame_from_mesh(mesh, width, height, camera, denoise_radius, seed)


def oracle_frame_from_mesh(mesh, width: int, height: int, camera=None, denoise_radius: int = 0, seed: int = 0):
    """(scene the mesh rasterises to under the oracle's rasteriser, params, the port oracle's frame images)."""
    from legitengine_b200 import raster

    m = scene.frame_matrices(width, height, camera=camera)
    port, ms = loader.port(), raster.host_mesh_desc(mesh)
    frags = np.zeros((height, width), dtype=abi.FRAGMENT_DTYPE)
    g = abi.GBufferBuilderData(abi.mat4(m.view), abi.mat4(m.proj), 0.0, 0.0)
    assert port.raster_gbuffer(C.byref(g), C.byref(ms), width, height, frags.ctypes.data, frags.strides[0], None) == 0
    shadow = np.zeros((scene.SHADOW_MAP_SIZE, scene.SHADOW_MAP_SIZE), dtype=np.float32)
    sp = abi.ShadowmapBuilderData(abi.mat4(m.light_view), abi.mat4(m.light_proj))
    assert port.raster_shadow_map(C.byref(sp), C.byref(ms), scene.SHADOW_MAP_SIZE, shadow.ctypes.data, shadow.strides[0]) == 0
    sc = scene.Scene(width, height, seed, m, frags, mesh.objects, shadow)
    p = passes.make_params(width, height, m, denoise_radius)
    fi = passes.FrameImages(width, height, images.HostImage)
    passes.run_pass_list(port, fi, p, passes.upload_inputs(fi, sc))
    return sc, p, fi
