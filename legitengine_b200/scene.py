"""Synthetic scenes and frame matrices for the SSVGI path (thin ctypes wrapper over host/synth_scene.cpp).

The scene stands in for the reference's rasteriser + Scene (src/Scene/Scene.h): it produces the per-pixel fragment
buffer that feeds the G-buffer resolve, the per-draw-call constants, and the light's depth map (ShadowPass output).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi

SHADOW_MAP_SIZE = 1024  # SSVGIRenderer.h:402
DEFAULT_BOXES = 64

# src/main.cpp:166-172
DEFAULT_CAMERA = dict(pos=(0.0, 0.5, -2.0), vert=0.0, hor=0.0)
DEFAULT_LIGHT = dict(pos=(0.0, 5.0, 0.0), vert=float(np.float32(3.1415) / np.float32(2.0)), hor=0.0)


def _f4(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@dataclass
class FrameMatrices:
    view: np.ndarray  # column-major float32[16]
    proj: np.ndarray
    light_view: np.ndarray
    light_proj: np.ndarray


def frame_matrices(width: int, height: int, camera=None, light=None) -> FrameMatrices:
    """view / proj / lightView / lightProj exactly as SSVGIRenderer::RenderFrame builds them (SSVGIRenderer.h:54-59)."""
    lib = abi.load_scene_lib()
    camera = camera or DEFAULT_CAMERA
    light = light or DEFAULT_LIGHT
    cp = np.asarray(camera["pos"], dtype=np.float32)
    lp = np.asarray(light["pos"], dtype=np.float32)
    out = [np.zeros(16, dtype=np.float32) for _ in range(4)]
    lib.lgs_frame_matrices(_f4(cp), camera["vert"], camera["hor"], _f4(lp), light["vert"], light["hor"], width, height, *[_f4(o) for o in out])
    return FrameMatrices(*out)


@dataclass
class Scene:
    width: int
    height: int
    seed: int
    matrices: FrameMatrices
    fragments: np.ndarray  # (height, width) structured abi.FRAGMENT_DTYPE
    objects: np.ndarray  # (n,) structured abi.DRAW_CALL_DTYPE
    shadow_map: np.ndarray  # (1024, 1024) float32


def scene_objects(seed: int, n_boxes: int = DEFAULT_BOXES) -> np.ndarray:
    lib = abi.load_scene_lib()
    n = lib.lgs_object_count(n_boxes)
    objects = np.zeros(n, dtype=abi.DRAW_CALL_DTYPE)
    abi.check(lib.lgs_scene_objects(seed, n_boxes, objects.ctypes.data, n), "lgs_scene_objects")
    return objects


def scene_fragments(seed: int, width: int, height: int, m: FrameMatrices, n_boxes: int = DEFAULT_BOXES, rows=None, out: np.ndarray | None = None) -> np.ndarray:
    lib = abi.load_scene_lib()
    frags = out if out is not None else np.zeros((height, width), dtype=abi.FRAGMENT_DTYPE)
    y0, y1 = rows if rows is not None else (0, height)
    abi.check(
        lib.lgs_scene_fragments(seed, n_boxes, width, height, _f4(m.view), _f4(m.proj), frags.ctypes.data, frags.strides[0], y0, y1),
        "lgs_scene_fragments",
    )
    return frags


def scene_shadow_map(seed: int, m: FrameMatrices, n_boxes: int = DEFAULT_BOXES, size: int = SHADOW_MAP_SIZE) -> np.ndarray:
    lib = abi.load_scene_lib()
    depth = np.zeros((size, size), dtype=np.float32)
    abi.check(lib.lgs_scene_shadow_map(seed, n_boxes, size, _f4(m.light_view), _f4(m.light_proj), depth.ctypes.data, depth.strides[0]), "lgs_scene_shadow_map")
    return depth


def make_scene(seed: int, width: int, height: int, n_boxes: int = DEFAULT_BOXES, camera=None, light=None, shadow_size: int = SHADOW_MAP_SIZE) -> Scene:
    m = frame_matrices(width, height, camera, light)
    return Scene(width, height, seed, m, scene_fragments(seed, width, height, m, n_boxes), scene_objects(seed, n_boxes), scene_shadow_map(seed, m, n_boxes, shadow_size))


@dataclass
class Mesh:
    """The scene as the reference's Scene holds it (src/Scene/Scene.h, src/Scene/Mesh.h:209-216): input of lgcu_raster_*."""
    vertices: np.ndarray  # (nv,) abi.VERTEX_DTYPE
    indices: np.ndarray  # (ni,) uint32
    draws: np.ndarray  # (nd,) abi.DRAW_DTYPE, firstTriangle filled
    objects: np.ndarray  # (no,) abi.DRAW_CALL_DTYPE

    @property
    def triangle_count(self) -> int:
        return int(self.draws["indexCount"].sum() // 3)


def scene_mesh(seed: int, n_boxes: int = DEFAULT_BOXES) -> Mesh:
    lib = abi.load_scene_lib()
    nv, ni, nd = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib.lgs_scene_mesh_counts(n_boxes, C.byref(nv), C.byref(ni), C.byref(nd))
    vertices = np.zeros(nv.value, dtype=abi.VERTEX_DTYPE)
    indices = np.zeros(ni.value, dtype=np.uint32)
    draws = np.zeros(nd.value, dtype=abi.DRAW_DTYPE)
    abi.check(lib.lgs_scene_mesh(seed, n_boxes, vertices.ctypes.data, indices.ctypes.data, draws.ctypes.data), "lgs_scene_mesh")
    return Mesh(vertices, indices, draws, scene_objects(seed, n_boxes))


def load_packed_mesh(path) -> "Mesh":
    """A mesh scene stored as .npz (vertices / indices / draws / objects as raw bytes of the include/lgcu.h structs) -> Mesh."""
    z = np.load(path)
    return Mesh(z["vertices"].view(abi.VERTEX_DTYPE).copy(), z["indices"].astype(np.uint32), z["draws"].view(abi.DRAW_DTYPE).copy(), z["objects"].view(abi.DRAW_CALL_DTYPE).copy())
