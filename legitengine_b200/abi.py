"""ctypes mirror of include/lgcu.h (the C ABI of the CUDA pass library) and loaders for the in-tree shared objects.

Nothing here computes anything: the structures are byte-for-byte the C ones (checked by tests/test_abi.py against
sizes the C side reports) and `load_lgcu()` fails loudly when the CUDA library has not been built — there is no
CPU fallback for the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_DIR = PKG_DIR / "lib"

LGCU_MAX_MIPS = 16
LGCU_NO_OBJECT = 0xFFFFFFFF

# lgcu_status
LGCU_OK = 0
LGCU_ERR_INVALID_ARGUMENT = -1
LGCU_ERR_UNSUPPORTED_FORMAT = -2
LGCU_ERR_CUDA = -3
LGCU_ERR_UNSUPPORTED = -4

# lgcu_format (VkFormat values)
FORMAT_B8G8R8A8_SRGB = 50
FORMAT_R16G16B16A16_SFLOAT = 97
FORMAT_R32G32_SFLOAT = 103
FORMAT_R32G32B32A32_SFLOAT = 109
FORMAT_D32_SFLOAT = 126

TEXEL_SIZE = {
    FORMAT_B8G8R8A8_SRGB: 4,
    FORMAT_R16G16B16A16_SFLOAT: 8,
    FORMAT_R32G32_SFLOAT: 8,
    FORMAT_R32G32B32A32_SFLOAT: 16,
    FORMAT_D32_SFLOAT: 4,
}

GI_DEFAULT = 0
GI_STRICT = 1


class LgcuImage(C.Structure):
    _fields_ = [
        ("base", C.c_void_p),
        ("format", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("imageMipCount", C.c_uint32),
        ("baseMip", C.c_uint32),
        ("mipCount", C.c_uint32),
        ("reserved0", C.c_uint32),
        ("reserved1", C.c_uint32),
        ("levelOffset", C.c_uint64 * LGCU_MAX_MIPS),
        ("levelPitch", C.c_uint32 * LGCU_MAX_MIPS),
    ]


class LgcuRows(C.Structure):
    _fields_ = [("y0", C.c_uint32), ("y1", C.c_uint32)]


class LgcuMat4(C.Structure):
    _pack_ = 1
    _fields_ = [("m", C.c_float * 16)]


def _packed(name, fields):
    return type(name, (C.Structure,), {"_pack_": 1, "_fields_": fields})


GBufferBuilderData = _packed("GBufferBuilderData", [("viewMatrix", LgcuMat4), ("projMatrix", LgcuMat4), ("time", C.c_float), ("bla", C.c_float)])
DrawCallData = _packed("DrawCallData", [("modelMatrix", LgcuMat4), ("albedoColor", C.c_float * 4), ("emissiveColor", C.c_float * 4)])
DirectLightingData = _packed("DirectLightingData", [("viewMatrix", LgcuMat4), ("projMatrix", LgcuMat4), ("lightViewMatrix", LgcuMat4), ("lightProjMatrix", LgcuMat4), ("time", C.c_float)])
MipLevelBuilderData = _packed("MipLevelBuilderData", [("filterType", C.c_float)])
BlurLayerBuilderData = _packed("BlurLayerBuilderData", [("size", C.c_int32 * 4), ("radius", C.c_int32)])
IndirectLightingData = _packed("IndirectLightingData", [("viewMatrix", LgcuMat4), ("projMatrix", LgcuMat4), ("viewportExtent", C.c_float * 4)])
DenoiserData = _packed("DenoiserData", [("viewMatrix", LgcuMat4), ("projMatrix", LgcuMat4), ("viewportExtent", C.c_float * 4), ("radius", C.c_int32)])
FinalGathererData = _packed("FinalGathererData", [("viewMatrix", LgcuMat4), ("projMatrix", LgcuMat4)])
InterleaveData = _packed("InterleaveData", [("gridSize", C.c_int32 * 4), ("viewportSize", C.c_int32 * 4)])
DebugQuadData = _packed("DebugQuadData", [("minmax", C.c_float * 4)])


class RowCopy(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("bytes", C.c_uint64)]


class ExchangeDesc(C.Structure):
    """lgcu_exchange_desc (include/lgcu.h): one fused exchange step of the strip protocol."""
    _fields_ = [("signalBefore", C.POINTER(C.c_void_p)), ("signalBeforeCount", C.c_uint32), ("wait", C.POINTER(C.c_void_p)), ("waitCount", C.c_uint32),
                ("lag", C.c_uint32), ("copies", C.POINTER(RowCopy)), ("copyCount", C.c_uint32), ("signalAfter", C.POINTER(C.c_void_p)),
                ("signalAfterCount", C.c_uint32), ("frameCounter", C.c_void_p), ("doneCounter", C.c_void_p), ("bump", C.c_uint32)]


class ClearValues(C.Structure):
    _fields_ = [("color", C.c_float * 4), ("depth", C.c_float)]


def default_clear() -> ClearValues:
    """Default attachment clear values of the G-buffer pass (LV/RenderGraph.h:469, 487)."""
    return ClearValues((C.c_float * 4)(1.0, 0.5, 0.0, 1.0), 1.0)


# numpy views of the two array-of-struct inputs
FRAGMENT_DTYPE = np.dtype(
    [("worldPos", "<f4", 3), ("worldNormal", "<f4", 3), ("objectId", "<u4"), ("ndcDepth", "<f4")]
)
DRAW_CALL_DTYPE = np.dtype([("modelMatrix", "<f4", 16), ("albedoColor", "<f4", 4), ("emissiveColor", "<f4", 4)])
# rasterisation front end (include/lgcu.h: lgcu_vertex, lgcu_draw, lgcu_mesh_scene, lgcu_shadowmap_builder_data)
VERTEX_DTYPE = np.dtype([("pos", "<f4", 3), ("normal", "<f4", 3), ("uv", "<f4", 2)])
DRAW_DTYPE = np.dtype([("firstIndex", "<u4"), ("indexCount", "<u4"), ("vertexOffset", "<u4"), ("objectId", "<u4"), ("firstTriangle", "<u4"), ("reserved", "<u4", 3)])
assert FRAGMENT_DTYPE.itemsize == 32 and DRAW_CALL_DTYPE.itemsize == 96 and VERTEX_DTYPE.itemsize == 32 and DRAW_DTYPE.itemsize == 32
ShadowmapBuilderData = _packed("ShadowmapBuilderData", [("lightViewMatrix", LgcuMat4), ("lightProjMatrix", LgcuMat4)])


class MeshScene(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("indices", C.c_void_p), ("draws", C.c_void_p), ("objects", C.c_void_p),
                ("nVertices", C.c_uint32), ("nIndices", C.c_uint32), ("nDraws", C.c_uint32), ("nObjects", C.c_uint32), ("nTriangles", C.c_uint32)]


def mat4(values) -> LgcuMat4:
    m = LgcuMat4()
    flat = np.asarray(values, dtype=np.float32).reshape(16)
    for i in range(16):
        m.m[i] = float(flat[i])
    return m


P = C.POINTER
IMG = P(LgcuImage)
ROWS = P(LgcuRows)

# rasterisation front end: (argtypes without the trailing stream); the oracle variants take host arrays and different targets
RASTER_SIGNATURES = {
    "raster_shadow_map": [P(ShadowmapBuilderData), P(MeshScene), C.c_void_p, C.c_uint64, IMG],
    "raster_gbuffer": [P(GBufferBuilderData), P(MeshScene), C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, ROWS],
}

# name -> (argtypes without the trailing stream) ; shared by the CUDA library (with stream) and both oracles (without)
PASS_SIGNATURES = {
    "gbuffer_resolve": [P(GBufferBuilderData), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, P(ClearValues), IMG, IMG, IMG, IMG, IMG, ROWS],
    "direct_light": [P(DirectLightingData), IMG, IMG, IMG, IMG, IMG, IMG, ROWS],
    "mip_level": [P(MipLevelBuilderData), IMG, IMG, ROWS],
    "blur_level": [P(BlurLayerBuilderData), IMG, IMG, ROWS],
    "gi_gather": [P(IndirectLightingData), IMG, IMG, IMG, IMG, IMG, C.c_uint32, ROWS],
    "denoise": [P(DenoiserData), IMG, IMG, IMG, IMG, ROWS],
    "final_gather": [P(FinalGathererData), IMG, IMG, IMG, IMG, IMG, ROWS],
    # interleaved rendering and the debug overlay
    "deinterleave": [P(InterleaveData), IMG, IMG, ROWS],
    "interleave": [P(InterleaveData), IMG, IMG, ROWS],
    "debug_overlay": [P(DebugQuadData), IMG, IMG, ROWS],
}
FUSED_SIGNATURES = {
    "gbuffer_direct_light": [P(GBufferBuilderData), P(DirectLightingData), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, P(ClearValues), IMG, IMG, IMG, IMG, IMG, IMG, IMG, ROWS],
    "mip_blur_chain": [IMG, IMG, C.c_int32, ROWS],
    "denoise_final_gather": [P(DenoiserData), P(FinalGathererData), IMG, IMG, IMG, IMG, IMG, IMG, IMG, IMG, ROWS],
    "frame_front": [P(GBufferBuilderData), P(DirectLightingData), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, P(ClearValues), IMG, IMG, IMG, IMG, IMG, IMG, IMG, IMG, IMG, ROWS],
    "frame_chains": [IMG, IMG, IMG, IMG, C.c_int32, ROWS],
    "gi_gather_pack": [P(IndirectLightingData), IMG, IMG, IMG, IMG, IMG, C.c_void_p, C.c_uint64, ROWS],
    "gi_gather_packed": [P(IndirectLightingData), IMG, IMG, IMG, IMG, IMG, C.c_void_p, C.c_uint64, ROWS],
}


class LibraryMissing(RuntimeError):
    pass


def _load(path: Path, what: str, how: str) -> C.CDLL:
    if not path.exists():
        raise LibraryMissing(f"{what} not built: {path} is missing ({how})")
    return C.CDLL(str(path), mode=getattr(os, "RTLD_NOW", 2) | getattr(os, "RTLD_GLOBAL", 0))


_lgcu = None


def lgcu_path() -> Path:
    return LIB_DIR / "liblgcu.so"


def load_lgcu() -> C.CDLL:
    """The CUDA pass library. Raises LibraryMissing if it has not been built — never falls back to the CPU."""
    global _lgcu
    if _lgcu is None:
        lib = _load(lgcu_path(), "CUDA pass library (liblgcu.so)", "run `python -c 'import __graft_entry__ as g; g.build()'`")
        for name, sig in {**PASS_SIGNATURES, **FUSED_SIGNATURES, **RASTER_SIGNATURES}.items():
            fn = getattr(lib, "lgcu_" + name)
            fn.argtypes = sig + [C.c_void_p]
            fn.restype = C.c_int
        lib.lgcu_abi_version.restype = C.c_int
        lib.lgcu_last_error.restype = C.c_char_p
        lib.lgcu_format_texel_size.argtypes = [C.c_uint32]
        lib.lgcu_format_texel_size.restype = C.c_uint32
        lib.lgcu_image_layout.argtypes = [IMG, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.lgcu_image_layout.restype = C.c_uint64
        lib.lgcu_copy_rows.argtypes = [P(RowCopy), C.c_uint32, C.c_void_p]
        lib.lgcu_frame_counter_bump.argtypes = [C.c_void_p, C.c_void_p]
        lib.lgcu_signal_flags.argtypes = [P(C.c_void_p), C.c_uint32, C.c_void_p, C.c_void_p]
        lib.lgcu_wait_flags.argtypes = [P(C.c_void_p), C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.lgcu_exchange.argtypes = [P(ExchangeDesc), C.c_void_p]
        for fn in (lib.lgcu_copy_rows, lib.lgcu_frame_counter_bump, lib.lgcu_signal_flags, lib.lgcu_wait_flags, lib.lgcu_exchange):
            fn.restype = C.c_int
        lib.lgcu_gather_scratch_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        lib.lgcu_gather_scratch_bytes.restype = C.c_uint64
        lib.lgcu_raster_scratch_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        lib.lgcu_raster_scratch_bytes.restype = C.c_uint64
        lib.lgcu_raster_prepare_draws.argtypes = [C.c_void_p, C.c_uint32]
        lib.lgcu_raster_prepare_draws.restype = C.c_uint32
        _lgcu = lib
    return _lgcu


_scene = None


def load_scene_lib() -> C.CDLL:
    global _scene
    if _scene is None:
        lib = _load(LIB_DIR / "liblgcu_scene.so", "synthetic scene library", "run build()")
        f4 = P(C.c_float)
        lib.lgs_object_count.argtypes = [C.c_uint32]
        lib.lgs_object_count.restype = C.c_uint32
        lib.lgs_scene_objects.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32]
        lib.lgs_scene_fragments.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, f4, f4, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
        lib.lgs_scene_shadow_map.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, f4, f4, C.c_void_p, C.c_uint64]
        lib.lgs_frame_matrices.argtypes = [f4, C.c_float, C.c_float, f4, C.c_float, C.c_float, C.c_uint32, C.c_uint32, f4, f4, f4, f4]
        lib.lgs_frame_matrices.restype = None
        lib.lgs_mat4_inverse.argtypes = [f4, f4]
        lib.lgs_mat4_mul.argtypes = [f4, f4, f4]
        u32p = P(C.c_uint32)
        lib.lgs_scene_mesh_counts.argtypes = [C.c_uint32, u32p, u32p, u32p]
        lib.lgs_scene_mesh_counts.restype = None
        lib.lgs_scene_mesh.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _scene = lib
    return _scene


def check(status: int, what: str = "lgcu call") -> None:
    if status != LGCU_OK:
        detail = ""
        if _lgcu is not None:
            detail = (_lgcu.lgcu_last_error() or b"").decode()
        raise RuntimeError(f"{what} failed with status {status} {detail}")
