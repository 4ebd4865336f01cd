"""ctypes loaders for the two CPU oracles — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  port() -> oracle/liboracle_port.so    plain-C restatement (oracle/ssvgi_oracle.c), symbols orc_<pass>
  ref()  -> oracle/_ref/libref_spirv.so the reference's own SPIR-V run through its vendored SPIRV-Cross C++ backend
                                        and GLM (built by oracle/Makefile from /root/reference), symbols ref_<pass>
Both expose the pass signatures of include/lgcu.h minus the trailing stream argument, on HOST images.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from legitengine_b200 import abi

HERE = Path(__file__).resolve().parent
PORT_PATH = HERE / "liboracle_port.so"
REF_PATH = HERE / "_ref" / "libref_spirv.so"


class Oracle:
    def __init__(self, lib: C.CDLL, prefix: str, kind: str):
        self.lib, self.prefix, self.kind = lib, prefix, kind
        for name, sig in abi.PASS_SIGNATURES.items():
            fn = getattr(lib, f"{prefix}_{name}")
            fn.argtypes = sig
            fn.restype = C.c_int
            setattr(self, name, fn)
        self.num_threads = getattr(lib, f"{prefix}_num_threads")
        self.num_threads.restype = C.c_int
        self.set_num_threads = getattr(lib, f"{prefix}_set_num_threads")
        self.set_num_threads.argtypes = [C.c_int]
        if kind == "port":  # rasterisation front end (oracle/raster_oracle.c): host scene arrays, host targets
            lib.orc_raster_shadow_map.argtypes = [C.POINTER(abi.ShadowmapBuilderData), C.POINTER(abi.MeshScene), C.c_uint32, C.c_void_p, C.c_uint64]
            lib.orc_raster_gbuffer.argtypes = [C.POINTER(abi.GBufferBuilderData), C.POINTER(abi.MeshScene), C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, abi.ROWS]
            lib.orc_vertex_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
            for fn in (lib.orc_raster_shadow_map, lib.orc_raster_gbuffer, lib.orc_vertex_stage):
                fn.restype = C.c_int
            self.raster_shadow_map, self.raster_gbuffer, self.vertex_stage = lib.orc_raster_shadow_map, lib.orc_raster_gbuffer, lib.orc_vertex_stage


_port = None
_ref = None


def port() -> Oracle:
    global _port
    if _port is None:
        if not PORT_PATH.exists():
            raise FileNotFoundError(f"{PORT_PATH} missing: run `make -C oracle port` (or __graft_entry__.build())")
        _port = Oracle(C.CDLL(str(PORT_PATH)), "orc", "port")
    return _port


def have_ref() -> bool:
    return REF_PATH.exists()


def ref() -> Oracle:
    global _ref
    if _ref is None:
        if not REF_PATH.exists():
            raise FileNotFoundError(f"{REF_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        lib = C.CDLL(str(REF_PATH))
        _ref = Oracle(lib, "ref", "reference")
        f4 = C.POINTER(C.c_float)
        lib.ref_frame_matrices.argtypes = [f4, C.c_float, C.c_float, f4, C.c_float, C.c_float, C.c_uint, C.c_uint, f4, f4, f4, f4]
        lib.ref_frame_matrices.restype = None
        lib.ref_mat4_inverse.argtypes = [f4, f4]
        lib.ref_mat4_mul.argtypes = [f4, f4, f4]
        for fn in (lib.ref_gbuffer_vertex_stage, lib.ref_shadowmap_vertex_stage):  # (DrawCallData*, UBO*, vertices, n, out[n][10])
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
            fn.restype = C.c_int
    return _ref
