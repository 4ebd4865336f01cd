"""ctypes wrapper of the headless frame harness (include/lgcu_harness.h, host/harness.cpp).

`Renderer` owns one C++ legit_cuda::SSVGIRenderer + RenderGraph on a CUDA stream. All arithmetic happens in the CUDA
kernels of liblgcu.so; this module only moves pointers around. PyTorch supplies pinned host memory and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import abi, images

MODE_PASS_GRANULAR = 0
MODE_FUSED = 1
STAGE_FRONT, STAGE_CHAINS, STAGE_GATHER, STAGE_FINAL, STAGE_ALL = 1, 2, 4, 8, 15

IMAGE_NAMES = (
    "albedo", "emissive", "normal", "depthMoments", "blurredDepthMoments", "depthStencil", "directLight",
    "blurredDirectLight", "shadowMap", "indirectLight", "denoisedIndirectLight", "swapchain",
)

_lib = None


def load_harness() -> C.CDLL:
    global _lib
    if _lib is None:
        abi.load_lgcu()  # dependency, loaded RTLD_GLOBAL first so $ORIGIN lookups are not needed
        lib = abi._load(abi.LIB_DIR / "liblegit_cuda.so", "host harness library (liblegit_cuda.so)", "run __graft_entry__.build()")
        R = C.c_void_p
        lib.lgh_last_error.restype = C.c_char_p
        lib.lgh_create.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        lib.lgh_create.restype = R
        lib.lgh_destroy.argtypes = [R]
        lib.lgh_destroy.restype = None
        f3 = C.POINTER(C.c_float)
        lib.lgh_set_camera.argtypes = [R, f3, C.c_float, C.c_float, f3, C.c_float, C.c_float]
        lib.lgh_upload_fragments.argtypes = [R, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
        lib.lgh_upload_objects.argtypes = [R, C.c_void_p, C.c_uint32]
        lib.lgh_upload_light_depth.argtypes = [R, C.c_void_p, C.c_uint32]
        lib.lgh_upload_mesh.argtypes = [R, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        lib.lgh_use_mesh.argtypes = [R, C.c_uint32]
        lib.lgh_set_debug_overlay.argtypes = [R, C.c_uint32]
        lib.lgh_set_external_swapchain.argtypes = [R, C.c_void_p]
        lib.lgh_render_frame.argtypes = [R, C.c_uint32, C.c_int32, C.c_uint32, abi.ROWS, C.c_uint32]
        lib.lgh_render_stages.argtypes = [R, C.c_uint32, C.c_int32, C.c_uint32, abi.ROWS, C.c_uint32]
        H64 = C.c_ubyte * 64
        lib.lgh_ipc_export_image.argtypes = [R, C.c_char_p, H64, C.POINTER(C.c_uint64)]
        lib.lgh_ipc_export_ptr.argtypes = [C.c_void_p, H64]
        lib.lgh_ipc_open.argtypes = [H64, C.POINTER(C.c_void_p)]
        lib.lgh_ipc_close.argtypes = [C.c_void_p]
        lib.lgh_device_alloc_zeroed.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
        lib.lgh_device_free.argtypes = [C.c_void_p]
        lib.lgh_capture_frame.argtypes = [R, C.c_uint32, C.c_int32, C.c_uint32, abi.ROWS]
        lib.lgh_replay_frame.argtypes = [R]
        lib.lgh_captured_kernel_count.argtypes = [R]
        lib.lgh_last_pass_count.argtypes = [R]
        lib.lgh_image_desc.argtypes = [R, C.c_char_p, abi.IMG]
        lib.lgh_download_image.argtypes = [R, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
        lib.lgh_run_interleave.argtypes = [R, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.lgh_sync.argtypes = [R]
        lib.lgh_get_profile.argtypes = [R, C.c_char_p, C.c_uint64, C.POINTER(C.c_float), C.c_uint32]
        lib.lgh_allocated_bytes.argtypes = [R]
        lib.lgh_allocated_bytes.restype = C.c_uint64
        _lib = lib
    return _lib


def _check(status: int, what: str) -> None:
    if status < 0:
        raise RuntimeError(f"{what} failed ({status}): {(load_harness().lgh_last_error() or b'').decode()}")


class Renderer:
    def __init__(self, width: int, height: int, stream: int = 0):
        self.lib = load_harness()
        self.width, self.height = width, height
        self.handle = self.lib.lgh_create(width, height, C.c_void_p(stream))
        if not self.handle:
            raise RuntimeError("lgh_create failed: " + (self.lib.lgh_last_error() or b"").decode())

    def close(self):
        if self.handle:
            self.lib.lgh_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene ---------------------------------------------------------------------------------------------------
    def set_camera(self, camera: dict, light: dict) -> None:
        cp = (C.c_float * 3)(*camera["pos"])
        lp = (C.c_float * 3)(*light["pos"])
        _check(self.lib.lgh_set_camera(self.handle, cp, camera["vert"], camera["hor"], lp, light["vert"], light["hor"]), "lgh_set_camera")

    def upload_fragments(self, host_ptr: int, pitch: int, rows: Optional[Tuple[int, int]] = None) -> None:
        y0, y1 = rows if rows is not None else (0, self.height)
        _check(self.lib.lgh_upload_fragments(self.handle, C.c_void_p(host_ptr), pitch, y0, y1), "lgh_upload_fragments")

    def upload_objects(self, host_ptr: int, count: int) -> None:
        _check(self.lib.lgh_upload_objects(self.handle, C.c_void_p(host_ptr), count), "lgh_upload_objects")

    def upload_light_depth(self, host_ptr: int, size: int = 1024) -> None:
        _check(self.lib.lgh_upload_light_depth(self.handle, C.c_void_p(host_ptr), size), "lgh_upload_light_depth")

    def upload_scene(self, sc) -> None:
        """Convenience for tests: upload a legitengine_b200.scene.Scene from pageable numpy memory and wait."""
        self.upload_fragments(sc.fragments.ctypes.data, sc.fragments.strides[0])
        self.upload_objects(sc.objects.ctypes.data, len(sc.objects))
        self.upload_light_depth(np.ascontiguousarray(sc.shadow_map).ctypes.data, sc.shadow_map.shape[0])
        self.sync()

    def upload_mesh(self, mesh) -> None:
        """Scene in the reference's form (legitengine_b200.scene.Mesh, or anything with the same four arrays — e.g. views of pinned
        memory): frames then start from the mesh (ShadowPass + GBufferRasterPass rasterise it on the device). Asynchronous."""
        _check(self.lib.lgh_upload_mesh(self.handle, C.c_void_p(mesh.vertices.ctypes.data), len(mesh.vertices), C.c_void_p(mesh.indices.ctypes.data), len(mesh.indices),
                                        C.c_void_p(mesh.draws.ctypes.data), len(mesh.draws), C.c_void_p(mesh.objects.ctypes.data), len(mesh.objects)), "lgh_upload_mesh")

    def set_debug_overlay(self, enable: bool) -> None:
        """DebugInfoPass (thumbnails of normal / albedo / indirectLight / denoisedIndirectLight) over the finished frame."""
        _check(self.lib.lgh_set_debug_overlay(self.handle, 1 if enable else 0), "lgh_set_debug_overlay")

    def set_external_swapchain(self, device_base: int) -> None:
        """The frame's swapchain target becomes the image at `device_base` (canonical layout of a W x H BGRA8 image; may be a peer GPU's
        memory through a CUDA-IPC mapping). 0 restores the renderer's own image."""
        _check(self.lib.lgh_set_external_swapchain(self.handle, C.c_void_p(device_base or None)), "lgh_set_external_swapchain")

    def use_mesh(self, enable: bool) -> None:
        _check(self.lib.lgh_use_mesh(self.handle, 1 if enable else 0), "lgh_use_mesh")

    # -- frames --------------------------------------------------------------------------------------------------
    @staticmethod
    def _rows(rows):
        return None if rows is None else C.byref(abi.LgcuRows(rows[0], rows[1]))

    def render_frame(self, mode: int = MODE_FUSED, denoiser_radius: int = 0, gi_flags: int = abi.GI_DEFAULT, rows=None, profile: bool = False) -> None:
        _check(self.lib.lgh_render_frame(self.handle, mode, denoiser_radius, gi_flags, self._rows(rows), 1 if profile else 0), "lgh_render_frame")

    def render_stages(self, stages: int, rows=None, denoiser_radius: int = 0, gi_flags: int = abi.GI_DEFAULT) -> None:
        """Selected stages of the fused frame (STAGE_* bits) on rows [y0, y1): the building block of the strip-sharded renderer."""
        _check(self.lib.lgh_render_stages(self.handle, MODE_FUSED, denoiser_radius, gi_flags, self._rows(rows), stages), "lgh_render_stages")

    def capture_frame(self, mode: int = MODE_FUSED, denoiser_radius: int = 0, gi_flags: int = abi.GI_DEFAULT, rows=None) -> None:
        _check(self.lib.lgh_capture_frame(self.handle, mode, denoiser_radius, gi_flags, self._rows(rows)), "lgh_capture_frame")

    def replay_frame(self) -> None:
        _check(self.lib.lgh_replay_frame(self.handle), "lgh_replay_frame")

    def captured_kernel_count(self) -> int:
        return self.lib.lgh_captured_kernel_count(self.handle)

    def last_pass_count(self) -> int:
        return self.lib.lgh_last_pass_count(self.handle)

    def sync(self) -> None:
        _check(self.lib.lgh_sync(self.handle), "lgh_sync")

    def profile(self) -> List[Tuple[str, float]]:
        names = C.create_string_buffer(8192)
        ms = (C.c_float * 256)()
        n = self.lib.lgh_get_profile(self.handle, names, 8192, ms, 256)
        _check(n, "lgh_get_profile")
        labels = names.value.decode().split("\n")
        return [(labels[i], float(ms[i])) for i in range(n)]

    def allocated_bytes(self) -> int:
        return int(self.lib.lgh_allocated_bytes(self.handle))

    # -- images --------------------------------------------------------------------------------------------------
    def image_desc(self, name: str) -> abi.LgcuImage:
        d = abi.LgcuImage()
        _check(self.lib.lgh_image_desc(self.handle, name.encode(), C.byref(d)), "lgh_image_desc")
        return d

    def download_swapchain(self, host_ptr: int, pitch: int, rows=None) -> None:
        y0, y1 = rows if rows is not None else (0, self.height)
        _check(self.lib.lgh_download_image(self.handle, b"swapchain", 0, C.c_void_p(host_ptr), pitch, y0, y1), "lgh_download_image")

    def run_interleave(self, name: str, grid=(4, 4)):
        """InterleaveBuilder::Deinterleave then ::Interleave of image `name` through the rendergraph: (de-interleaved, round trip) as
        raw (h, w, texel bytes) uint8 arrays."""
        d = self.image_desc(name)
        ts = abi.TEXEL_SIZE[d.format]
        a = np.zeros((d.height, d.width, ts), dtype=np.uint8)
        b = np.zeros_like(a)
        _check(self.lib.lgh_run_interleave(self.handle, name.encode(), grid[0], grid[1], C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), d.width * ts), "lgh_run_interleave")
        return a, b

    def download_image(self, name: str) -> images.HostImage:
        """Whole image (all levels) into a HostImage with the canonical layout; synchronises."""
        d = self.image_desc(name)
        host = images.HostImage(d.format, d.width, d.height, d.imageMipCount)
        for l in range(d.imageMipCount):
            w, h = host.level_size(l)
            if w <= 0 or h <= 0:
                continue
            ptr = host.buf.ctypes.data + host.desc.levelOffset[l]
            _check(self.lib.lgh_download_image(self.handle, name.encode(), l, C.c_void_p(ptr), host.desc.levelPitch[l], 0, h), "lgh_download_image")
        self.sync()
        return host


# ---- CUDA IPC helpers (peer-to-peer strip exchange) -----------------------------------------------------------------------
def ipc_export_image(renderer: "Renderer", name: str) -> bytes:
    h = (C.c_ubyte * 64)()
    _check(renderer.lib.lgh_ipc_export_image(renderer.handle, name.encode(), h, None), "lgh_ipc_export_image")
    return bytes(h)


def ipc_export_ptr(ptr: int) -> bytes:
    h = (C.c_ubyte * 64)()
    _check(load_harness().lgh_ipc_export_ptr(C.c_void_p(ptr), h), "lgh_ipc_export_ptr")
    return bytes(h)


def ipc_open(handle: bytes) -> int:
    out = C.c_void_p()
    _check(load_harness().lgh_ipc_open((C.c_ubyte * 64).from_buffer_copy(handle), C.byref(out)), "lgh_ipc_open")
    return int(out.value)


def ipc_close(ptr: int) -> None:
    load_harness().lgh_ipc_close(C.c_void_p(ptr))


def device_alloc_zeroed(nbytes: int) -> int:
    out = C.c_void_p()
    _check(load_harness().lgh_device_alloc_zeroed(nbytes, C.byref(out)), "lgh_device_alloc_zeroed")
    return int(out.value)


def device_free(ptr: int) -> None:
    load_harness().lgh_device_free(C.c_void_p(ptr))
