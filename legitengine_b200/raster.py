"""Device-side scene for the rasterisation front end (include/lgcu.h: lgcu_raster_*).

`DeviceMesh` keeps what the reference's Scene keeps on the GPU — vertex / index buffers, the per-object constants and the draw
list (src/Scene/Scene.h:141-147, src/Scene/Mesh.h:244-262) — plus the scratch the raster kernels need, and calls the C ABI.
PyTorch only owns the device memory; every computation happens in csrc/k_raster.cu.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi, scene


class DeviceMesh:
    def __init__(self, mesh: scene.Mesh, device="cuda:0"):
        import torch

        self.lib = abi.load_lgcu()
        self.device = torch.device(device)
        draws = mesh.draws.copy()
        self.triangles = int(self.lib.lgcu_raster_prepare_draws(draws.ctypes.data, len(draws)))
        up = lambda a: torch.from_numpy(np.frombuffer(np.ascontiguousarray(a).tobytes(), dtype=np.uint8).copy()).to(self.device)
        self.vertices, self.indices, self.draws, self.objects = up(mesh.vertices), up(mesh.indices), up(draws), up(mesh.objects)
        self.desc = abi.MeshScene(self.vertices.data_ptr(), self.indices.data_ptr(), self.draws.data_ptr(), self.objects.data_ptr(),
                                  len(mesh.vertices), len(mesh.indices), len(draws), len(mesh.objects), self.triangles)
        self._scratch = None

    def scratch(self, width: int, height: int):
        import torch

        need = int(self.lib.lgcu_raster_scratch_bytes(self.triangles, width, height))
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = self._scratch.data_ptr()
        return (base + 255) // 256 * 256, need

    def raster_gbuffer(self, view, proj, width: int, height: int, fragments, rows=None, stream: int = 0) -> None:
        """fragments: uint8 device tensor of shape (height, width * 32) receiving lgcu_fragment records."""
        g = abi.GBufferBuilderData(abi.mat4(view), abi.mat4(proj), 0.0, 0.0)
        ptr, nbytes = self.scratch(width, height)
        r = C.byref(abi.LgcuRows(rows[0], rows[1])) if rows is not None else None
        abi.check(self.lib.lgcu_raster_gbuffer(C.byref(g), C.byref(self.desc), C.c_void_p(ptr), nbytes, width, height, C.c_void_p(fragments.data_ptr()),
                                               fragments.stride(0), r, C.c_void_p(stream)), "lgcu_raster_gbuffer")

    def raster_shadow_map(self, light_view, light_proj, shadow_map, stream: int = 0) -> None:
        """shadow_map: images.DeviceImage (D32F)."""
        p = abi.ShadowmapBuilderData(abi.mat4(light_view), abi.mat4(light_proj))
        w, h = shadow_map.level_size(0)
        ptr, nbytes = self.scratch(w, h)
        view = shadow_map.view()
        abi.check(self.lib.lgcu_raster_shadow_map(C.byref(p), C.byref(self.desc), C.c_void_p(ptr), nbytes, C.byref(view), C.c_void_p(stream)), "lgcu_raster_shadow_map")


def host_mesh_desc(mesh: scene.Mesh) -> abi.MeshScene:
    """lgcu_mesh_scene over HOST arrays (what the CPU oracle takes). The arrays must outlive the descriptor."""
    return abi.MeshScene(mesh.vertices.ctypes.data, mesh.indices.ctypes.data, mesh.draws.ctypes.data, mesh.objects.ctypes.data,
                         len(mesh.vertices), len(mesh.indices), len(mesh.draws), len(mesh.objects), mesh.triangle_count)
