"""The SSVGI frame as a list of pass calls, written once for any backend with the include/lgcu.h pass signatures.

`FrameImages` declares the eleven screen images exactly as SSVGIRenderer::ViewportResources does
(src/Render/Renderers/SSVGIRenderer.h:393-404: formats, 10-level MippedProxy chains, 1024² shadow map) plus the
swapchain target (LV/Swapchain.h:108). `run_pass_list` issues the passes in the order of
SSVGIRenderer::RenderFrame (:107-342): GBuffer, Light, 2x9 MipBuilder, 2x10 Blur, IndirectLight, Denoiser, Gathering.

A backend is any object with callables gbuffer_resolve, direct_light, mip_level, blur_level, gi_gather, denoise,
final_gather taking (params..., images..., rows): the CPU oracles (oracle.loader) or `CudaPasses` below, which appends
the stream and raises on a non-zero status. This module contains no arithmetic.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import abi, images

MIPS = 10  # MipBuilder.h:21


@dataclass
class FrameImages:
    width: int
    height: int
    make: type  # images.HostImage or images.DeviceImage
    kwargs: dict = field(default_factory=dict)
    indirect_format: int = abi.FORMAT_R16G16B16A16_SFLOAT
    shadow_size: int = 1024

    def __post_init__(self):
        W, H, mk, kw = self.width, self.height, self.make, self.kwargs
        F16, RG32, D32 = abi.FORMAT_R16G16B16A16_SFLOAT, abi.FORMAT_R32G32_SFLOAT, abi.FORMAT_D32_SFLOAT
        self.albedo = mk(F16, W, H, 1, **kw)
        self.emissive = mk(F16, W, H, 1, **kw)
        self.normal = mk(F16, W, H, 1, **kw)
        self.depthMoments = mk(RG32, W, H, MIPS, **kw)
        self.blurredDepthMoments = mk(RG32, W, H, MIPS, **kw)
        self.depthStencil = mk(D32, W, H, 1, **kw)
        self.directLight = mk(F16, W, H, MIPS, **kw)
        self.blurredDirectLight = mk(F16, W, H, MIPS, **kw)
        self.shadowMap = mk(D32, self.shadow_size, self.shadow_size, 1, **kw)
        self.indirectLight = mk(self.indirect_format, W, H, 1, **kw)
        self.denoisedIndirectLight = mk(self.indirect_format, W, H, 1, **kw)
        self.swapchain = mk(abi.FORMAT_B8G8R8A8_SRGB, W, H, 1, **kw)

    NAMES = (
        "albedo", "emissive", "normal", "depthMoments", "blurredDepthMoments", "depthStencil", "directLight",
        "blurredDirectLight", "shadowMap", "indirectLight", "denoisedIndirectLight", "swapchain",
    )

    def items(self):
        return [(n, getattr(self, n)) for n in self.NAMES]

    def nbytes(self) -> int:
        return sum(img.nbytes for _, img in self.items())


def mip_levels_built(width: int, height: int, mips: int = MIPS) -> int:
    """Number of levels MipBuilder::BuildMips actually produces (loop stops when a dimension hits 0, MipBuilder.h:151-152)."""
    n = 1
    for l in range(1, mips):
        if (width >> l) <= 0 or (height >> l) <= 0:
            break
        n += 1
    return n


@dataclass
class FrameParams:
    gbuffer: abi.GBufferBuilderData
    light: abi.DirectLightingData
    mip: abi.MipLevelBuilderData
    indirect: abi.IndirectLightingData
    denoiser: abi.DenoiserData
    final: abi.FinalGathererData
    clear: abi.ClearValues


def make_params(width: int, height: int, m, denoise_radius: int = 0) -> FrameParams:
    """UBO contents as the record lambdas of SSVGIRenderer::RenderFrame fill them (:127-131, :179-185, :240-244, :283-288, :321-324)."""
    view, proj = abi.mat4(m.view), abi.mat4(m.proj)
    lview, lproj = abi.mat4(m.light_view), abi.mat4(m.light_proj)
    ext = (C.c_float * 4)(float(width), float(height), 0.0, 0.0)
    return FrameParams(
        gbuffer=abi.GBufferBuilderData(view, proj, 0.0, 0.0),
        light=abi.DirectLightingData(view, proj, lview, lproj, 0.0),
        mip=abi.MipLevelBuilderData(0.0),
        indirect=abi.IndirectLightingData(view, proj, ext),
        denoiser=abi.DenoiserData(view, proj, ext, denoise_radius),
        final=abi.FinalGathererData(view, proj),
        clear=abi.default_clear(),
    )


class CudaPasses:
    """Adapter giving liblgcu.so the backend interface: appends the stream, checks the status."""

    kind = "cuda"

    def __init__(self, stream: int = 0):
        self.lib = abi.load_lgcu()
        self.stream = stream

    def __getattr__(self, name):
        fn = getattr(self.lib, "lgcu_" + name)

        def call(*args):
            status = fn(*args, C.c_void_p(self.stream))
            abi.check(status, "lgcu_" + name)
            return status

        return call


def _rows(rows):
    return None if rows is None else C.byref(abi.LgcuRows(rows[0], rows[1]))


def upload_inputs(fi: FrameImages, sc) -> Dict[str, object]:
    """Place the scene's fragment buffer, object table and shadow map where the backend can read them.
    Host backends get numpy pointers, device backends get torch CUDA byte tensors."""
    out: Dict[str, object] = {}
    if fi.make is images.HostImage:
        out["fragments"], out["objects"] = sc.fragments, sc.objects
        out["fragments_ptr"], out["objects_ptr"] = sc.fragments.ctypes.data, sc.objects.ctypes.data
        fi.shadowMap.set_level(0, sc.shadow_map[..., None])
    else:
        import torch

        dev = fi.kwargs.get("device", "cuda:0")
        ft = torch.from_numpy(sc.fragments.view(np.uint8).reshape(-1)).to(dev)
        ot = torch.from_numpy(sc.objects.view(np.uint8).reshape(-1)).to(dev)
        out["fragments"], out["objects"] = ft, ot
        out["fragments_ptr"], out["objects_ptr"] = ft.data_ptr(), ot.data_ptr()
        host_shadow = images.HostImage(abi.FORMAT_D32_SFLOAT, sc.shadow_map.shape[1], sc.shadow_map.shape[0], 1)
        host_shadow.set_level(0, sc.shadow_map[..., None])
        fi.shadowMap.tensor.copy_(torch.from_numpy(host_shadow.buf))
    out["pitch"] = sc.fragments.strides[0]
    out["n_objects"] = len(sc.objects)
    return out


def run_pass_list(be, fi: FrameImages, p: FrameParams, inputs: Dict[str, object], rows: Optional[tuple] = None,
                  gi_flags: int = abi.GI_DEFAULT, stop_after: Optional[str] = None) -> None:
    """SSVGIRenderer::RenderFrame, pass by pass (unfused)."""
    W, H = fi.width, fi.height
    r = _rows(rows)
    v = lambda img, base=0, n=None: C.byref(img.view(base, n))
    be.gbuffer_resolve(C.byref(p.gbuffer), inputs["objects_ptr"], inputs["n_objects"], inputs["fragments_ptr"], inputs["pitch"],
                       C.byref(p.clear), v(fi.albedo), v(fi.emissive), v(fi.normal), v(fi.depthMoments, 0, 1), v(fi.depthStencil), r)
    if stop_after == "gbuffer":
        return
    be.direct_light(C.byref(p.light), v(fi.albedo), v(fi.emissive), v(fi.normal), v(fi.depthStencil), v(fi.shadowMap),
                    v(fi.directLight, 0, 1), r)
    if stop_after == "light":
        return
    levels = mip_levels_built(W, H)
    for chain in (fi.directLight, fi.depthMoments):  # SSVGIRenderer.h:207-208
        for l in range(1, levels):
            be.mip_level(C.byref(p.mip), v(chain, l - 1, 1), v(chain, l, 1), r)
    if stop_after == "mips":
        return
    for src, dst in ((fi.directLight, fi.blurredDirectLight), (fi.depthMoments, fi.blurredDepthMoments)):  # :209-221
        for l in range(MIPS):
            w, h = images.mip_size(W, H, l)
            if w <= 0 or h <= 0:
                continue  # the reference would create a zero-sized render area here; only reachable below 512 px
            bp = abi.BlurLayerBuilderData((C.c_int32 * 4)(w, h, 0, 0), 0 if l == 0 else 2)
            be.blur_level(C.byref(bp), v(src, l, 1), v(dst, l, 1), r)
    if stop_after == "blur":
        return
    be.gi_gather(C.byref(p.indirect), v(fi.blurredDirectLight), v(fi.blurredDepthMoments), v(fi.normal), v(fi.depthStencil),
                 v(fi.indirectLight), gi_flags, r)
    if stop_after == "gather":
        return
    be.denoise(C.byref(p.denoiser), v(fi.indirectLight), v(fi.normal), v(fi.depthMoments), v(fi.denoisedIndirectLight), r)
    be.final_gather(C.byref(p.final), v(fi.directLight), v(fi.blurredDirectLight), v(fi.albedo), v(fi.denoisedIndirectLight),
                    v(fi.swapchain), r)


def debug_tiles(count: int):
    """QuadData of the first `count` tiles, in the fp32 arithmetic of DebugRenderer.h:27-57."""
    f = np.float32
    size, pad = f(0.1), f(0.02)
    x, y = pad, pad
    out = []
    for _ in range(count):
        out.append(abi.DebugQuadData((C.c_float * 4)(float(x), float(y), float(f(x + size)), float(f(y + size)))))
        x = f(x + f(size + pad))
        if f(x + size) > f(1.0):
            x = pad
            y = f(y + f(size + pad))
    return out
