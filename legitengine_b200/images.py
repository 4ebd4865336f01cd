"""Image containers in the library's canonical linear layout (include/lgcu.h, `lgcu_image`).

`HostImage` lives in numpy memory (oracle side, fixtures), `DeviceImage` in a torch CUDA byte tensor (PyTorch is only
the allocator here). Both expose `desc` (an `LgcuImage`) and `view(base_mip, mip_count)` — the analogue of
`RenderGraph::AddImageView(image, baseMip, mipCount, 0, 1)` (LV/RenderGraph.h:313-325).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import abi

PITCH_ALIGN = 128
LEVEL_ALIGN = 256


def _align(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def mip_size(width: int, height: int, level: int) -> Tuple[int, int]:
    """Size of mip `level` (integer floor halving; LV/RenderGraph.h:373-374, LV/Image.h:90-100)."""
    return width >> level, height >> level


def make_layout(fmt: int, width: int, height: int, mips: int) -> Tuple[abi.LgcuImage, int]:
    """Python twin of lgcu_image_layout(): returns (descriptor without base, allocation bytes)."""
    texel = abi.TEXEL_SIZE[fmt]
    d = abi.LgcuImage()
    d.base = None
    d.format = fmt
    d.width, d.height = width, height
    d.imageMipCount = mips
    d.baseMip, d.mipCount = 0, mips
    off = 0
    for l in range(mips):
        w, h = mip_size(width, height, l)
        pitch = _align(max(w, 1) * texel, PITCH_ALIGN)
        d.levelOffset[l] = off
        d.levelPitch[l] = pitch
        off += _align(pitch * max(h, 1), LEVEL_ALIGN)
    return d, off


def _copy_desc(d: abi.LgcuImage) -> abi.LgcuImage:
    out = abi.LgcuImage()
    C.memmove(C.byref(out), C.byref(d), C.sizeof(abi.LgcuImage))
    return out


class _ImageBase:
    desc: abi.LgcuImage
    nbytes: int

    @property
    def format(self) -> int:
        return self.desc.format

    @property
    def mips(self) -> int:
        return self.desc.imageMipCount

    def level_size(self, level: int) -> Tuple[int, int]:
        return mip_size(self.desc.width, self.desc.height, level)

    def view(self, base_mip: int = 0, mip_count: int | None = None) -> abi.LgcuImage:
        v = _copy_desc(self.desc)
        v.baseMip = base_mip
        v.mipCount = self.desc.imageMipCount - base_mip if mip_count is None else mip_count
        return v


class HostImage(_ImageBase):
    def __init__(self, fmt: int, width: int, height: int, mips: int = 1, fill: int = 0xCD):
        self.desc, self.nbytes = make_layout(fmt, width, height, mips)
        # poison fill so that texels a pass forgets to write show up in comparisons
        self.buf = np.full(self.nbytes, fill, dtype=np.uint8)
        self.desc.base = self.buf.ctypes.data

    # raw texel bytes of one level as (h, w, texel_size) uint8 view
    def level_bytes(self, level: int) -> np.ndarray:
        w, h = self.level_size(level)
        texel = abi.TEXEL_SIZE[self.format]
        off, pitch = self.desc.levelOffset[level], self.desc.levelPitch[level]
        rows = np.lib.stride_tricks.as_strided(self.buf[off:], shape=(h, w, texel), strides=(pitch, texel, 1), writeable=True)
        return rows

    def level_raw(self, level: int) -> np.ndarray:
        """Level as its storage type: float16 (h,w,4), float32 (h,w,2|4|1) or uint8 (h,w,4)."""
        b = self.level_bytes(level)
        w, h = self.level_size(level)
        fmt = self.format
        if fmt == abi.FORMAT_R16G16B16A16_SFLOAT:
            return np.ascontiguousarray(b).view(np.float16).reshape(h, w, 4)
        if fmt == abi.FORMAT_R32G32_SFLOAT:
            return np.ascontiguousarray(b).view(np.float32).reshape(h, w, 2)
        if fmt == abi.FORMAT_R32G32B32A32_SFLOAT:
            return np.ascontiguousarray(b).view(np.float32).reshape(h, w, 4)
        if fmt == abi.FORMAT_D32_SFLOAT:
            return np.ascontiguousarray(b).view(np.float32).reshape(h, w, 1)
        if fmt == abi.FORMAT_B8G8R8A8_SRGB:
            return np.ascontiguousarray(b).reshape(h, w, 4)
        raise ValueError(fmt)

    def level_f32(self, level: int) -> np.ndarray:
        return self.level_raw(level).astype(np.float32)

    def set_level(self, level: int, values: np.ndarray) -> None:
        """Store `values` (h,w,channels) converted to the storage type (numpy float16 cast is RTNE)."""
        fmt = self.format
        w, h = self.level_size(level)
        if fmt == abi.FORMAT_R16G16B16A16_SFLOAT:
            raw = np.ascontiguousarray(np.asarray(values, dtype=np.float32).astype(np.float16).reshape(h, w, 4)).view(np.uint8)
        elif fmt in (abi.FORMAT_R32G32_SFLOAT, abi.FORMAT_R32G32B32A32_SFLOAT, abi.FORMAT_D32_SFLOAT):
            ch = {abi.FORMAT_R32G32_SFLOAT: 2, abi.FORMAT_R32G32B32A32_SFLOAT: 4, abi.FORMAT_D32_SFLOAT: 1}[fmt]
            raw = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(h, w, ch)).view(np.uint8)
        else:
            raw = np.ascontiguousarray(np.asarray(values, dtype=np.uint8).reshape(h, w, 4))
        self.level_bytes(level)[...] = raw.reshape(h, w, -1)

    def levels_equal(self, other: "HostImage", level: int) -> bool:
        return bool(np.array_equal(self.level_bytes(level), other.level_bytes(level)))


class DeviceImage(_ImageBase):
    """Device-resident image backed by a torch uint8 CUDA tensor (torch = allocator / stream plumbing only)."""

    def __init__(self, fmt: int, width: int, height: int, mips: int = 1, device="cuda:0", fill: int | None = 0xCD):
        import torch

        self.desc, self.nbytes = make_layout(fmt, width, height, mips)
        self.tensor = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        if fill is not None:
            self.tensor.fill_(fill)
        self.desc.base = self.tensor.data_ptr()

    @classmethod
    def from_host(cls, host: HostImage, device="cuda:0") -> "DeviceImage":
        import torch

        d = cls(host.format, host.desc.width, host.desc.height, host.mips, device=device, fill=None)
        d.tensor.copy_(torch.from_numpy(host.buf))
        return d

    def to_host(self) -> HostImage:
        h = HostImage(self.format, self.desc.width, self.desc.height, self.mips)
        h.buf[...] = self.tensor.cpu().numpy()
        return h


def image_bytes(images: List[_ImageBase]) -> int:
    return sum(i.nbytes for i in images)
