"""CPU tests of the Vulkan hand-back shim (include/lgcu_interop.h, include/lgcu_vulkan.h): the library exports
every declared entry point, the layout conversion (pure host code) accepts what vkGetImageSubresourceLayout reports for the images the
passes touch and rejects what the kernels could not address, the import calls fail cleanly without a device, and the Vulkan-typed
header compiles and runs against tests/stubs/vulkan/vulkan_core.h (the image has no Vulkan SDK; the header only converts types)."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import pytest

from legitengine_b200 import abi, images
from tests.test_abi_cpu import _declared

ROOT = Path(__file__).resolve().parent.parent


def _lib():
    lib = abi.load_lgcu()
    lib.lgcu_image_from_linear_layout.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(abi.LgcuImage)]
    lib.lgcu_import_memory_fd.argtypes = [C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.lgcu_import_timeline_semaphore_fd.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.lgcu_semaphore_wait.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    return lib


def test_interop_entry_points_are_exported():
    lib = abi.load_lgcu()
    names = _declared("lgcu_interop.h", "lgcu")
    assert len(names) == 7, names
    for n in names:
        assert hasattr(lib, n), f"liblgcu.so does not export {n}"


def test_linear_layout_to_image():
    lib = _lib()
    W, Hh, mips = 1920, 1080, 10
    want, nbytes = images.make_layout(abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, mips)  # a layout the kernels are known to take
    offs = (C.c_uint64 * mips)(*[want.levelOffset[l] for l in range(mips)])
    pitch = (C.c_uint64 * mips)(*[want.levelPitch[l] for l in range(mips)])
    img = abi.LgcuImage()
    base = 0x7F0000000000
    assert lib.lgcu_image_from_linear_layout(C.c_void_p(base), abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, mips, offs, pitch, C.byref(img)) == 0
    assert img.base == base and (img.width, img.height, img.imageMipCount, img.baseMip, img.mipCount) == (W, Hh, mips, 0, mips)
    assert [img.levelOffset[l] for l in range(mips)] == [want.levelOffset[l] for l in range(mips)]
    assert [img.levelPitch[l] for l in range(mips)] == [want.levelPitch[l] for l in range(mips)]
    # a driver's own row pitch (wider than the tight one) is fine as long as it is a multiple of 16 bytes
    pitch[0] = 1920 * 8 + 256
    offs2 = (C.c_uint64 * mips)(*[want.levelOffset[l] + (0 if l == 0 else 256 * 1080) for l in range(mips)])
    assert lib.lgcu_image_from_linear_layout(C.c_void_p(base), abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, mips, offs2, pitch, C.byref(img)) == 0
    assert img.levelPitch[0] == 1920 * 8 + 256
    # rejected: row pitch shorter than a row, unaligned pitch, overlapping levels, unknown format, unaligned base
    bad = (C.c_uint64 * mips)(*pitch)
    bad[0] = 1920 * 8 - 16
    assert lib.lgcu_image_from_linear_layout(C.c_void_p(base), abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, mips, offs2, bad, C.byref(img)) == abi.LGCU_ERR_INVALID_ARGUMENT
    bad[0] = 1920 * 8 + 8
    assert lib.lgcu_image_from_linear_layout(C.c_void_p(base), abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, mips, offs2, bad, C.byref(img)) == abi.LGCU_ERR_INVALID_ARGUMENT
    overlap = (C.c_uint64 * mips)(*offs)
    overlap[1] = overlap[0] + 16
    pitch[0] = want.levelPitch[0]
    assert lib.lgcu_image_from_linear_layout(C.c_void_p(base), abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, mips, overlap, pitch, C.byref(img)) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_image_from_linear_layout(C.c_void_p(base), 37, W, Hh, mips, offs, pitch, C.byref(img)) == abi.LGCU_ERR_UNSUPPORTED_FORMAT
    assert lib.lgcu_image_from_linear_layout(C.c_void_p(base + 8), abi.FORMAT_R16G16B16A16_SFLOAT, W, Hh, mips, offs, pitch, C.byref(img)) == abi.LGCU_ERR_INVALID_ARGUMENT


def test_imports_fail_cleanly_without_a_device_or_a_valid_fd():
    lib = _lib()
    mem, ptr, sem = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.lgcu_import_memory_fd(-1, 4096, 0, C.byref(mem), C.byref(ptr)) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_import_timeline_semaphore_fd(-1, C.byref(sem)) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_semaphore_wait(None, 1, None) == abi.LGCU_ERR_INVALID_ARGUMENT
    assert lib.lgcu_release_memory(None) == 0 and lib.lgcu_release_semaphore(None) == 0


RUNNER = r"""
#define LGCU_WITH_VULKAN 1
#include "lgcu_vulkan.h"
#include <stdio.h>
static int record(void *user, void *stream) { (void)stream; *(int *)user += 1; return LGCU_OK; }
int main(void) {
  VkSubresourceLayout layouts[3] = {{0, 0, 2048, 0, 0}, {2048 * 100, 0, 1024, 0, 0}, {2048 * 100 + 1024 * 50, 0, 512, 0, 0}};
  VkExtent3D extent = {200, 100, 1};
  static char memory[1 << 20] __attribute__((aligned(256)));
  lgcu_image img;
  int st = lgcu_vk_image(memory, 256, VK_FORMAT_R32G32_SFLOAT, extent, 3, layouts, &img);
  if (st != LGCU_OK) { printf("lgcu_vk_image failed %d: %s\n", st, lgcu_last_error()); return 1; }
  if (img.base != memory + 256 || img.format != LGCU_FORMAT_R32G32_SFLOAT || img.levelPitch[1] != 1024 || img.levelOffset[2] != 2048 * 100 + 1024 * 50) return 2;
  if (lgcu_vk_image(memory, 256, (VkFormat)37, extent, 3, layouts, &img) != LGCU_ERR_UNSUPPORTED_FORMAT) return 3;
  if (lgcu_vk_format(VK_FORMAT_B8G8R8A8_SRGB) != LGCU_FORMAT_B8G8R8A8_SRGB || lgcu_vk_format(VK_FORMAT_UNDEFINED) != LGCU_FORMAT_UNDEFINED) return 4;
  int calls = 0; /* a null timeline is refused before anything is recorded */
  if (lgcu_vk_cuda_section(NULL, 7, NULL, record, &calls) != LGCU_ERR_INVALID_ARGUMENT || calls != 0) return 5;
  printf("ok\n");
  return 0;
}
"""


@pytest.mark.parametrize("compiler,std", [("gcc", "-std=c11"), ("g++", "-std=c++17")])
def test_vulkan_typed_header_compiles_and_runs_against_the_stub_sdk(tmp_path, compiler, std):
    if not shutil.which(compiler):
        pytest.skip(f"{compiler} not installed")
    abi.load_lgcu()
    src = tmp_path / ("runner.c" if compiler == "gcc" else "runner.cpp")
    src.write_text(RUNNER)
    exe = tmp_path / "runner"
    libdir = ROOT / "legitengine_b200" / "lib"
    cmd = [compiler, std, "-Wall", "-Werror", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tests' / 'stubs'}", str(src), "-o", str(exe), f"-L{libdir}", "-llgcu", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip() == "ok", (run.returncode, run.stdout, run.stderr)
    # without the guard the header is empty: including it never needs a Vulkan SDK
    plain = tmp_path / "plain.c"
    plain.write_text('#include "lgcu_vulkan.h"\nint main(void) { return 0; }\n')
    assert subprocess.run(["gcc", "-fsyntax-only", f"-I{ROOT / 'include'}", str(plain)], capture_output=True).returncode == 0
