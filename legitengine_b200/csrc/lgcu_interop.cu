// lgcu_interop.cu — external memory / timeline-semaphore interop (include/lgcu_interop.h): thin, checked wrappers over the CUDA
// runtime's external-resource API. Host code only; compiled into liblgcu.so so that the engine links one library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/lgcu_interop.h"

extern "C" int lgcu_set_last_error(int status, const char *fmt, ...); // lgcu_api.cu

struct lgcu_external_memory {
  cudaExternalMemory_t handle;
  void *devicePtr;
  uint64_t size;
};
struct lgcu_external_semaphore {
  cudaExternalSemaphore_t handle;
};

namespace {
uint32_t texelBytes(uint32_t format) {
  switch (format) {
  case LGCU_FORMAT_B8G8R8A8_SRGB: return 4;
  case LGCU_FORMAT_R16G16B16A16_SFLOAT: return 8;
  case LGCU_FORMAT_R32G32_SFLOAT: return 8;
  case LGCU_FORMAT_R32G32B32A32_SFLOAT: return 16;
  case LGCU_FORMAT_D32_SFLOAT: return 4;
  default: return 0;
  }
}
} // namespace

extern "C" {

int lgcu_import_memory_fd(int fd, uint64_t allocationSize, int dedicated, lgcu_external_memory **memory, void **devicePtr) {
  if (fd < 0 || allocationSize == 0 || !memory || !devicePtr) return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "import_memory_fd: bad fd / size / null output");
  cudaExternalMemoryHandleDesc desc;
  std::memset(&desc, 0, sizeof(desc));
  desc.type = cudaExternalMemoryHandleTypeOpaqueFd;
  desc.handle.fd = fd;
  desc.size = allocationSize;
  desc.flags = dedicated ? cudaExternalMemoryDedicated : 0;
  cudaExternalMemory_t handle;
  cudaError_t e = cudaImportExternalMemory(&handle, &desc);
  if (e != cudaSuccess) return lgcu_set_last_error(LGCU_ERR_CUDA, "cudaImportExternalMemory: %s", cudaGetErrorString(e));
  cudaExternalMemoryBufferDesc buf;
  std::memset(&buf, 0, sizeof(buf));
  buf.offset = 0;
  buf.size = allocationSize;
  void *ptr = nullptr;
  e = cudaExternalMemoryGetMappedBuffer(&ptr, handle, &buf);
  if (e != cudaSuccess) {
    cudaDestroyExternalMemory(handle);
    return lgcu_set_last_error(LGCU_ERR_CUDA, "cudaExternalMemoryGetMappedBuffer: %s", cudaGetErrorString(e));
  }
  lgcu_external_memory *m = new (std::nothrow) lgcu_external_memory{handle, ptr, allocationSize};
  if (!m) {
    cudaDestroyExternalMemory(handle);
    return lgcu_set_last_error(LGCU_ERR_CUDA, "import_memory_fd: out of host memory");
  }
  *memory = m;
  *devicePtr = ptr;
  return LGCU_OK;
}

int lgcu_release_memory(lgcu_external_memory *memory) {
  if (!memory) return LGCU_OK;
  cudaFree(memory->devicePtr); // a mapped buffer is released with cudaFree before the memory object is destroyed
  const cudaError_t e = cudaDestroyExternalMemory(memory->handle);
  delete memory;
  return e == cudaSuccess ? LGCU_OK : lgcu_set_last_error(LGCU_ERR_CUDA, "cudaDestroyExternalMemory: %s", cudaGetErrorString(e));
}

int lgcu_image_from_linear_layout(void *base, uint32_t format, uint32_t width, uint32_t height, uint32_t mips, const uint64_t *levelOffsets,
                                  const uint64_t *rowPitches, lgcu_image *image) {
  const uint32_t texel = texelBytes(format);
  if (!texel) return lgcu_set_last_error(LGCU_ERR_UNSUPPORTED_FORMAT, "image_from_linear_layout: format %u", format);
  if (!base || !image || !levelOffsets || !rowPitches || width == 0 || height == 0 || mips == 0 || mips > LGCU_MAX_MIPS)
    return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "image_from_linear_layout: null pointer, empty image or %u levels", mips);
  if (reinterpret_cast<uintptr_t>(base) % 16) return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "image_from_linear_layout: base not 16-byte aligned");
  std::memset(image, 0, sizeof(*image));
  image->base = base;
  image->format = format;
  image->width = width;
  image->height = height;
  image->imageMipCount = mips;
  image->baseMip = 0;
  image->mipCount = mips;
  for (uint32_t l = 0; l < mips; l++) {
    const uint64_t w = (width >> l) ? (width >> l) : 1, h = (height >> l) ? (height >> l) : 1;
    if (rowPitches[l] < w * texel || rowPitches[l] > 0xffffffffull || (rowPitches[l] % 16) != 0 || (levelOffsets[l] % 16) != 0)
      return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "image_from_linear_layout: level %u: offset %llu / pitch %llu (need pitch >= %llu, both multiples of 16)", l,
                                 (unsigned long long)levelOffsets[l], (unsigned long long)rowPitches[l], (unsigned long long)(w * texel));
    for (uint32_t k = 0; k < l; k++) { // levels must not overlap
      const uint64_t hk = (height >> k) ? (height >> k) : 1;
      const uint64_t a0 = levelOffsets[k], a1 = a0 + rowPitches[k] * hk, b0 = levelOffsets[l], b1 = b0 + rowPitches[l] * h;
      if (a0 < b1 && b0 < a1) return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "image_from_linear_layout: levels %u and %u overlap", k, l);
    }
    image->levelOffset[l] = levelOffsets[l];
    image->levelPitch[l] = (uint32_t)rowPitches[l];
  }
  return LGCU_OK;
}

int lgcu_import_timeline_semaphore_fd(int fd, lgcu_external_semaphore **semaphore) {
  if (fd < 0 || !semaphore) return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "import_timeline_semaphore_fd: bad fd / null output");
  cudaExternalSemaphoreHandleDesc desc;
  std::memset(&desc, 0, sizeof(desc));
  desc.type = cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd;
  desc.handle.fd = fd;
  cudaExternalSemaphore_t handle;
  const cudaError_t e = cudaImportExternalSemaphore(&handle, &desc);
  if (e != cudaSuccess) return lgcu_set_last_error(LGCU_ERR_CUDA, "cudaImportExternalSemaphore: %s", cudaGetErrorString(e));
  lgcu_external_semaphore *s = new (std::nothrow) lgcu_external_semaphore{handle};
  if (!s) {
    cudaDestroyExternalSemaphore(handle);
    return lgcu_set_last_error(LGCU_ERR_CUDA, "import_timeline_semaphore_fd: out of host memory");
  }
  *semaphore = s;
  return LGCU_OK;
}

int lgcu_release_semaphore(lgcu_external_semaphore *semaphore) {
  if (!semaphore) return LGCU_OK;
  const cudaError_t e = cudaDestroyExternalSemaphore(semaphore->handle);
  delete semaphore;
  return e == cudaSuccess ? LGCU_OK : lgcu_set_last_error(LGCU_ERR_CUDA, "cudaDestroyExternalSemaphore: %s", cudaGetErrorString(e));
}

int lgcu_semaphore_wait(lgcu_external_semaphore *semaphore, uint64_t value, void *stream) {
  if (!semaphore) return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "semaphore_wait: null semaphore");
  cudaExternalSemaphoreWaitParams p;
  std::memset(&p, 0, sizeof(p));
  p.params.fence.value = value;
  const cudaError_t e = cudaWaitExternalSemaphoresAsync(&semaphore->handle, &p, 1, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? LGCU_OK : lgcu_set_last_error(LGCU_ERR_CUDA, "cudaWaitExternalSemaphoresAsync: %s", cudaGetErrorString(e));
}

int lgcu_semaphore_signal(lgcu_external_semaphore *semaphore, uint64_t value, void *stream) {
  if (!semaphore) return lgcu_set_last_error(LGCU_ERR_INVALID_ARGUMENT, "semaphore_signal: null semaphore");
  cudaExternalSemaphoreSignalParams p;
  std::memset(&p, 0, sizeof(p));
  p.params.fence.value = value;
  const cudaError_t e = cudaSignalExternalSemaphoresAsync(&semaphore->handle, &p, 1, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? LGCU_OK : lgcu_set_last_error(LGCU_ERR_CUDA, "cudaSignalExternalSemaphoresAsync: %s", cudaGetErrorString(e));
}

} // extern "C"
