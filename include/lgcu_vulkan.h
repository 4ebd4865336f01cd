/*
 * lgcu_vulkan.h — Vulkan-typed convenience layer over lgcu_interop.h. Compile-guarded: define LGCU_WITH_VULKAN and have
 * <vulkan/vulkan_core.h> on the include path (the development image has neither a Vulkan SDK nor a loader — SURVEY.md F4 — so the
 * repository compiles this header against tests/stubs/vulkan/vulkan_core.h only; nothing in it calls into Vulkan, it converts types).
 *
 * What a LegitEngine maintainer adds on the Vulkan side (INTEGRATION.md §4 has the full listing):
 *   LV/CoreImpl.h:44-47      device extensions += VK_KHR_external_memory_fd, VK_KHR_external_semaphore_fd (+ timeline semaphores, core 1.2)
 *   LV/Image.h:230-248       image create: VkExternalMemoryImageCreateInfo{OPAQUE_FD}, tiling LINEAR for the images the CUDA passes touch;
 *                            allocate: VkExportMemoryAllocateInfo{OPAQUE_FD} (+ VkMemoryDedicatedAllocateInfo); vkGetMemoryFdKHR -> fd
 *   LV/PresentQueue.h:122-166  one timeline VkSemaphore with VkExportSemaphoreCreateInfo{OPAQUE_FD}; vkGetSemaphoreFdKHR -> fd; EndFrame
 *                            submits twice around the CUDA section (signal S = 2f+1 / wait S = 2f+2)
 */
#ifndef LGCU_VULKAN_H
#define LGCU_VULKAN_H

#ifdef LGCU_WITH_VULKAN

#include <vulkan/vulkan_core.h>

#include "lgcu_interop.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lgcu_format values ARE the VkFormat enumerants (lgcu.h), so the mapping is a range check. */
static inline uint32_t lgcu_vk_format(VkFormat format) {
  switch (format) {
  case VK_FORMAT_B8G8R8A8_SRGB:
  case VK_FORMAT_R16G16B16A16_SFLOAT:
  case VK_FORMAT_R32G32_SFLOAT:
  case VK_FORMAT_R32G32B32A32_SFLOAT:
  case VK_FORMAT_D32_SFLOAT:
    return (uint32_t)format;
  default:
    return LGCU_FORMAT_UNDEFINED;
  }
}

/* An exported LINEAR image of `mips` levels bound at `bindOffset` of the imported allocation mapped at devicePtr:
 * layouts[l] = vkGetImageSubresourceLayout(device, image, {aspect, mipLevel l, arrayLayer 0}). -> the lgcu_image the passes take. */
static inline int lgcu_vk_image(void *devicePtr, VkDeviceSize bindOffset, VkFormat format, VkExtent3D extent, uint32_t mips, const VkSubresourceLayout *layouts,
                                lgcu_image *image) {
  uint64_t offsets[LGCU_MAX_MIPS], pitches[LGCU_MAX_MIPS];
  if (!layouts || mips == 0 || mips > LGCU_MAX_MIPS || extent.depth != 1) return LGCU_ERR_INVALID_ARGUMENT;
  for (uint32_t l = 0; l < mips; l++) {
    offsets[l] = (uint64_t)layouts[l].offset;
    pitches[l] = (uint64_t)layouts[l].rowPitch;
  }
  return lgcu_image_from_linear_layout((char *)devicePtr + bindOffset, lgcu_vk_format(format), extent.width, extent.height, mips, offsets, pitches, image);
}

/* The frame's CUDA section between the two submits of EndFrame: wait for submit A (value 2f+1), run `record` (the lgcu_* calls or one
 * cudaGraphLaunch on `stream`), signal for submit B (value 2f+2). */
static inline int lgcu_vk_cuda_section(lgcu_external_semaphore *timeline, uint64_t frameIndex, void *stream, int (*record)(void *user, void *stream), void *user) {
  int st = lgcu_semaphore_wait(timeline, 2 * frameIndex + 1, stream);
  if (st != LGCU_OK) return st;
  st = record(user, stream);
  if (st != LGCU_OK) return st;
  return lgcu_semaphore_signal(timeline, 2 * frameIndex + 2, stream);
}

#ifdef __cplusplus
}
#endif

#endif /* LGCU_WITH_VULKAN */
#endif
