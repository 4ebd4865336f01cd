/*
 * lgcu_interop.h — hand-back to Vulkan: external memory / timeline-semaphore interop of the CUDA passes.
 *
 * The reference allocates every image inside Vulkan (LV/Image.h:230-248: vk::MemoryAllocateInfo without an export chain) and
 * submits one command buffer per frame (LV/PresentQueue.h:122-166). For the CUDA passes to work on the same memory and to be
 * ordered against the two halves of the split submit, the engine exports
 *   - the memory behind every image / buffer the passes touch as an opaque fd (VK_KHR_external_memory_fd, vkGetMemoryFdKHR), and
 *   - one timeline semaphore as an opaque fd (VK_KHR_external_semaphore_fd, vkGetSemaphoreFdKHR),
 * and this layer imports them: cudaImportExternalMemory / cudaExternalMemoryGetMappedBuffer, cudaImportExternalSemaphore,
 * cudaWaitExternalSemaphoresAsync / cudaSignalExternalSemaphoresAsync. Nothing here needs a Vulkan header: the fds and the
 * subresource layouts are plain integers. `lgcu_vulkan.h` adds the Vulkan-typed convenience layer (compile-guarded).
 *
 * All entry points return lgcu_status. Imports and releases are host-synchronous set-up calls (once per swapchain re-creation);
 * the semaphore wait / signal calls only enqueue on `stream`, like every pass.
 */
#ifndef LGCU_INTEROP_H
#define LGCU_INTEROP_H

#include "lgcu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lgcu_external_memory lgcu_external_memory;       /* one imported VkDeviceMemory */
typedef struct lgcu_external_semaphore lgcu_external_semaphore; /* one imported timeline VkSemaphore */

/* Imports the VkDeviceMemory behind an image or buffer (the fd of vkGetMemoryFdKHR with
 * VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT; replaces nothing in the reference — LV/Image.h:230-248 gains a
 * VkExportMemoryAllocateInfo in its pNext chain). On success CUDA owns the fd (do not close it), *devicePtr maps the whole
 * allocation [0, allocationSize). dedicated != 0 for memory allocated with VkMemoryDedicatedAllocateInfo. */
int lgcu_import_memory_fd(int fd, uint64_t allocationSize, int dedicated, lgcu_external_memory **memory, void **devicePtr);
int lgcu_release_memory(lgcu_external_memory *memory);

/* Describes a LINEAR-tiling image (or a buffer laid out like one) that lives at `base` — typically devicePtr + the image's bind
 * offset — as an lgcu_image: level l at base + levelOffsets[l] with rows rowPitches[l] bytes apart (VkSubresourceLayout.offset /
 * .rowPitch of vkGetImageSubresourceLayout for VK_IMAGE_ASPECT_COLOR/DEPTH, mip l; what PassContext::GetImageView resolves to,
 * LV/RenderGraph.h:421-451). Validates what the kernels rely on: a known format, 1..LGCU_MAX_MIPS levels, rows of at least
 * width_l texels, pitches and level offsets that are multiples of 16 bytes (128-bit accesses), base 16-byte aligned. Pure host code. */
int lgcu_image_from_linear_layout(void *base, uint32_t format, uint32_t width, uint32_t height, uint32_t mips, const uint64_t *levelOffsets,
                                  const uint64_t *rowPitches, lgcu_image *image);

/* Imports a timeline semaphore (fd of vkGetSemaphoreFdKHR, VK_SEMAPHORE_TYPE_TIMELINE, OPAQUE_FD). CUDA owns the fd on success. */
int lgcu_import_timeline_semaphore_fd(int fd, lgcu_external_semaphore **semaphore);
int lgcu_release_semaphore(lgcu_external_semaphore *semaphore);
/* Stream-ordered: the work enqueued on `stream` after the wait starts once the semaphore reaches `value` (submit A of
 * INTEGRATION.md §4 signals it); the signal sets it to `value` once the work enqueued before it is done (submit B waits for it).
 * Replaces the implicit ordering inside the reference's single submit (LV/PresentQueue.h:134-166). */
int lgcu_semaphore_wait(lgcu_external_semaphore *semaphore, uint64_t value, void *stream);
int lgcu_semaphore_signal(lgcu_external_semaphore *semaphore, uint64_t value, void *stream);

#ifdef __cplusplus
}
#endif
#endif
