// legit_cuda::ExternalImage / TimelineSemaphore — the rendergraph-side face of the Vulkan hand-back (include/lgcu_interop.h).
//
// In the reference every image lives in Vulkan memory (LV/Image.h:230-248) and reaches a pass through an image-view proxy; images
// the graph does not own enter it with RenderGraph::AddExternalImageView(ImageView *, usage) (LV/RenderGraph.h:305-335), which is how
// the swapchain image gets in (LV/PresentQueue.h:106-110). An ExternalImage is exactly such an ImageView over memory that Vulkan
// allocated and exported: imported once (cudaImportExternalMemory), described with the layout Vulkan reports for it, and then used like
// any other view — `graph->AddExternalImageView(external.GetView(), usage)`. The TimelineSemaphore orders the CUDA section of a frame
// between the two submits of InFlightQueue::EndFrame (LV/PresentQueue.h:122-166).
#pragma once

#include <memory>
#include <stdexcept>
#include <string>

#include "../../../include/lgcu_interop.h"
#include "Device.h"

namespace legit_cuda {

class ExternalImage {
public:
  // fd: vkGetMemoryFdKHR(OPAQUE_FD) of the image's memory (ownership passes to CUDA); bindOffset: vkBindImageMemory's offset;
  // levelOffsets / rowPitches: VkSubresourceLayout.offset / .rowPitch of every mip level of the LINEAR image.
  ExternalImage(int fd, uint64_t allocationSize, bool dedicated, uint64_t bindOffset, vk::Format format, glm::uvec2 size, uint32_t mips, const uint64_t *levelOffsets,
                const uint64_t *rowPitches) {
    void *mapped = nullptr;
    Check(lgcu_import_memory_fd(fd, allocationSize, dedicated ? 1 : 0, &memory_, &mapped), "lgcu_import_memory_fd");
    lgcu_image desc;
    const int st = lgcu_image_from_linear_layout(static_cast<char *>(mapped) + bindOffset, uint32_t(format), size.x, size.y, mips, levelOffsets, rowPitches, &desc);
    if (st != LGCU_OK) {
      lgcu_release_memory(memory_);
      Check(st, "lgcu_image_from_linear_layout");
    }
    image_.reset(new ImageData(desc)); // non-owning: the memory belongs to the Vulkan allocation
    view_.reset(new ImageView(image_.get(), 0, mips));
  }
  ExternalImage(const ExternalImage &) = delete;
  ExternalImage &operator=(const ExternalImage &) = delete;
  ~ExternalImage() {
    view_.reset();
    image_.reset();
    lgcu_release_memory(memory_);
  }
  ImageData *GetImageData() const { return image_.get(); }
  ImageView *GetView() const { return view_.get(); } // for RenderGraph::AddExternalImageView

private:
  static void Check(int status, const char *what) {
    if (status != LGCU_OK) throw std::runtime_error(std::string(what) + ": " + lgcu_last_error());
  }
  lgcu_external_memory *memory_ = nullptr;
  std::unique_ptr<ImageData> image_;
  std::unique_ptr<ImageView> view_;
};

class TimelineSemaphore {
public:
  explicit TimelineSemaphore(int fd) { // vkGetSemaphoreFdKHR(OPAQUE_FD) of a VK_SEMAPHORE_TYPE_TIMELINE semaphore
    if (lgcu_import_timeline_semaphore_fd(fd, &semaphore_) != LGCU_OK) throw std::runtime_error(std::string("lgcu_import_timeline_semaphore_fd: ") + lgcu_last_error());
  }
  TimelineSemaphore(const TimelineSemaphore &) = delete;
  TimelineSemaphore &operator=(const TimelineSemaphore &) = delete;
  ~TimelineSemaphore() { lgcu_release_semaphore(semaphore_); }
  // RenderGraph::Execute(stream) of frame f sits between these two (submit A signals 2f+1, submit B waits for 2f+2)
  void WaitForVulkan(uint64_t frameIndex, cudaStream_t stream) { Check(lgcu_semaphore_wait(semaphore_, 2 * frameIndex + 1, stream), "lgcu_semaphore_wait"); }
  void SignalVulkan(uint64_t frameIndex, cudaStream_t stream) { Check(lgcu_semaphore_signal(semaphore_, 2 * frameIndex + 2, stream), "lgcu_semaphore_signal"); }

private:
  static void Check(int status, const char *what) {
    if (status != LGCU_OK) throw std::runtime_error(std::string(what) + ": " + lgcu_last_error());
  }
  lgcu_external_semaphore *semaphore_ = nullptr;
};

} // namespace legit_cuda
