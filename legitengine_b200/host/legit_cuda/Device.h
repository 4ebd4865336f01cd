// Device.h — device-side resources behind the rendergraph: images (linear mip chains in HBM), image views, buffers.
// They play the role of legit::ImageData / legit::ImageView / legit::Buffer (LV/Image.h, LV/ImageView.h, LV/Buffer.h);
// the metadata contract (format, per-level size, view sub-range) is the same, the storage is plain cudaMalloc memory in
// the canonical layout of lgcu_image_layout() instead of an opaque VkImage.
#pragma once

#include <cuda_runtime_api.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../../include/lgcu.h"
#include "Vk.h"

namespace legit_cuda {

inline void CudaCheck(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
// C-ABI status -> exception, like the reference's vk::*Error exceptions (SURVEY.md §8b "Error convention")
inline void LgcuCheck(int status, const char *what) {
  if (status != LGCU_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(status) + "): " + lgcu_last_error());
}

inline bool IsDepthFormat(vk::Format f) { return f == vk::Format::eD32Sfloat; }

class ImageData {
public:
  // owning: allocates `mips` levels of width x height
  ImageData(vk::Format format, glm::uvec2 size, uint32_t mips) : owns_(true) {
    const uint64_t bytes = lgcu_image_layout(&desc_, uint32_t(format), size.x, size.y, mips);
    if (!bytes) throw std::runtime_error("ImageData: unsupported format / mip count");
    CudaCheck(cudaMalloc(&desc_.base, bytes), "cudaMalloc(image)");
    bytes_ = bytes;
  }
  // non-owning: wraps externally allocated memory (e.g. imported from Vulkan) that already follows `desc`
  explicit ImageData(const lgcu_image &desc) : desc_(desc), owns_(false) {}
  ImageData(const ImageData &) = delete;
  ImageData &operator=(const ImageData &) = delete;
  ~ImageData() {
    if (owns_ && desc_.base) cudaFree(desc_.base);
  }
  vk::Format GetFormat() const { return vk::Format(desc_.format); }
  uint32_t GetMipsCount() const { return desc_.imageMipCount; }
  glm::uvec2 GetMipSize(uint32_t level) const { return glm::uvec2(desc_.width >> level, desc_.height >> level); }
  const lgcu_image &GetDesc() const { return desc_; }
  uint64_t GetByteSize() const { return bytes_; }
  void *GetLevelPointer(uint32_t level) const { return static_cast<uint8_t *>(desc_.base) + desc_.levelOffset[level]; }
  uint32_t GetLevelPitch(uint32_t level) const { return desc_.levelPitch[level]; }

private:
  lgcu_image desc_{};
  uint64_t bytes_ = 0;
  bool owns_;
};

// A mip sub-range of an image (LV/ImageView.h:21-31). `GetDesc()` is what the lgcu_* entry points take.
class ImageView {
public:
  ImageView(ImageData *image, uint32_t baseMipLevel, uint32_t mipLevelsCount) : image_(image), desc_(image->GetDesc()) {
    desc_.baseMip = baseMipLevel;
    desc_.mipCount = mipLevelsCount;
  }
  ImageData *GetImageData() const { return image_; }
  uint32_t GetBaseMipLevel() const { return desc_.baseMip; }
  uint32_t GetMipLevelsCount() const { return desc_.mipCount; }
  const lgcu_image *GetDesc() const { return &desc_; }

private:
  ImageData *image_;
  lgcu_image desc_;
};

class Buffer {
public:
  explicit Buffer(size_t bytes) : bytes_(bytes), owns_(true) { CudaCheck(cudaMalloc(&ptr_, bytes ? bytes : 1), "cudaMalloc(buffer)"); }
  Buffer(void *devicePtr, size_t bytes) : ptr_(devicePtr), bytes_(bytes), owns_(false) {}
  Buffer(const Buffer &) = delete;
  Buffer &operator=(const Buffer &) = delete;
  ~Buffer() {
    if (owns_ && ptr_) cudaFree(ptr_);
  }
  void *GetHandle() const { return ptr_; }
  size_t GetSize() const { return bytes_; }

private:
  void *ptr_ = nullptr;
  size_t bytes_ = 0;
  bool owns_;
};

// Sampler state is metadata only: the kernels implement clamp-to-edge + linear / nearest filtering in software
// (SURVEY.md Appendix B). Kept so renderer constructors read like the reference's (SSVGIRenderer.h:17-18).
enum struct SamplerAddressMode { eClampToEdge };
enum struct Filter { eNearest, eLinear };
enum struct SamplerMipmapMode { eNearest, eLinear };
struct Sampler {
  Sampler(SamplerAddressMode a, Filter f, SamplerMipmapMode m, bool compare = false) : addressMode(a), filter(f), mipMode(m), useComparison(compare) {}
  SamplerAddressMode addressMode;
  Filter filter;
  SamplerMipmapMode mipMode;
  bool useComparison;
};

} // namespace legit_cuda
