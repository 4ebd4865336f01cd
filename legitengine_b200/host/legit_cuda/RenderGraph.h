// RenderGraph.h — legit_cuda::RenderGraph: the reference's rendergraph API (LV/RenderGraph.h:285-776) with a CUDA
// stream executor behind it.
//
// Kept from the reference: proxy ids and RAII proxy handles, AddImage / AddImageView / AddExternalImage(View) /
// AddBuffer / AddExternalBuffer / GetMipSize, the RenderPassDesc / ComputePassDesc / TransferPassDesc builders with the
// same setter names and defaults (default colour clear (1, .5, 0, 1), depth clear 1.0, load op DontCare), AddPass
// overloads, strict AddPass-order execution (LV/RenderGraph.h:784-788: no reordering, culling or aliasing), transient
// image pooling keyed by (format, mips, layers, usage, size) with per-frame use counters (ImageCache, :29-86), and
// the rule that a pass callback only sees the views it declared (:797-819).
//
// Replaced: Execute() takes a cudaStream_t instead of a vk::CommandBuffer. Passes are enqueued on that one stream in
// order, so stream order is the barrier (the reference infers vkCmdPipelineBarrier from declared usage, :824-863);
// there are no render passes / framebuffers / descriptor sets. The callback receives resolved image views whose
// GetDesc() is the lgcu_image the C ABI (include/lgcu.h) consumes.
//
// Necessary extension: in Vulkan the colour / depth attachments are bound implicitly through the
// framebuffer and never handed to the callback; a CUDA pass writes them explicitly, so attachment views are resolved
// for the callback too (GetImageView(attachmentId), GetColorAttachment(i), GetDepthAttachment()), together with their
// load op and clear value. A pass is a full-screen kernel that overwrites its whole render area; for eClear the kernel
// writes the clear value where it covers nothing (that is how lgcu_gbuffer_resolve treats uncovered pixels).
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "Device.h"
#include "Handles.h"
#include "Pool.h"
#include "Profiler.h"
#include "Vk.h"

namespace legit_cuda {

class RenderGraph {
private:
  struct ImageProxy;
  struct ImageViewProxy;
  struct BufferProxy;

public:
  using ImageProxyId = Utils::Pool<ImageProxy>::Id;
  using ImageViewProxyId = Utils::Pool<ImageViewProxy>::Id;
  using BufferProxyId = Utils::Pool<BufferProxy>::Id;

  struct ImageHandleInfo {
    ImageHandleInfo() = default;
    ImageHandleInfo(RenderGraph *g, ImageProxyId id) : graph(g), id(id) {}
    void Reset() { graph->imageProxies_.Release(id); }
    ImageProxyId Id() const { return id; }
    void SetDebugName(const std::string &name) const { graph->imageProxies_.Get(id).debugName = name; }

  private:
    RenderGraph *graph = nullptr;
    ImageProxyId id;
  };
  struct ImageViewHandleInfo {
    ImageViewHandleInfo() = default;
    ImageViewHandleInfo(RenderGraph *g, ImageViewProxyId id) : graph(g), id(id) {}
    void Reset() { graph->imageViewProxies_.Release(id); }
    ImageViewProxyId Id() const { return id; }
    void SetDebugName(const std::string &name) const { graph->imageViewProxies_.Get(id).debugName = name; }

  private:
    RenderGraph *graph = nullptr;
    ImageViewProxyId id;
  };
  struct BufferHandleInfo {
    BufferHandleInfo() = default;
    BufferHandleInfo(RenderGraph *g, BufferProxyId id) : graph(g), id(id) {}
    void Reset() { graph->bufferProxies_.Release(id); }
    BufferProxyId Id() const { return id; }

  private:
    RenderGraph *graph = nullptr;
    BufferProxyId id;
  };

  using ImageProxyUnique = UniqueHandle<ImageHandleInfo, RenderGraph>;
  using ImageViewProxyUnique = UniqueHandle<ImageViewHandleInfo, RenderGraph>;
  using BufferProxyUnique = UniqueHandle<BufferHandleInfo, RenderGraph>;

  RenderGraph() = default;
  RenderGraph(const RenderGraph &) = delete;
  RenderGraph &operator=(const RenderGraph &) = delete;

  // ---- resource declaration (LV/RenderGraph.h:285-335, 400-419) -------------------------------------------------
  ImageProxyUnique AddImage(vk::Format format, uint32_t mipsCount, uint32_t arrayLayersCount, glm::uvec2 size, vk::ImageUsageFlags usageFlags) {
    return AddImage(format, mipsCount, arrayLayersCount, glm::uvec3(size.x, size.y, uint32_t(-1)), usageFlags);
  }
  ImageProxyUnique AddImage(vk::Format format, uint32_t mipsCount, uint32_t arrayLayersCount, glm::uvec3 size, vk::ImageUsageFlags usageFlags) {
    if (size.z != uint32_t(-1) || arrayLayersCount != 1) throw std::runtime_error("legit_cuda::RenderGraph: only single-layer 2D images are supported (hot path scope)");
    ImageProxy proxy;
    proxy.external = nullptr;
    proxy.key = ImageKey{format, usageFlags, mipsCount, arrayLayersCount, size};
    ImageProxyId id = imageProxies_.Add(std::move(proxy));
    imageProxies_.Get(id).debugName = "Graph image [" + std::to_string(size.x) + ", " + std::to_string(size.y) + ", Id=" + std::to_string(id.asInt) + "]";
    return ImageProxyUnique(ImageHandleInfo(this, id));
  }
  ImageProxyUnique AddExternalImage(ImageData *image) {
    ImageProxy proxy;
    proxy.external = image;
    proxy.debugName = "External graph image";
    return ImageProxyUnique(ImageHandleInfo(this, imageProxies_.Add(std::move(proxy))));
  }
  ImageViewProxyUnique AddImageView(ImageProxyId imageProxyId, uint32_t baseMipLevel, uint32_t mipLevelsCount, uint32_t baseArrayLayer, uint32_t arrayLayersCount) {
    ImageViewProxy proxy;
    proxy.imageProxyId = imageProxyId;
    proxy.baseMipLevel = baseMipLevel;
    proxy.mipLevelsCount = mipLevelsCount;
    proxy.baseArrayLayer = baseArrayLayer;
    proxy.arrayLayersCount = arrayLayersCount;
    proxy.debugName = "View";
    return ImageViewProxyUnique(ImageViewHandleInfo(this, imageViewProxies_.Add(std::move(proxy))));
  }
  ImageViewProxyUnique AddExternalImageView(ImageView *imageView, ImageUsageTypes usageType = ImageUsageTypes::Unknown) {
    ImageViewProxy proxy;
    proxy.external = imageView;
    proxy.externalUsageType = usageType;
    proxy.debugName = "External view";
    return ImageViewProxyUnique(ImageViewHandleInfo(this, imageViewProxies_.Add(std::move(proxy))));
  }

  // LV/RenderGraph.h:362-397
  glm::uvec2 GetMipSize(ImageProxyId imageProxyId, uint32_t mipLevel) {
    const ImageProxy &proxy = imageProxies_.Get(imageProxyId);
    if (proxy.external) return proxy.external->GetMipSize(mipLevel);
    const uint32_t mipMult = 1u << mipLevel;
    return glm::uvec2(proxy.key.size.x / mipMult, proxy.key.size.y / mipMult);
  }
  glm::uvec2 GetMipSize(ImageViewProxyId imageViewProxyId, uint32_t mipOffset) {
    const ImageViewProxy &proxy = imageViewProxies_.Get(imageViewProxyId);
    if (proxy.external) return proxy.external->GetImageData()->GetMipSize(proxy.external->GetBaseMipLevel() + mipOffset);
    return GetMipSize(proxy.imageProxyId, proxy.baseMipLevel + mipOffset);
  }

  template <typename BufferType> BufferProxyUnique AddBuffer(uint32_t count) {
    BufferProxy proxy;
    proxy.elementSize = uint32_t(sizeof(BufferType));
    proxy.elementsCount = count;
    return BufferProxyUnique(BufferHandleInfo(this, bufferProxies_.Add(std::move(proxy))));
  }
  BufferProxyUnique AddExternalBuffer(Buffer *buffer) {
    BufferProxy proxy;
    proxy.external = buffer;
    proxy.elementSize = proxy.elementsCount = uint32_t(-1);
    return BufferProxyUnique(BufferHandleInfo(this, bufferProxies_.Add(std::move(proxy))));
  }

  // ---- pass callback contexts (LV/RenderGraph.h:421-451) ----------------------------------------------------------
  struct PassContext {
    ImageView *GetImageView(ImageViewProxyId id) { return resolvedImageViews[id.asInt]; }
    Buffer *GetBuffer(BufferProxyId id) { return resolvedBuffers[id.asInt]; }
    cudaStream_t GetStream() { return stream; }
    cudaStream_t GetCommandBuffer() { return stream; } // the stream stands where the reference hands out a vk::CommandBuffer

  private:
    std::vector<ImageView *> resolvedImageViews;
    std::vector<Buffer *> resolvedBuffers;
    cudaStream_t stream = nullptr;
    friend class RenderGraph;
  };

  struct RenderPassDesc;
  struct RenderPassContext : public PassContext {
    // Attachment access: the extension described in the header comment.
    size_t GetColorAttachmentsCount() const { return colorViews.size(); }
    ImageView *GetColorAttachment(size_t index) { return colorViews[index]; }
    ImageView *GetDepthAttachment() { return depthView; }
    vk::AttachmentLoadOp GetColorLoadOp(size_t index) const;
    vk::ClearValue GetColorClearValue(size_t index) const;
    vk::AttachmentLoadOp GetDepthLoadOp() const;
    vk::ClearValue GetDepthClearValue() const;
    vk::Extent2D GetRenderAreaExtent() const;
    const RenderPassDesc *GetRenderPass() const { return desc; } // where the reference returns a legit::RenderPass*

  private:
    std::vector<ImageView *> colorViews;
    ImageView *depthView = nullptr;
    const RenderPassDesc *desc = nullptr;
    friend class RenderGraph;
  };

  // ---- pass descriptions (LV/RenderGraph.h:453-551, 591-708) ------------------------------------------------------
  struct RenderPassDesc {
    RenderPassDesc() : profilerTaskName("RenderPass"), profilerTaskColor(Colors::orange) {}
    struct Attachment {
      ImageViewProxyId imageViewProxyId;
      vk::AttachmentLoadOp loadOp = vk::AttachmentLoadOp::eDontCare;
      vk::ClearValue clearValue;
    };
    RenderPassDesc &SetColorAttachments(const std::vector<ImageViewProxyId> &views, vk::AttachmentLoadOp loadOp = vk::AttachmentLoadOp::eDontCare,
                                        vk::ClearValue clearValue = vk::ClearColorValue(std::array<float, 4>{1.0f, 0.5f, 0.0f, 1.0f})) {
      colorAttachments.clear();
      for (const ImageViewProxyId &v : views) colorAttachments.push_back(Attachment{v, loadOp, clearValue});
      return *this;
    }
    RenderPassDesc &SetColorAttachments(std::vector<Attachment> &&attachments) {
      colorAttachments = std::move(attachments);
      return *this;
    }
    RenderPassDesc &SetDepthAttachment(ImageViewProxyId view, vk::AttachmentLoadOp loadOp = vk::AttachmentLoadOp::eDontCare,
                                       vk::ClearValue clearValue = vk::ClearDepthStencilValue(1.0f, 0)) {
      depthAttachment = Attachment{view, loadOp, clearValue};
      return *this;
    }
    RenderPassDesc &SetDepthAttachment(Attachment attachment) {
      depthAttachment = attachment;
      return *this;
    }
    RenderPassDesc &SetVertexBuffers(std::vector<BufferProxyId> &&buffers) {
      vertexBufferProxies = std::move(buffers);
      return *this;
    }
    RenderPassDesc &SetInputImages(std::vector<ImageViewProxyId> &&views) {
      inputImageViewProxies = std::move(views);
      return *this;
    }
    RenderPassDesc &SetStorageBuffers(std::vector<BufferProxyId> &&buffers) {
      inoutStorageBufferProxies = std::move(buffers);
      return *this;
    }
    RenderPassDesc &SetStorageImages(std::vector<ImageViewProxyId> &&views) {
      inoutStorageImageProxies = std::move(views);
      return *this;
    }
    RenderPassDesc &SetRenderAreaExtent(vk::Extent2D extent) {
      renderAreaExtent = extent;
      return *this;
    }
    RenderPassDesc &SetRecordFunc(std::function<void(RenderPassContext)> func) {
      recordFunc = std::move(func);
      return *this;
    }
    RenderPassDesc &SetProfilerInfo(uint32_t taskColor, std::string taskName) {
      profilerTaskColor = taskColor;
      profilerTaskName = std::move(taskName);
      return *this;
    }

    std::vector<Attachment> colorAttachments;
    Attachment depthAttachment;
    std::vector<ImageViewProxyId> inputImageViewProxies;
    std::vector<BufferProxyId> vertexBufferProxies;
    std::vector<BufferProxyId> inoutStorageBufferProxies;
    std::vector<ImageViewProxyId> inoutStorageImageProxies;
    vk::Extent2D renderAreaExtent;
    std::function<void(RenderPassContext)> recordFunc;
    std::string profilerTaskName;
    uint32_t profilerTaskColor;
  };

  struct ComputePassDesc {
    ComputePassDesc() : profilerTaskName("ComputePass"), profilerTaskColor(Colors::belizeHole) {}
    ComputePassDesc &SetInputImages(std::vector<ImageViewProxyId> &&views) {
      inputImageViewProxies = std::move(views);
      return *this;
    }
    ComputePassDesc &SetStorageBuffers(std::vector<BufferProxyId> &&buffers) {
      inoutStorageBufferProxies = std::move(buffers);
      return *this;
    }
    ComputePassDesc &SetStorageImages(std::vector<ImageViewProxyId> &&views) {
      inoutStorageImageProxies = std::move(views);
      return *this;
    }
    ComputePassDesc &SetRecordFunc(std::function<void(PassContext)> func) {
      recordFunc = std::move(func);
      return *this;
    }
    ComputePassDesc &SetProfilerInfo(uint32_t taskColor, std::string taskName) {
      profilerTaskColor = taskColor;
      profilerTaskName = std::move(taskName);
      return *this;
    }
    std::vector<BufferProxyId> inoutStorageBufferProxies;
    std::vector<ImageViewProxyId> inputImageViewProxies;
    std::vector<ImageViewProxyId> inoutStorageImageProxies;
    std::function<void(PassContext)> recordFunc;
    std::string profilerTaskName;
    uint32_t profilerTaskColor;
  };

  struct TransferPassDesc {
    TransferPassDesc() : profilerTaskName("TransferPass"), profilerTaskColor(Colors::silver) {}
    TransferPassDesc &SetSrcImages(std::vector<ImageViewProxyId> &&views) {
      srcImageViewProxies = std::move(views);
      return *this;
    }
    TransferPassDesc &SetDstImages(std::vector<ImageViewProxyId> &&views) {
      dstImageViewProxies = std::move(views);
      return *this;
    }
    TransferPassDesc &SetSrcBuffers(std::vector<BufferProxyId> &&buffers) {
      srcBufferProxies = std::move(buffers);
      return *this;
    }
    TransferPassDesc &SetDstBuffers(std::vector<BufferProxyId> &&buffers) {
      dstBufferProxies = std::move(buffers);
      return *this;
    }
    TransferPassDesc &SetRecordFunc(std::function<void(PassContext)> func) {
      recordFunc = std::move(func);
      return *this;
    }
    TransferPassDesc &SetProfilerInfo(uint32_t taskColor, std::string taskName) {
      profilerTaskColor = taskColor;
      profilerTaskName = std::move(taskName);
      return *this;
    }
    std::vector<BufferProxyId> srcBufferProxies, dstBufferProxies;
    std::vector<ImageViewProxyId> srcImageViewProxies, dstImageViewProxies;
    std::function<void(PassContext)> recordFunc;
    std::string profilerTaskName;
    uint32_t profilerTaskColor;
  };

  struct ImagePresentPassDesc {
    ImagePresentPassDesc &SetImage(ImageViewProxyId id) {
      presentImageViewProxyId = id;
      return *this;
    }
    ImageViewProxyId presentImageViewProxyId;
  };
  struct FrameSyncBeginPassDesc {};
  struct FrameSyncEndPassDesc {};

  // ---- pass registration (LV/RenderGraph.h:553-774) ---------------------------------------------------------------
  void AddPass(RenderPassDesc &desc) {
    tasks_.push_back(Task{Task::Type::RenderPass, renderPassDescs_.size()});
    renderPassDescs_.emplace_back(desc);
  }
  void AddPass(ComputePassDesc &desc) {
    tasks_.push_back(Task{Task::Type::ComputePass, computePassDescs_.size()});
    computePassDescs_.emplace_back(desc);
  }
  void AddPass(TransferPassDesc &desc) {
    tasks_.push_back(Task{Task::Type::TransferPass, transferPassDescs_.size()});
    transferPassDescs_.emplace_back(desc);
  }
  void AddPass(ImagePresentPassDesc &&desc) {
    tasks_.push_back(Task{Task::Type::ImagePresent, imagePresentDescs_.size()});
    imagePresentDescs_.push_back(desc);
  }
  void AddPass(FrameSyncBeginPassDesc &&) { tasks_.push_back(Task{Task::Type::FrameSyncBegin, 0}); }
  void AddPass(FrameSyncEndPassDesc &&) { tasks_.push_back(Task{Task::Type::FrameSyncEnd, 0}); }
  void AddImagePresent(ImageViewProxyId presentImageViewProxyId) {
    ImagePresentPassDesc desc;
    desc.presentImageViewProxyId = presentImageViewProxyId;
    AddPass(std::move(desc));
  }

  void AddRenderPass(std::vector<ImageViewProxyId> colorAttachmentImageProxies, ImageViewProxyId depthAttachmentImageProxy,
                     std::vector<ImageViewProxyId> inputImageViewProxies, vk::Extent2D renderAreaExtent, vk::AttachmentLoadOp loadOp,
                     std::function<void(RenderPassContext)> recordFunc) {
    RenderPassDesc desc;
    for (const auto &proxy : colorAttachmentImageProxies)
      desc.colorAttachments.push_back(RenderPassDesc::Attachment{proxy, loadOp, vk::ClearColorValue(std::array<float, 4>{0.03f, 0.03f, 0.03f, 1.0f})});
    desc.depthAttachment = RenderPassDesc::Attachment{depthAttachmentImageProxy, loadOp, vk::ClearDepthStencilValue(1.0f, 0)};
    desc.inputImageViewProxies = std::move(inputImageViewProxies);
    desc.renderAreaExtent = renderAreaExtent;
    desc.recordFunc = std::move(recordFunc);
    AddPass(desc);
  }
  void AddComputePass(std::vector<BufferProxyId> inoutBufferProxies, std::vector<ImageViewProxyId> inputImageViewProxies,
                      std::function<void(PassContext)> recordFunc) {
    ComputePassDesc desc;
    desc.inoutStorageBufferProxies = std::move(inoutBufferProxies);
    desc.inputImageViewProxies = std::move(inputImageViewProxies);
    desc.recordFunc = std::move(recordFunc);
    AddPass(desc);
  }

  // Drops every declared pass and pooled allocation (LV/RenderGraph.h:586-589); live proxy handles become dangling,
  // exactly as in the reference, so callers destroy their proxies first.
  void Clear() {
    ClearPasses();
    imageCache_.clear();
    imageViewCache_.clear();
    bufferCache_.clear();
  }

  size_t GetPendingPassCount() const { return tasks_.size(); }

  // ---- execution (replaces LV/RenderGraph.h:776-1091) ---------------------------------------------------------------
  // Resolves proxies to device memory, then invokes every pass callback synchronously on the calling thread, in
  // AddPass order; the callbacks enqueue kernels on `stream`. Returns after enqueueing (no host/device sync).
  void Execute(cudaStream_t stream, CpuProfiler *cpuProfiler = nullptr, GpuProfiler *gpuProfiler = nullptr) {
    ResolveImages();
    ResolveImageViews();
    ResolveBuffers();
    if (cpuProfiler) cpuProfiler->StartFrame();
    if (gpuProfiler) gpuProfiler->StartFrame();

    for (const Task &task : tasks_) {
      switch (task.type) {
        case Task::Type::RenderPass: {
          RenderPassDesc &desc = renderPassDescs_[task.index];
          if (gpuProfiler) gpuProfiler->StartTask(stream, desc.profilerTaskName, desc.profilerTaskColor);
          const size_t cpuTask = cpuProfiler ? cpuProfiler->StartTask(desc.profilerTaskName, desc.profilerTaskColor) : 0;
          RenderPassContext ctx;
          InitContext(ctx, stream);
          for (auto id : desc.inputImageViewProxies) Bind(ctx, id);
          for (auto id : desc.inoutStorageImageProxies) Bind(ctx, id);
          for (auto id : desc.inoutStorageBufferProxies) Bind(ctx, id);
          for (auto id : desc.vertexBufferProxies) Bind(ctx, id);
          for (auto &attachment : desc.colorAttachments) {
            Bind(ctx, attachment.imageViewProxyId);
            ctx.colorViews.push_back(ctx.resolvedImageViews[attachment.imageViewProxyId.asInt]);
          }
          if (desc.depthAttachment.imageViewProxyId.IsValid()) {
            Bind(ctx, desc.depthAttachment.imageViewProxyId);
            ctx.depthView = ctx.resolvedImageViews[desc.depthAttachment.imageViewProxyId.asInt];
          }
          ctx.desc = &desc;
          if (desc.recordFunc) desc.recordFunc(ctx);
          if (cpuProfiler) cpuProfiler->EndTask(cpuTask);
        } break;
        case Task::Type::ComputePass: {
          ComputePassDesc &desc = computePassDescs_[task.index];
          if (gpuProfiler) gpuProfiler->StartTask(stream, desc.profilerTaskName, desc.profilerTaskColor);
          const size_t cpuTask = cpuProfiler ? cpuProfiler->StartTask(desc.profilerTaskName, desc.profilerTaskColor) : 0;
          PassContext ctx;
          InitContext(ctx, stream);
          for (auto id : desc.inputImageViewProxies) Bind(ctx, id);
          for (auto id : desc.inoutStorageImageProxies) Bind(ctx, id);
          for (auto id : desc.inoutStorageBufferProxies) Bind(ctx, id);
          if (desc.recordFunc) desc.recordFunc(ctx);
          if (cpuProfiler) cpuProfiler->EndTask(cpuTask);
        } break;
        case Task::Type::TransferPass: {
          TransferPassDesc &desc = transferPassDescs_[task.index];
          if (gpuProfiler) gpuProfiler->StartTask(stream, desc.profilerTaskName, desc.profilerTaskColor);
          const size_t cpuTask = cpuProfiler ? cpuProfiler->StartTask(desc.profilerTaskName, desc.profilerTaskColor) : 0;
          PassContext ctx;
          InitContext(ctx, stream);
          for (auto id : desc.srcImageViewProxies) Bind(ctx, id);
          for (auto id : desc.dstImageViewProxies) Bind(ctx, id);
          for (auto id : desc.srcBufferProxies) Bind(ctx, id);
          for (auto id : desc.dstBufferProxies) Bind(ctx, id);
          if (desc.recordFunc) desc.recordFunc(ctx);
          if (cpuProfiler) cpuProfiler->EndTask(cpuTask);
        } break;
        case Task::Type::ImagePresent:   // layout transition to PresentSrc in the reference; nothing to do on a stream
        case Task::Type::FrameSyncBegin: // external-image barriers in the reference
        case Task::Type::FrameSyncEnd:
          break;
      }
    }
    if (gpuProfiler) gpuProfiler->EndFrame(stream);
    ClearPasses(); // includes the compute pass list, which the reference forgets to clear
  }

  // Resolved view of a proxy after the last Execute (harness / tests read images back through this).
  ImageView *GetResolvedImageView(ImageViewProxyId id) { return imageViewProxies_.Get(id).resolved; }

  uint64_t GetAllocatedBytes() const {
    uint64_t total = 0;
    for (const auto &entry : imageCache_)
      for (const auto &img : entry.second.images) total += img->GetByteSize();
    for (const auto &entry : bufferCache_)
      for (const auto &buf : entry.second.buffers) total += buf->GetSize();
    return total;
  }

private:
  // -- proxies
  struct ImageKey {
    vk::Format format = vk::Format::eUndefined;
    vk::ImageUsageFlags usageFlags;
    uint32_t mipsCount = 0, arrayLayersCount = 0;
    glm::uvec3 size;
    bool operator<(const ImageKey &o) const {
      return std::tie(format, mipsCount, arrayLayersCount, usageFlags.mask, size.x, size.y, size.z) <
             std::tie(o.format, o.mipsCount, o.arrayLayersCount, o.usageFlags.mask, o.size.x, o.size.y, o.size.z);
    }
  };
  struct ImageProxy {
    ImageKey key;
    ImageData *external = nullptr;
    ImageData *resolved = nullptr;
    std::string debugName;
  };
  struct ImageViewProxy {
    ImageProxyId imageProxyId;
    uint32_t baseMipLevel = 0, mipLevelsCount = 0, baseArrayLayer = 0, arrayLayersCount = 0;
    ImageView *external = nullptr;
    ImageUsageTypes externalUsageType = ImageUsageTypes::Unknown;
    ImageView *resolved = nullptr;
    std::string debugName;
  };
  struct BufferProxy {
    uint32_t elementSize = 0, elementsCount = 0;
    Buffer *external = nullptr;
    Buffer *resolved = nullptr;
  };

  // -- per-frame pooled allocations (ImageCache / ImageViewCache / BufferCache of the reference)
  struct ImageCacheEntry {
    std::vector<std::unique_ptr<ImageData>> images;
    size_t usedCount = 0;
  };
  struct BufferCacheEntry {
    std::vector<std::unique_ptr<Buffer>> buffers;
    size_t usedCount = 0;
  };
  using ViewKey = std::tuple<ImageData *, uint32_t, uint32_t>;

  void ResolveImages() {
    for (auto &entry : imageCache_) entry.second.usedCount = 0;
    imageProxies_.ForEach([&](ImageProxyId, ImageProxy &proxy) {
      if (proxy.external) {
        proxy.resolved = proxy.external;
        return;
      }
      ImageCacheEntry &entry = imageCache_[proxy.key];
      if (entry.usedCount == entry.images.size())
        entry.images.emplace_back(new ImageData(proxy.key.format, glm::uvec2(proxy.key.size.x, proxy.key.size.y), proxy.key.mipsCount));
      proxy.resolved = entry.images[entry.usedCount++].get();
    });
  }
  void ResolveImageViews() {
    imageViewProxies_.ForEach([&](ImageViewProxyId, ImageViewProxy &proxy) {
      if (proxy.external) {
        proxy.resolved = proxy.external;
        return;
      }
      ImageData *image = imageProxies_.Get(proxy.imageProxyId).resolved;
      std::unique_ptr<ImageView> &view = imageViewCache_[ViewKey(image, proxy.baseMipLevel, proxy.mipLevelsCount)];
      if (!view) view.reset(new ImageView(image, proxy.baseMipLevel, proxy.mipLevelsCount));
      proxy.resolved = view.get();
    });
  }
  void ResolveBuffers() {
    for (auto &entry : bufferCache_) entry.second.usedCount = 0;
    bufferProxies_.ForEach([&](BufferProxyId, BufferProxy &proxy) {
      if (proxy.external) {
        proxy.resolved = proxy.external;
        return;
      }
      BufferCacheEntry &entry = bufferCache_[std::make_pair(proxy.elementSize, proxy.elementsCount)];
      if (entry.usedCount == entry.buffers.size()) entry.buffers.emplace_back(new Buffer(size_t(proxy.elementSize) * proxy.elementsCount));
      proxy.resolved = entry.buffers[entry.usedCount++].get();
    });
  }

  void InitContext(PassContext &ctx, cudaStream_t stream) {
    ctx.resolvedImageViews.assign(imageViewProxies_.GetSize(), nullptr);
    ctx.resolvedBuffers.assign(bufferProxies_.GetSize(), nullptr);
    ctx.stream = stream;
  }
  void Bind(PassContext &ctx, ImageViewProxyId id) { ctx.resolvedImageViews[id.asInt] = imageViewProxies_.Get(id).resolved; }
  void Bind(PassContext &ctx, BufferProxyId id) { ctx.resolvedBuffers[id.asInt] = bufferProxies_.Get(id).resolved; }

  void ClearPasses() {
    renderPassDescs_.clear();
    computePassDescs_.clear();
    transferPassDescs_.clear();
    imagePresentDescs_.clear();
    tasks_.clear();
  }

  struct Task {
    enum struct Type { RenderPass, ComputePass, TransferPass, ImagePresent, FrameSyncBegin, FrameSyncEnd };
    Type type;
    size_t index;
  };

  Utils::Pool<ImageProxy> imageProxies_;
  Utils::Pool<ImageViewProxy> imageViewProxies_;
  Utils::Pool<BufferProxy> bufferProxies_;
  std::map<ImageKey, ImageCacheEntry> imageCache_;
  std::map<ViewKey, std::unique_ptr<ImageView>> imageViewCache_;
  std::map<std::pair<uint32_t, uint32_t>, BufferCacheEntry> bufferCache_;
  std::vector<Task> tasks_;
  std::vector<RenderPassDesc> renderPassDescs_;
  std::vector<ComputePassDesc> computePassDescs_;
  std::vector<TransferPassDesc> transferPassDescs_;
  std::vector<ImagePresentPassDesc> imagePresentDescs_;
};

inline vk::AttachmentLoadOp RenderGraph::RenderPassContext::GetColorLoadOp(size_t index) const { return desc->colorAttachments[index].loadOp; }
inline vk::ClearValue RenderGraph::RenderPassContext::GetColorClearValue(size_t index) const { return desc->colorAttachments[index].clearValue; }
inline vk::AttachmentLoadOp RenderGraph::RenderPassContext::GetDepthLoadOp() const { return desc->depthAttachment.loadOp; }
inline vk::ClearValue RenderGraph::RenderPassContext::GetDepthClearValue() const { return desc->depthAttachment.clearValue; }
inline vk::Extent2D RenderGraph::RenderPassContext::GetRenderAreaExtent() const { return desc->renderAreaExtent; }

// Owner of the graph and the stream the frame runs on; stands where legit::Core stands in renderer constructors
// (LV/Core.h: GetRenderGraph()). Device / queue / pipeline-cache management has no CUDA analogue.
class Core {
public:
  explicit Core(cudaStream_t stream = nullptr) : stream_(stream) {}
  RenderGraph *GetRenderGraph() { return &renderGraph_; }
  cudaStream_t GetStream() const { return stream_; }
  void SetStream(cudaStream_t s) { stream_ = s; }
  void WaitIdle() { CudaCheck(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"); }

private:
  RenderGraph renderGraph_;
  cudaStream_t stream_;
};

} // namespace legit_cuda
