// MipBuilder.h — image-proxy wrappers and the mip-chain pass list, CUDA-backed.
// Mirrors src/Render/Common/MipBuilder.h: UnmippedProxy (:3-14), MippedProxy (:16-33, 10 levels, one whole-chain view
// plus one view per level) and MipBuilder::BuildMips for 2-D images (:142-181: one "MipBuilderPass" per level, render
// area = destination level, loop ends when a dimension reaches 0). The record lambda calls lgcu_mip_level where the
// reference binds mipLevelBuilder.frag and draws a full-screen quad. The 3-D (compute) overload and FilterTypes::Depth
// have no live caller in the reference and are not provided.
#pragma once

#include "RenderGraph.h"
#include "ShaderMemoryPool.h"

namespace legit_cuda {

struct UnmippedProxy {
  UnmippedProxy(RenderGraph *renderGraph, vk::Format format, glm::uvec2 _baseSize, vk::ImageUsageFlags usageFlags) : baseSize(_baseSize) {
    imageProxy = renderGraph->AddImage(format, 1, 1, _baseSize, usageFlags);
    imageViewProxy = renderGraph->AddImageView(imageProxy->Id(), 0, 1, 0, 1);
  }
  RenderGraph::ImageProxyUnique imageProxy;
  RenderGraph::ImageViewProxyUnique imageViewProxy;
  glm::uvec2 baseSize;
};

struct MippedProxy {
  MippedProxy(RenderGraph *renderGraph, vk::Format format, glm::uvec2 _baseSize, vk::ImageUsageFlags usageFlags) : baseSize(_baseSize) {
    const uint32_t mipsCount = 10;
    imageProxy = renderGraph->AddImage(format, mipsCount, 1, _baseSize, usageFlags);
    imageViewProxy = renderGraph->AddImageView(imageProxy->Id(), 0, mipsCount, 0, 1);
    for (uint32_t mipIndex = 0; mipIndex < mipsCount; mipIndex++) mipImageViewProxies.push_back(renderGraph->AddImageView(imageProxy->Id(), mipIndex, 1, 0, 1));
  }
  RenderGraph::ImageProxyUnique imageProxy;
  RenderGraph::ImageViewProxyUnique imageViewProxy;
  std::vector<RenderGraph::ImageViewProxyUnique> mipImageViewProxies;
  glm::uvec2 baseSize;
};

class MipBuilder {
public:
  explicit MipBuilder(Core *_core) : core(_core), imageSpaceSampler(SamplerAddressMode::eClampToEdge, Filter::eNearest, SamplerMipmapMode::eNearest) {}
  enum struct FilterTypes { Avg, Depth };

  void BuildMips(RenderGraph *renderGraph, ShaderMemoryPool *memoryPool, const MippedProxy &mippedProxy, FilterTypes filterType = FilterTypes::Avg) {
    vk::Extent2D layerSize(mippedProxy.baseSize.x, mippedProxy.baseSize.y);
    for (size_t mipIndex = 1; mipIndex < mippedProxy.mipImageViewProxies.size(); mipIndex++) {
      layerSize.width /= 2;
      layerSize.height /= 2;
      if (layerSize.width <= 0 || layerSize.height <= 0) break;
      auto srcProxyId = mippedProxy.mipImageViewProxies[mipIndex - 1]->Id();
      auto dstProxyId = mippedProxy.mipImageViewProxies[mipIndex]->Id();
      renderGraph->AddPass(RenderGraph::RenderPassDesc()
                               .SetColorAttachments({dstProxyId})
                               .SetInputImages({srcProxyId})
                               .SetRenderAreaExtent(layerSize)
                               .SetProfilerInfo(Colors::nephritis, "MipBuilderPass")
                               .SetRecordFunc([memoryPool, srcProxyId, filterType](RenderGraph::RenderPassContext passContext) {
                                 memoryPool->BeginSet();
                                 auto shaderDataBuffer = memoryPool->GetUniformBufferData<lgcu_mip_level_builder_data>("MipLevelBuilderData");
                                 shaderDataBuffer->filterType = (filterType == FilterTypes::Avg) ? 0.0f : 1.0f;
                                 memoryPool->EndSet();
                                 auto prevMipView = passContext.GetImageView(srcProxyId); // "prevLevelSampler"
                                 LgcuCheck(lgcu_mip_level(shaderDataBuffer, prevMipView->GetDesc(), passContext.GetColorAttachment(0)->GetDesc(), nullptr,
                                                          passContext.GetStream()),
                                           "MipBuilderPass");
                               }));
    }
  }

  void ReloadShaders() {} // kernels are linked in; nothing to load

private:
  Core *core;
  Sampler imageSpaceSampler;
};

} // namespace legit_cuda
