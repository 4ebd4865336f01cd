// BlurBuilder.h — one "BlurPass" per call, CUDA-backed. Mirrors src/Render/Common/BlurBuilder.h:14-46:
// size = GetMipSize(src, 0) (must equal dst), UBO {ivec4 size; int radius}, render area = that size; the record lambda
// calls lgcu_blur_level where the reference binds blurLayerBuilder.frag and draws a full-screen quad.
#pragma once

#include <cassert>

#include "RenderGraph.h"
#include "ShaderMemoryPool.h"

namespace legit_cuda {

class BlurBuilder {
public:
  explicit BlurBuilder(Core *_core) : core(_core), imageSpaceSampler(SamplerAddressMode::eClampToEdge, Filter::eNearest, SamplerMipmapMode::eNearest) {}

  void ApplyBlur(RenderGraph *renderGraph, ShaderMemoryPool *memoryPool, RenderGraph::ImageViewProxyId srcProxyId, RenderGraph::ImageViewProxyId dstProxyId,
                 int radius) {
    glm::uvec2 viewportSize = renderGraph->GetMipSize(srcProxyId, 0);
    assert(viewportSize == renderGraph->GetMipSize(dstProxyId, 0));
    if (viewportSize.x == 0 || viewportSize.y == 0) return; // the reference would open a zero-area render pass here (only below 512 px)
    vk::Extent2D layerSize(viewportSize.x, viewportSize.y);
    renderGraph->AddPass(RenderGraph::RenderPassDesc()
                             .SetColorAttachments({dstProxyId})
                             .SetInputImages({srcProxyId})
                             .SetRenderAreaExtent(layerSize)
                             .SetProfilerInfo(Colors::wisteria, "BlurPass")
                             .SetRecordFunc([memoryPool, srcProxyId, viewportSize, radius](RenderGraph::RenderPassContext passContext) {
                               memoryPool->BeginSet();
                               auto shaderDataBuffer = memoryPool->GetUniformBufferData<lgcu_blur_layer_builder_data>("BlurLayerBuilderData");
                               shaderDataBuffer->size[0] = int32_t(viewportSize.x);
                               shaderDataBuffer->size[1] = int32_t(viewportSize.y);
                               shaderDataBuffer->size[2] = shaderDataBuffer->size[3] = 0;
                               shaderDataBuffer->radius = radius;
                               memoryPool->EndSet();
                               LgcuCheck(lgcu_blur_level(shaderDataBuffer, passContext.GetImageView(srcProxyId)->GetDesc(), // "srcSampler"
                                                         passContext.GetColorAttachment(0)->GetDesc(), nullptr, passContext.GetStream()),
                                         "BlurPass");
                             }));
  }

  void ReloadShaders() {}

private:
  Core *core;
  Sampler imageSpaceSampler;
};

} // namespace legit_cuda
