// InterleaveBuilder.h — interleaved-rendering passes, CUDA-backed. Mirrors src/Render/Common/InterleaveBuilder.h:14-80:
// Deinterleave(graph, pool, interleavedId, deinterleavedId, gridSize) and Interleave(graph, pool, deinterleavedId, interleavedId,
// gridSize) each add one full-screen pass whose render area is the size of the interleaved view (:16, :49) and whose UBO is
// {ivec4 gridSize; ivec4 viewportSize} (:99-123). The record lambdas call lgcu_deinterleave / lgcu_interleave where the
// reference binds deinterleave.frag / interleave.frag and draws a quad.
//
// One deliberate difference: the reference's Interleave() binds the DE-interleave program and then asks it for a uniform block
// named "InterleaveData", which that program does not have (InterleaveBuilder.h:60-65) — its only callers are renderers that are
// not compiled in. This mirror runs the shipped interleave.frag, i.e. what the function means to do.
#pragma once

#include "RenderGraph.h"
#include "ShaderMemoryPool.h"

namespace legit_cuda {

class InterleaveBuilder {
public:
  explicit InterleaveBuilder(Core *_core) : core(_core), imageSpaceSampler(SamplerAddressMode::eClampToEdge, Filter::eNearest, SamplerMipmapMode::eNearest) {}

  void Deinterleave(RenderGraph *renderGraph, ShaderMemoryPool *memoryPool, RenderGraph::ImageViewProxyId interleavedProxyId,
                    RenderGraph::ImageViewProxyId deinterleavedProxyId, glm::uvec2 gridSize) {
    auto viewportSize = renderGraph->GetMipSize(interleavedProxyId, 0);
    vk::Extent2D viewportExtent(viewportSize.x, viewportSize.y);
    renderGraph->AddPass(RenderGraph::RenderPassDesc()
                             .SetColorAttachments({deinterleavedProxyId}, vk::AttachmentLoadOp::eDontCare)
                             .SetInputImages({interleavedProxyId})
                             .SetRenderAreaExtent(viewportExtent)
                             .SetProfilerInfo(Colors::turqoise, "DeinterleavePass")
                             .SetRecordFunc([memoryPool, interleavedProxyId, gridSize, viewportSize](RenderGraph::RenderPassContext passContext) {
                               auto shaderDataBuffer = Fill(memoryPool, "DeinterleaveData", gridSize, viewportSize);
                               LgcuCheck(lgcu_deinterleave(shaderDataBuffer, passContext.GetImageView(interleavedProxyId)->GetDesc(), // "interleavedSampler"
                                                           passContext.GetColorAttachment(0)->GetDesc(), nullptr, passContext.GetStream()),
                                         "DeinterleavePass");
                             }));
  }

  void Interleave(RenderGraph *renderGraph, ShaderMemoryPool *memoryPool, RenderGraph::ImageViewProxyId deinterleavedProxyId,
                  RenderGraph::ImageViewProxyId interleavedProxyId, glm::uvec2 gridSize) {
    auto viewportSize = renderGraph->GetMipSize(interleavedProxyId, 0);
    vk::Extent2D viewportExtent(viewportSize.x, viewportSize.y);
    renderGraph->AddPass(RenderGraph::RenderPassDesc()
                             .SetColorAttachments({interleavedProxyId}, vk::AttachmentLoadOp::eDontCare)
                             .SetInputImages({deinterleavedProxyId})
                             .SetRenderAreaExtent(viewportExtent)
                             .SetProfilerInfo(Colors::turqoise, "InterleavePass")
                             .SetRecordFunc([memoryPool, deinterleavedProxyId, gridSize, viewportSize](RenderGraph::RenderPassContext passContext) {
                               auto shaderDataBuffer = Fill(memoryPool, "InterleaveData", gridSize, viewportSize);
                               LgcuCheck(lgcu_interleave(shaderDataBuffer, passContext.GetImageView(deinterleavedProxyId)->GetDesc(), // "deinterleavedSampler"
                                                         passContext.GetColorAttachment(0)->GetDesc(), nullptr, passContext.GetStream()),
                                         "InterleavePass");
                             }));
  }

  void ReloadShaders() {}

private:
  static lgcu_interleave_data *Fill(ShaderMemoryPool *memoryPool, const char *name, glm::uvec2 gridSize, glm::uvec2 viewportSize) {
    memoryPool->BeginSet();
    auto shaderDataBuffer = memoryPool->GetUniformBufferData<lgcu_interleave_data>(name);
    shaderDataBuffer->gridSize[0] = int32_t(gridSize.x); // glm::ivec4(gridSize, 0.0f, 0.0f)
    shaderDataBuffer->gridSize[1] = int32_t(gridSize.y);
    shaderDataBuffer->gridSize[2] = shaderDataBuffer->gridSize[3] = 0;
    shaderDataBuffer->viewportSize[0] = int32_t(viewportSize.x);
    shaderDataBuffer->viewportSize[1] = int32_t(viewportSize.y);
    shaderDataBuffer->viewportSize[2] = shaderDataBuffer->viewportSize[3] = 0;
    memoryPool->EndSet();
    return shaderDataBuffer;
  }
  Core *core;
  Sampler imageSpaceSampler;
};

} // namespace legit_cuda
