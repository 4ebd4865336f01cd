// DebugRenderer.h — "DebugInfoPass", CUDA-backed. Mirrors src/Render/Common/DebugRenderer.h:13-63: one pass over the target
// (loadOp eLoad) that draws every debug view as a 0.1 x 0.1 tile with 0.02 padding, left to right, wrapping to the next row when
// the next tile would cross the right edge (:27-57). Per tile the reference fills QuadData.minmax and draws a 4-vertex quad with
// debugRenderer.vert / debugRenderer.frag; here the same loop issues one lgcu_debug_overlay call per tile.
#pragma once

#include <vector>

#include "RenderGraph.h"
#include "ShaderMemoryPool.h"

namespace legit_cuda {

class DebugRenderer {
public:
  explicit DebugRenderer(Core *_core) : core(_core), imageSpaceSampler(SamplerAddressMode::eClampToEdge, Filter::eLinear, SamplerMipmapMode::eNearest) {}

  void RenderImageViews(RenderGraph *renderGraph, ShaderMemoryPool *memoryPool, RenderGraph::ImageViewProxyId targetProxyId,
                        std::vector<RenderGraph::ImageViewProxyId> debugProxies, const lgcu_rows *rows = nullptr) {
    glm::uvec2 viewportSize = renderGraph->GetMipSize(targetProxyId, 0);
    vk::Extent2D viewportExtent(viewportSize.x, viewportSize.y);
    const bool useRows = rows != nullptr;
    const lgcu_rows rowsCopy = rows ? *rows : lgcu_rows{0, 0};
    renderGraph->AddPass(RenderGraph::RenderPassDesc()
                             .SetColorAttachments({targetProxyId}, vk::AttachmentLoadOp::eLoad)
                             .SetInputImages(std::vector<RenderGraph::ImageViewProxyId>(debugProxies))
                             .SetRenderAreaExtent(viewportExtent)
                             .SetProfilerInfo(Colors::clouds, "DebugInfoPass")
                             .SetRecordFunc([memoryPool, debugProxies, useRows, rowsCopy](RenderGraph::RenderPassContext passContext) {
                               const float tileSize[2] = {0.1f, 0.1f}, tilePadding[2] = {0.02f, 0.02f};
                               float currMin[2] = {tilePadding[0], tilePadding[1]};
                               for (auto debugProxyId : debugProxies) {
                                 memoryPool->BeginSet();
                                 auto quadDataBuffer = memoryPool->GetUniformBufferData<lgcu_debug_quad_data>("QuadData");
                                 quadDataBuffer->minmax[0] = currMin[0];
                                 quadDataBuffer->minmax[1] = currMin[1];
                                 quadDataBuffer->minmax[2] = currMin[0] + tileSize[0];
                                 quadDataBuffer->minmax[3] = currMin[1] + tileSize[1];
                                 memoryPool->EndSet();
                                 LgcuCheck(lgcu_debug_overlay(quadDataBuffer, passContext.GetImageView(debugProxyId)->GetDesc(), // "srcSampler"
                                                              passContext.GetColorAttachment(0)->GetDesc(), useRows ? &rowsCopy : nullptr, passContext.GetStream()),
                                           "DebugInfoPass");
                                 currMin[0] += tileSize[0] + tilePadding[0];
                                 if (currMin[0] + tileSize[0] > 1.0f) {
                                   currMin[0] = tilePadding[0];
                                   currMin[1] += tileSize[1] + tilePadding[1];
                                 }
                               }
                             }));
  }

  void ReloadShaders() {}

private:
  Core *core;
  Sampler imageSpaceSampler;
};

} // namespace legit_cuda
