// SSVGIRenderer.h — the reference's SSVGI renderer (src/Render/Renderers/SSVGIRenderer.h) on legit_cuda::RenderGraph.
//
// Same screen images (ViewportResources, :391-419), same UBO structs (include/lgcu.h mirrors :33-38, :422-519), same
// pass list in the same AddPass order with the same attachments / inputs / render areas / profiler names and colours
// (:63-342). Each record lambda fills its parameter block through the ShaderMemoryPool like the reference and then
// calls one lgcu_* entry point where the reference binds a pipeline + descriptor set and draws a full-screen quad.
//
// What differs, and why:
//  * Scene. The reference rasterises meshes (Scene::IterateObjects, :84-102, :138-156) with the fixed-function rasteriser. Here
//    a Scene comes in one of two forms:
//      - mesh (Scene::SetMesh): vertex / index buffers + draw list + per-object constants, like the reference's Scene.
//        "ShadowPass" rasterises the light's depth map (lgcu_raster_shadow_map) and "GBufferRasterPass" writes the per-pixel
//        fragment buffer (lgcu_raster_gbuffer) that the fragment stage then reads;
//      - already rasterised (§8d's synthetic input): the fragment buffer, the per-draw-call constants and the light's depth map
//        arrive from outside; "ShadowPass" copies that depth map into the shadowMap image.
//    In both forms "GBufferPass" runs the fragment stage (lgcu_gbuffer_resolve) over the fragment buffer.
//  * FrameOptions::mode == Fused replaces groups of passes by the fused entry points (frame front = K1+K2+level-0 blur+mips 1..4,
//    frame chains = remaining blur/mip work, packed GI gather, K6+K7): identical images, 6 launches instead of 46 passes.
//    PassGranular is the 1:1 pass list (47 passes incl. shadow).
//  * The debug overlay (DebugRenderer, :344-350) is drawn when FrameOptions::debugOverlay is set (the reference always draws it; the
//    parity tests and the bench compare the frame before it).
#pragma once

#include <memory>

#include "BlurBuilder.h"
#include "Camera.h"
#include "DebugRenderer.h"
#include "MipBuilder.h"

namespace legit_cuda {

// Mirrors legit::InFlightQueue::FrameInfo (LV/PresentQueue.h:71-80): what BeginFrame hands to RenderFrame.
struct FrameInfo {
  ShaderMemoryPool *memoryPool = nullptr;
  RenderGraph::ImageViewProxyId swapchainImageViewProxyId;
};

// The rasterised scene (see header comment). Buffers are device memory registered as external buffers.
struct Scene {
  Scene(RenderGraph *graph, Buffer *fragments_, uint64_t fragmentPitch_, Buffer *objects_, uint32_t objectsCount_, Buffer *lightDepth_, uint32_t lightDepthSize_)
      : fragments(fragments_), objects(objects_), lightDepth(lightDepth_), fragmentPitch(fragmentPitch_), objectsCount(objectsCount_), lightDepthSize(lightDepthSize_) {
    fragmentsProxy = graph->AddExternalBuffer(fragments);
    objectsProxy = graph->AddExternalBuffer(objects);
    lightDepthProxy = graph->AddExternalBuffer(lightDepth);
  }
  // Mesh form: the buffers of desc are device memory owned by the caller; scratch must hold lgcu_raster_scratch_bytes for both the
  // viewport and the shadow map. The draw-call constants the fragment stage reads are desc.objects.
  void SetMesh(RenderGraph *graph, const lgcu_mesh_scene &desc, Buffer *rasterScratch_) {
    mesh = desc;
    if (rasterScratch != rasterScratch_ || !rasterScratchProxy.IsAttached()) {
      rasterScratch = rasterScratch_;
      rasterScratchProxy = graph->AddExternalBuffer(rasterScratch);
    }
    hasMesh = true;
  }
  void ClearMesh() { hasMesh = false; }
  const lgcu_draw_call_data *DrawCallData(Buffer *resolvedObjects) const {
    return hasMesh ? mesh.objects : static_cast<const lgcu_draw_call_data *>(resolvedObjects->GetHandle());
  }
  uint32_t DrawCallCount() const { return hasMesh ? mesh.nObjects : objectsCount; }
  Buffer *fragments, *objects, *lightDepth;
  uint64_t fragmentPitch;
  uint32_t objectsCount, lightDepthSize;
  RenderGraph::BufferProxyUnique fragmentsProxy, objectsProxy, lightDepthProxy;
  bool hasMesh = false;
  lgcu_mesh_scene mesh{};
  Buffer *rasterScratch = nullptr;
  RenderGraph::BufferProxyUnique rasterScratchProxy;
};

struct FrameOptions {
  enum struct Mode { PassGranular, Fused };
  Mode mode = Mode::Fused;
  int denoiserRadius = 0;               // key 'A' held -> 2 in the reference (:288)
  uint32_t giFlags = LGCU_GI_DEFAULT;   // LGCU_GI_STRICT selects the shader-order parity kernel
  bool useRows = false;                 // multi-GPU strip: only rows [rows.y0, rows.y1) of the frame are produced
  lgcu_rows rows{0, 0};
  // Multi-GPU strips run the fused frame in stages with a halo exchange between them; a single GPU runs all.
  enum Stage : uint32_t { StageFront = 1, StageChains = 2, StageGather = 4, StageFinal = 8, StageAll = 15 };
  uint32_t stages = StageAll;
  bool debugOverlay = false; // DebugInfoPass: thumbnails of normal / albedo / indirectLight / denoisedIndirectLight over the frame (:344-350)
};

class SSVGIRenderer {
public:
  explicit SSVGIRenderer(Core *_core)
      : mipBuilder(_core), blurBuilder(_core), debugRenderer(_core), screenspaceSampler(SamplerAddressMode::eClampToEdge, Filter::eLinear, SamplerMipmapMode::eLinear),
        shadowmapSampler(SamplerAddressMode::eClampToEdge, Filter::eLinear, SamplerMipmapMode::eNearest, true), core(_core) {}

  void RecreateSceneResources(Scene *) {}
  void RecreateSwapchainResources(vk::Extent2D _viewportExtent, size_t /*inFlightFramesCount*/) {
    viewportExtent = _viewportExtent;
    viewportResources.reset(new ViewportResources(core->GetRenderGraph(), glm::uvec2(viewportExtent.width, viewportExtent.height)));
  }

  void RenderFrame(const FrameInfo &frameInfo, const Camera &camera, const Camera &light, Scene *scene, const FrameOptions &options = FrameOptions()) {
    struct PassData {
      ShaderMemoryPool *memoryPool;
      lgcu_mat4 viewMatrix, projMatrix, lightViewMatrix, lightProjMatrix;
      Scene *scene;
      const lgcu_rows *rows;
      lgcu_rows rowsStorage;
    } passData;
    passData.memoryPool = frameInfo.memoryPool;
    passData.scene = scene;
    const FrameMatrices frame = MakeFrameMatrices(camera, light, viewportExtent.width, viewportExtent.height); // :54-59
    passData.viewMatrix = frame.viewMatrix;
    passData.projMatrix = frame.projMatrix;
    passData.lightViewMatrix = frame.lightViewMatrix;
    passData.lightProjMatrix = frame.lightProjMatrix;
    passData.rowsStorage = options.rows;
    RenderGraph *graph = core->GetRenderGraph();
    ViewportResources *res = viewportResources.get();
    const bool useRows = options.useRows;
    auto rowsOf = [useRows](const PassData &pd) -> const lgcu_rows * { return useRows ? &pd.rowsStorage : nullptr; };

    const uint32_t stages = options.mode == FrameOptions::Mode::Fused ? options.stages : uint32_t(FrameOptions::StageAll);
    // rendering shadow map (:61-104)
    vk::Extent2D shadowMapExtent(res->shadowMap.baseSize.x, res->shadowMap.baseSize.y);
    if ((stages & FrameOptions::StageFront) && scene->hasMesh)
      graph->AddPass(RenderGraph::RenderPassDesc()
                       .SetDepthAttachment(res->shadowMap.imageViewProxy->Id(), vk::AttachmentLoadOp::eClear)
                       .SetStorageBuffers({scene->rasterScratchProxy->Id()})
                       .SetRenderAreaExtent(shadowMapExtent)
                       .SetProfilerInfo(Colors::amethyst, "ShadowPass")
                       .SetRecordFunc([passData](RenderGraph::RenderPassContext passContext) {
                         passData.memoryPool->BeginSet();
                         auto shaderDataBuffer = passData.memoryPool->GetUniformBufferData<lgcu_shadowmap_builder_data>("ShadowmapBuilderData");
                         shaderDataBuffer->lightViewMatrix = passData.lightViewMatrix; // :77-78
                         shaderDataBuffer->lightProjMatrix = passData.lightProjMatrix;
                         passData.memoryPool->EndSet();
                         Buffer *scratch = passContext.GetBuffer(passData.scene->rasterScratchProxy->Id());
                         // Scene::IterateObjects + drawIndexed per object (:84-102) = the draw list of the mesh scene
                         LgcuCheck(lgcu_raster_shadow_map(shaderDataBuffer, &passData.scene->mesh, scratch->GetHandle(), scratch->GetSize(),
                                                          passContext.GetDepthAttachment()->GetDesc(), passContext.GetStream()),
                                   "ShadowPass");
                       }));
    else if (stages & FrameOptions::StageFront) // the light's depth arrives rasterised with the scene
      graph->AddPass(RenderGraph::RenderPassDesc()
                       .SetDepthAttachment(res->shadowMap.imageViewProxy->Id(), vk::AttachmentLoadOp::eClear)
                       .SetStorageBuffers({scene->lightDepthProxy->Id()})
                       .SetRenderAreaExtent(shadowMapExtent)
                       .SetProfilerInfo(Colors::amethyst, "ShadowPass")
                       .SetRecordFunc([passData](RenderGraph::RenderPassContext passContext) {
                         ImageView *shadowMap = passContext.GetDepthAttachment();
                         const uint32_t size = passData.scene->lightDepthSize;
                         if (shadowMap->GetImageData()->GetMipSize(0).x != size) throw std::runtime_error("ShadowPass: light depth size != shadow map size");
                         CudaCheck(cudaMemcpy2DAsync(shadowMap->GetImageData()->GetLevelPointer(0), shadowMap->GetImageData()->GetLevelPitch(0),
                                                     passContext.GetBuffer(passData.scene->lightDepthProxy->Id())->GetHandle(), size_t(size) * 4, size_t(size) * 4, size,
                                                     cudaMemcpyDeviceToDevice, passContext.GetStream()),
                                   "ShadowPass copy");
                       }));

    auto fillGBufferData = [](const PassData &pd) {
      pd.memoryPool->BeginSet();
      auto shaderDataBuffer = pd.memoryPool->GetUniformBufferData<lgcu_gbuffer_builder_data>("GBufferBuilderData");
      shaderDataBuffer->time = 0.0f;
      shaderDataBuffer->projMatrix = pd.projMatrix;
      shaderDataBuffer->viewMatrix = pd.viewMatrix;
      pd.memoryPool->EndSet();
      return shaderDataBuffer;
    };
    auto fillLightData = [](const PassData &pd) {
      pd.memoryPool->BeginSet();
      auto shaderDataBuffer = pd.memoryPool->GetUniformBufferData<lgcu_direct_lighting_data>("DirectLightingData");
      shaderDataBuffer->viewMatrix = pd.viewMatrix;
      shaderDataBuffer->projMatrix = pd.projMatrix;
      shaderDataBuffer->lightViewMatrix = pd.lightViewMatrix;
      shaderDataBuffer->lightProjMatrix = pd.lightProjMatrix;
      shaderDataBuffer->time = 0.0f;
      pd.memoryPool->EndSet();
      return shaderDataBuffer;
    };
    auto clearOf = [](RenderGraph::RenderPassContext &ctx) {
      lgcu_clear_values clear;
      const vk::ClearValue color = ctx.GetColorClearValue(0);
      for (int i = 0; i < 4; i++) clear.color[i] = color.color.float32[size_t(i)];
      clear.depth = ctx.GetDepthClearValue().depthStencil.depth;
      return clear;
    };

    // rendering gbuffer (:106-158), raster half: vertex stage + rasteriser + depth test -> per-pixel fragment buffer
    if (scene->hasMesh && (stages & FrameOptions::StageFront)) {
      const uint32_t vw = viewportExtent.width, vh = viewportExtent.height;
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetStorageBuffers({scene->fragmentsProxy->Id(), scene->rasterScratchProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::belizeHole, "GBufferRasterPass")
                         .SetRecordFunc([passData, fillGBufferData, rowsOf, vw, vh](RenderGraph::RenderPassContext passContext) {
                           auto shaderDataBuffer = fillGBufferData(passData);
                           Scene *sc = passData.scene;
                           Buffer *scratch = passContext.GetBuffer(sc->rasterScratchProxy->Id());
                           LgcuCheck(lgcu_raster_gbuffer(shaderDataBuffer, &sc->mesh, scratch->GetHandle(), scratch->GetSize(), vw, vh,
                                                         static_cast<lgcu_fragment *>(passContext.GetBuffer(sc->fragmentsProxy->Id())->GetHandle()), sc->fragmentPitch,
                                                         rowsOf(passData), passContext.GetStream()),
                                     "GBufferRasterPass");
                         }));
    }

    if (options.mode == FrameOptions::Mode::PassGranular) {
      // rendering gbuffer (:106-158): fragment stage over the rasterised fragments
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetColorAttachments({res->albedo.imageViewProxy->Id(),                  // location = 0
                                               res->emissive.imageViewProxy->Id(),                // location = 1
                                               res->normal.imageViewProxy->Id(),                  // location = 2
                                               res->depthMoments.mipImageViewProxies[0]->Id()},   // location = 3
                                              vk::AttachmentLoadOp::eClear)
                         .SetDepthAttachment(res->depthStencil.imageViewProxy->Id(), vk::AttachmentLoadOp::eClear)
                         .SetStorageBuffers({scene->fragmentsProxy->Id(), scene->objectsProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::belizeHole, "GBufferPass")
                         .SetRecordFunc([passData, fillGBufferData, clearOf, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto shaderDataBuffer = fillGBufferData(passData);
                           const lgcu_clear_values clear = clearOf(passContext);
                           Scene *sc = passData.scene;
                           LgcuCheck(lgcu_gbuffer_resolve(shaderDataBuffer, sc->DrawCallData(passContext.GetBuffer(sc->objectsProxy->Id())),
                                                          sc->DrawCallCount(), static_cast<const lgcu_fragment *>(passContext.GetBuffer(sc->fragmentsProxy->Id())->GetHandle()),
                                                          sc->fragmentPitch, &clear, passContext.GetColorAttachment(0)->GetDesc(), passContext.GetColorAttachment(1)->GetDesc(),
                                                          passContext.GetColorAttachment(2)->GetDesc(), passContext.GetColorAttachment(3)->GetDesc(),
                                                          passContext.GetDepthAttachment()->GetDesc(), rowsOf(passData), passContext.GetStream()),
                                     "GBufferPass");
                         }));

      // applying direct lighting (:160-205)
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetColorAttachments({res->directLight.mipImageViewProxies[0]->Id()})
                         .SetInputImages({res->albedo.imageViewProxy->Id(), res->emissive.imageViewProxy->Id(), res->normal.imageViewProxy->Id(),
                                          res->depthStencil.imageViewProxy->Id(), res->shadowMap.imageViewProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::orange, "LightPass")
                         .SetRecordFunc([this, passData, fillLightData, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto shaderDataBuffer = fillLightData(passData);
                           ViewportResources *r = this->viewportResources.get();
                           LgcuCheck(lgcu_direct_light(shaderDataBuffer, passContext.GetImageView(r->albedo.imageViewProxy->Id())->GetDesc(),  // "albedoSampler"
                                                       passContext.GetImageView(r->emissive.imageViewProxy->Id())->GetDesc(),                   // "emissiveSampler"
                                                       passContext.GetImageView(r->normal.imageViewProxy->Id())->GetDesc(),                     // "normalSampler"
                                                       passContext.GetImageView(r->depthStencil.imageViewProxy->Id())->GetDesc(),               // "depthStencilSampler"
                                                       passContext.GetImageView(r->shadowMap.imageViewProxy->Id())->GetDesc(),                  // "shadowmapSampler"
                                                       passContext.GetColorAttachment(0)->GetDesc(), rowsOf(passData), passContext.GetStream()),
                                     "LightPass");
                         }));

      // :207-221 — with a row strip the per-level passes would need per-level row ranges; strips always use Fused
      if (useRows) throw std::runtime_error("SSVGIRenderer: row strips require FrameOptions::Mode::Fused");
      mipBuilder.BuildMips(graph, frameInfo.memoryPool, res->directLight);
      mipBuilder.BuildMips(graph, frameInfo.memoryPool, res->depthMoments);
      for (uint32_t mipLevel = 0; mipLevel < res->blurredDirectLight.mipImageViewProxies.size(); mipLevel++)
        blurBuilder.ApplyBlur(graph, frameInfo.memoryPool, res->directLight.mipImageViewProxies[mipLevel]->Id(),
                              res->blurredDirectLight.mipImageViewProxies[mipLevel]->Id(), mipLevel == 0 ? 0 : 2);
      for (uint32_t mipLevel = 0; mipLevel < res->blurredDirectLight.mipImageViewProxies.size(); mipLevel++)
        blurBuilder.ApplyBlur(graph, frameInfo.memoryPool, res->depthMoments.mipImageViewProxies[mipLevel]->Id(),
                              res->blurredDepthMoments.mipImageViewProxies[mipLevel]->Id(), mipLevel == 0 ? 0 : 2);
    } else if (!useRows || (options.rows.y0 % 16 == 0 && (options.rows.y1 % 16 == 0 || options.rows.y1 == viewportExtent.height))) {
      // K1 + K2 + level-0 blur copies + mip levels 1..4 of both chains in one pass over the fragments (lgcu_frame_front), then the
      // remaining blur / mip work of both chains in one launch (lgcu_frame_chains): 2 kernels for 41 reference passes (:106-221)
      if (stages & FrameOptions::StageFront)
        graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetColorAttachments({res->albedo.imageViewProxy->Id(), res->emissive.imageViewProxy->Id(), res->normal.imageViewProxy->Id(),
                                               res->depthMoments.imageViewProxy->Id(), res->directLight.imageViewProxy->Id(),
                                               res->blurredDirectLight.imageViewProxy->Id(), res->blurredDepthMoments.imageViewProxy->Id()},
                                              vk::AttachmentLoadOp::eClear)
                         .SetDepthAttachment(res->depthStencil.imageViewProxy->Id(), vk::AttachmentLoadOp::eClear)
                         .SetInputImages({res->shadowMap.imageViewProxy->Id()})
                         .SetStorageBuffers({scene->fragmentsProxy->Id(), scene->objectsProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::belizeHole, "FrameFrontPass")
                         .SetRecordFunc([this, passData, fillGBufferData, fillLightData, clearOf, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto gbufferData = fillGBufferData(passData);
                           auto lightData = fillLightData(passData);
                           const lgcu_clear_values clear = clearOf(passContext);
                           Scene *sc = passData.scene;
                           LgcuCheck(lgcu_frame_front(gbufferData, lightData, sc->DrawCallData(passContext.GetBuffer(sc->objectsProxy->Id())),
                                                      sc->DrawCallCount(), static_cast<const lgcu_fragment *>(passContext.GetBuffer(sc->fragmentsProxy->Id())->GetHandle()),
                                                      sc->fragmentPitch, &clear, passContext.GetColorAttachment(0)->GetDesc(), passContext.GetColorAttachment(1)->GetDesc(),
                                                      passContext.GetColorAttachment(2)->GetDesc(), passContext.GetColorAttachment(3)->GetDesc(),
                                                      passContext.GetDepthAttachment()->GetDesc(),
                                                      passContext.GetImageView(this->viewportResources->shadowMap.imageViewProxy->Id())->GetDesc(),
                                                      passContext.GetColorAttachment(4)->GetDesc(), passContext.GetColorAttachment(5)->GetDesc(),
                                                      passContext.GetColorAttachment(6)->GetDesc(), rowsOf(passData), passContext.GetStream()),
                                     "FrameFrontPass");
                         }));
      if (stages & FrameOptions::StageChains)
        graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetStorageImages({res->directLight.imageViewProxy->Id(), res->blurredDirectLight.imageViewProxy->Id(), res->depthMoments.imageViewProxy->Id(),
                                            res->blurredDepthMoments.imageViewProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::nephritis, "FrameChainsPass")
                         .SetRecordFunc([this, passData, rowsOf](RenderGraph::RenderPassContext passContext) {
                           ViewportResources *r = this->viewportResources.get();
                           LgcuCheck(lgcu_frame_chains(passContext.GetImageView(r->directLight.imageViewProxy->Id())->GetDesc(),
                                                       passContext.GetImageView(r->blurredDirectLight.imageViewProxy->Id())->GetDesc(),
                                                       passContext.GetImageView(r->depthMoments.imageViewProxy->Id())->GetDesc(),
                                                       passContext.GetImageView(r->blurredDepthMoments.imageViewProxy->Id())->GetDesc(), 2, rowsOf(passData),
                                                       passContext.GetStream()),
                                     "FrameChainsPass");
                         }));
    } else {
      // row strips that do not sit on the 16-row tile grid of the frame-front kernel: K1+K2 and one chain call per MippedProxy
      // K1 + K2 fused: everything GBufferPass and LightPass write, from one read of the fragments
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetColorAttachments({res->albedo.imageViewProxy->Id(), res->emissive.imageViewProxy->Id(), res->normal.imageViewProxy->Id(),
                                               res->depthMoments.mipImageViewProxies[0]->Id(), res->directLight.mipImageViewProxies[0]->Id()},
                                              vk::AttachmentLoadOp::eClear)
                         .SetDepthAttachment(res->depthStencil.imageViewProxy->Id(), vk::AttachmentLoadOp::eClear)
                         .SetInputImages({res->shadowMap.imageViewProxy->Id()})
                         .SetStorageBuffers({scene->fragmentsProxy->Id(), scene->objectsProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::belizeHole, "GBufferLightPass")
                         .SetRecordFunc([this, passData, fillGBufferData, fillLightData, clearOf, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto gbufferData = fillGBufferData(passData);
                           auto lightData = fillLightData(passData);
                           const lgcu_clear_values clear = clearOf(passContext);
                           Scene *sc = passData.scene;
                           LgcuCheck(lgcu_gbuffer_direct_light(
                                         gbufferData, lightData, sc->DrawCallData(passContext.GetBuffer(sc->objectsProxy->Id())),
                                         sc->DrawCallCount(), static_cast<const lgcu_fragment *>(passContext.GetBuffer(sc->fragmentsProxy->Id())->GetHandle()), sc->fragmentPitch,
                                         &clear, passContext.GetColorAttachment(0)->GetDesc(), passContext.GetColorAttachment(1)->GetDesc(),
                                         passContext.GetColorAttachment(2)->GetDesc(), passContext.GetColorAttachment(3)->GetDesc(), passContext.GetDepthAttachment()->GetDesc(),
                                         passContext.GetImageView(this->viewportResources->shadowMap.imageViewProxy->Id())->GetDesc(),
                                         passContext.GetColorAttachment(4)->GetDesc(), rowsOf(passData), passContext.GetStream()),
                                     "GBufferLightPass");
                         }));
      // K3 + K4 fused per chain (BuildMips + ten ApplyBlur, :207-221)
      auto addChain = [&](const MippedProxy &chain, const MippedProxy &blurred) {
        auto chainId = chain.imageViewProxy->Id(), blurredId = blurred.imageViewProxy->Id();
        graph->AddPass(RenderGraph::RenderPassDesc()
                           .SetStorageImages({chainId, blurredId})
                           .SetRenderAreaExtent(viewportExtent)
                           .SetProfilerInfo(Colors::nephritis, "MipBlurChainPass")
                           .SetRecordFunc([passData, chainId, blurredId, rowsOf](RenderGraph::RenderPassContext passContext) {
                             LgcuCheck(lgcu_mip_blur_chain(passContext.GetImageView(chainId)->GetDesc(), passContext.GetImageView(blurredId)->GetDesc(), 2, rowsOf(passData),
                                                           passContext.GetStream()),
                                       "MipBlurChainPass");
                           }));
      };
      addChain(res->directLight, res->blurredDirectLight);
      addChain(res->depthMoments, res->blurredDepthMoments);
    }

    // calculating indirect lighting (:223-263)
    const uint32_t giFlags = options.giFlags;
    const bool packedGather = options.mode == FrameOptions::Mode::Fused && !(giFlags & LGCU_GI_STRICT);
    auto fillIndirectData = [this](const PassData &pd) {
      pd.memoryPool->BeginSet();
      auto shaderDataBuffer = pd.memoryPool->GetUniformBufferData<lgcu_indirect_lighting_data>("IndirectLightingData");
      shaderDataBuffer->viewMatrix = pd.viewMatrix;
      shaderDataBuffer->projMatrix = pd.projMatrix;
      shaderDataBuffer->viewportExtent[0] = float(this->viewportExtent.width);
      shaderDataBuffer->viewportExtent[1] = float(this->viewportExtent.height);
      shaderDataBuffer->viewportExtent[2] = shaderDataBuffer->viewportExtent[3] = 0.0f;
      pd.memoryPool->EndSet();
      return shaderDataBuffer;
    };
    if (packedGather && (stages & FrameOptions::StageGather)) {
      // builds the gather's private acceleration structure (quad-packed depth pyramid) in a transient buffer
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetInputImages({res->blurredDirectLight.imageViewProxy->Id(), res->blurredDepthMoments.imageViewProxy->Id(), res->normal.imageViewProxy->Id(),
                                          res->depthStencil.imageViewProxy->Id(), res->indirectLight.imageViewProxy->Id()})
                         .SetStorageBuffers({res->gatherScratch->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::carrot, "GatherPackPass")
                         .SetRecordFunc([this, passData, fillIndirectData, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto shaderDataBuffer = fillIndirectData(passData);
                           ViewportResources *r = this->viewportResources.get();
                           Buffer *scratch = passContext.GetBuffer(r->gatherScratch->Id());
                           LgcuCheck(lgcu_gi_gather_pack(shaderDataBuffer, passContext.GetImageView(r->blurredDirectLight.imageViewProxy->Id())->GetDesc(),
                                                         passContext.GetImageView(r->blurredDepthMoments.imageViewProxy->Id())->GetDesc(),
                                                         passContext.GetImageView(r->normal.imageViewProxy->Id())->GetDesc(),
                                                         passContext.GetImageView(r->depthStencil.imageViewProxy->Id())->GetDesc(),
                                                         passContext.GetImageView(r->indirectLight.imageViewProxy->Id())->GetDesc(), scratch->GetHandle(),
                                                         scratch->GetSize(), rowsOf(passData), passContext.GetStream()),
                                     "GatherPackPass");
                         }));
    }
    if (stages & FrameOptions::StageGather) {
      RenderGraph::RenderPassDesc desc;
      desc.SetColorAttachments({res->indirectLight.imageViewProxy->Id()})
          .SetInputImages({res->blurredDirectLight.imageViewProxy->Id(), res->blurredDepthMoments.imageViewProxy->Id(), res->normal.imageViewProxy->Id(),
                           res->depthStencil.imageViewProxy->Id()})
          .SetRenderAreaExtent(viewportExtent)
          .SetProfilerInfo(Colors::sunFlower, "IndirectLightPass");
      if (packedGather) desc.SetStorageBuffers({res->gatherScratch->Id()});
      desc.SetRecordFunc([this, passData, giFlags, packedGather, fillIndirectData, rowsOf](RenderGraph::RenderPassContext passContext) {
        auto shaderDataBuffer = fillIndirectData(passData);
        ViewportResources *r = this->viewportResources.get();
        const lgcu_image *light = passContext.GetImageView(r->blurredDirectLight.imageViewProxy->Id())->GetDesc();     // "blurredDirectLightSampler"
        const lgcu_image *moments = passContext.GetImageView(r->blurredDepthMoments.imageViewProxy->Id())->GetDesc();  // "blurredDepthMomentsSampler"
        const lgcu_image *normal = passContext.GetImageView(r->normal.imageViewProxy->Id())->GetDesc();                // "normalSampler"
        const lgcu_image *depth = passContext.GetImageView(r->depthStencil.imageViewProxy->Id())->GetDesc();           // "depthStencilSampler"
        if (packedGather) {
          Buffer *scratch = passContext.GetBuffer(r->gatherScratch->Id());
          LgcuCheck(lgcu_gi_gather_packed(shaderDataBuffer, light, moments, normal, depth, passContext.GetColorAttachment(0)->GetDesc(), scratch->GetHandle(),
                                          scratch->GetSize(), rowsOf(passData), passContext.GetStream()),
                    "IndirectLightPass");
        } else {
          LgcuCheck(lgcu_gi_gather(shaderDataBuffer, light, moments, normal, depth, passContext.GetColorAttachment(0)->GetDesc(), giFlags, rowsOf(passData),
                                   passContext.GetStream()),
                    "IndirectLightPass");
        }
      });
      graph->AddPass(desc);
    }

    const int denoiserRadius = options.denoiserRadius;
    // The radius-2 denoiser reads indirectLight and depthMoments on rows -2..+1 around its own. A row strip of a frame that is sharded
    // over GPUs does not own those rows of indirectLight and no stage exchanges them (sharding.py plans halos for radius 0), so the
    // combination is refused instead of fitting stale rows at the seams.
    if (useRows && denoiserRadius != 0 && (stages & FrameOptions::StageFinal) && (options.rows.y0 > 0 || options.rows.y1 < viewportExtent.height))
      throw std::runtime_error("SSVGIRenderer: denoiserRadius != 0 on a row strip needs indirectLight halo rows that are not exchanged; render the frame whole");
    auto fillDenoiserData = [this, denoiserRadius](const PassData &pd) {
      pd.memoryPool->BeginSet();
      auto shaderDataBuffer = pd.memoryPool->GetUniformBufferData<lgcu_denoiser_data>("DenoiserData");
      shaderDataBuffer->viewMatrix = pd.viewMatrix;
      shaderDataBuffer->projMatrix = pd.projMatrix;
      shaderDataBuffer->viewportExtent[0] = float(this->viewportExtent.width);
      shaderDataBuffer->viewportExtent[1] = float(this->viewportExtent.height);
      shaderDataBuffer->viewportExtent[2] = shaderDataBuffer->viewportExtent[3] = 0.0f;
      shaderDataBuffer->radius = denoiserRadius;
      pd.memoryPool->EndSet();
      return shaderDataBuffer;
    };
    auto fillFinalData = [](const PassData &pd) {
      pd.memoryPool->BeginSet();
      auto shaderDataBuffer = pd.memoryPool->GetUniformBufferData<lgcu_final_gatherer_data>("FinalGathererData");
      shaderDataBuffer->viewMatrix = pd.viewMatrix;
      shaderDataBuffer->projMatrix = pd.projMatrix;
      pd.memoryPool->EndSet();
      return shaderDataBuffer;
    };

    if (options.mode == FrameOptions::Mode::PassGranular) {
      // denoising indirect lighting (:265-302)
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetColorAttachments({res->denoisedIndirectLight.imageViewProxy->Id()})
                         .SetInputImages({res->depthMoments.imageViewProxy->Id(), res->normal.imageViewProxy->Id(), res->indirectLight.imageViewProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::greenSea, "DenoiserPass")
                         .SetRecordFunc([this, passData, fillDenoiserData, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto shaderDataBuffer = fillDenoiserData(passData);
                           ViewportResources *r = this->viewportResources.get();
                           LgcuCheck(lgcu_denoise(shaderDataBuffer, passContext.GetImageView(r->indirectLight.imageViewProxy->Id())->GetDesc(), // "noisySampler"
                                                  passContext.GetImageView(r->normal.imageViewProxy->Id())->GetDesc(),                          // "normalSampler"
                                                  passContext.GetImageView(r->depthMoments.imageViewProxy->Id())->GetDesc(),                    // "depthStencilSampler" (:293)
                                                  passContext.GetColorAttachment(0)->GetDesc(), rowsOf(passData), passContext.GetStream()),
                                     "DenoiserPass");
                         }));
      // final gathering (:304-342)
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetColorAttachments({frameInfo.swapchainImageViewProxyId})
                         .SetInputImages({res->directLight.imageViewProxy->Id(), res->blurredDirectLight.imageViewProxy->Id(), res->albedo.imageViewProxy->Id(),
                                          res->denoisedIndirectLight.imageViewProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::pomegranate, "GatheringPass")
                         .SetRecordFunc([this, passData, fillFinalData, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto shaderDataBuffer = fillFinalData(passData);
                           ViewportResources *r = this->viewportResources.get();
                           LgcuCheck(lgcu_final_gather(shaderDataBuffer, passContext.GetImageView(r->directLight.imageViewProxy->Id())->GetDesc(),  // "directLightSampler"
                                                       passContext.GetImageView(r->blurredDirectLight.imageViewProxy->Id())->GetDesc(),             // "blurredDirectLightSampler"
                                                       passContext.GetImageView(r->albedo.imageViewProxy->Id())->GetDesc(),                         // "albedoSampler"
                                                       passContext.GetImageView(r->denoisedIndirectLight.imageViewProxy->Id())->GetDesc(),          // "indirectLightSampler"
                                                       passContext.GetColorAttachment(0)->GetDesc(), rowsOf(passData), passContext.GetStream()),
                                     "GatheringPass");
                         }));
    } else if (stages & FrameOptions::StageFinal) {
      // K6 + K7 fused
      graph->AddPass(RenderGraph::RenderPassDesc()
                         .SetColorAttachments({res->denoisedIndirectLight.imageViewProxy->Id(), frameInfo.swapchainImageViewProxyId})
                         .SetInputImages({res->depthMoments.imageViewProxy->Id(), res->normal.imageViewProxy->Id(), res->indirectLight.imageViewProxy->Id(),
                                          res->directLight.imageViewProxy->Id(), res->blurredDirectLight.imageViewProxy->Id(), res->albedo.imageViewProxy->Id()})
                         .SetRenderAreaExtent(viewportExtent)
                         .SetProfilerInfo(Colors::pomegranate, "DenoiseGatheringPass")
                         .SetRecordFunc([this, passData, fillDenoiserData, fillFinalData, rowsOf](RenderGraph::RenderPassContext passContext) {
                           auto denoiserData = fillDenoiserData(passData);
                           auto finalData = fillFinalData(passData);
                           ViewportResources *r = this->viewportResources.get();
                           LgcuCheck(lgcu_denoise_final_gather(denoiserData, finalData, passContext.GetImageView(r->indirectLight.imageViewProxy->Id())->GetDesc(),
                                                               passContext.GetImageView(r->normal.imageViewProxy->Id())->GetDesc(),
                                                               passContext.GetImageView(r->depthMoments.imageViewProxy->Id())->GetDesc(),
                                                               passContext.GetColorAttachment(0)->GetDesc(),
                                                               passContext.GetImageView(r->directLight.imageViewProxy->Id())->GetDesc(),
                                                               passContext.GetImageView(r->blurredDirectLight.imageViewProxy->Id())->GetDesc(),
                                                               passContext.GetImageView(r->albedo.imageViewProxy->Id())->GetDesc(),
                                                               passContext.GetColorAttachment(1)->GetDesc(), rowsOf(passData), passContext.GetStream()),
                                     "DenoiseGatheringPass");
                         }));
    }

    if (options.debugOverlay && (options.mode == FrameOptions::Mode::PassGranular || (stages & FrameOptions::StageFinal))) {
      std::vector<RenderGraph::ImageViewProxyId> debugProxies; // :344-350
      debugProxies.push_back(res->normal.imageViewProxy->Id());
      debugProxies.push_back(res->albedo.imageViewProxy->Id());
      debugProxies.push_back(res->indirectLight.imageViewProxy->Id());
      debugProxies.push_back(res->denoisedIndirectLight.imageViewProxy->Id());
      debugRenderer.RenderImageViews(graph, frameInfo.memoryPool, frameInfo.swapchainImageViewProxyId, debugProxies, useRows ? &options.rows : nullptr);
    }
  }

  void ReloadShaders() {
    mipBuilder.ReloadShaders();
    blurBuilder.ReloadShaders();
    debugRenderer.ReloadShaders();
  }

  // :391-419
  struct ViewportResources {
    ViewportResources(RenderGraph *renderGraph, glm::uvec2 screenSize)
        : albedo(renderGraph, vk::Format::eR16G16B16A16Sfloat, screenSize, colorImageUsage),
          emissive(renderGraph, vk::Format::eR16G16B16A16Sfloat, screenSize, colorImageUsage),
          normal(renderGraph, vk::Format::eR16G16B16A16Sfloat, screenSize, colorImageUsage),
          depthMoments(renderGraph, vk::Format::eR32G32Sfloat, screenSize, colorImageUsage),
          blurredDepthMoments(renderGraph, vk::Format::eR32G32Sfloat, screenSize, colorImageUsage),
          depthStencil(renderGraph, vk::Format::eD32Sfloat, screenSize, depthImageUsage),
          directLight(renderGraph, vk::Format::eR16G16B16A16Sfloat, screenSize, colorImageUsage),
          blurredDirectLight(renderGraph, vk::Format::eR16G16B16A16Sfloat, screenSize, colorImageUsage),
          shadowMap(renderGraph, vk::Format::eD32Sfloat, glm::uvec2(1024, 1024), depthImageUsage),
          indirectLight(renderGraph, vk::Format::eR16G16B16A16Sfloat, screenSize, colorImageUsage),
          denoisedIndirectLight(renderGraph, vk::Format::eR16G16B16A16Sfloat, screenSize, colorImageUsage),
          gatherScratch(renderGraph->AddBuffer<GatherQuad>(uint32_t(lgcu_gather_scratch_bytes(screenSize.x, screenSize.y, 10) / sizeof(GatherQuad)))) {}
    struct GatherQuad { float tap[4]; };
    UnmippedProxy albedo;
    UnmippedProxy emissive;
    UnmippedProxy normal;
    MippedProxy depthMoments;
    MippedProxy blurredDepthMoments;
    UnmippedProxy depthStencil;
    MippedProxy directLight;
    MippedProxy blurredDirectLight;
    UnmippedProxy shadowMap;
    UnmippedProxy indirectLight;
    UnmippedProxy denoisedIndirectLight;
    RenderGraph::BufferProxyUnique gatherScratch; // transient: quad-packed depth pyramid of the GI gather (lgcu_gi_gather_pack)
  };
  ViewportResources *GetViewportResources() { return viewportResources.get(); }
  vk::Extent2D GetViewportExtent() const { return viewportExtent; }

private:
  std::unique_ptr<ViewportResources> viewportResources;
  vk::Extent2D viewportExtent;
  MipBuilder mipBuilder;
  BlurBuilder blurBuilder;
  DebugRenderer debugRenderer;
  Sampler screenspaceSampler;
  Sampler shadowmapSampler;
  Core *core;
};

} // namespace legit_cuda
