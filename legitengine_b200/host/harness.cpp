// harness.cpp — headless frame harness over the legit_cuda renderer (C ABI in include/lgcu_harness.h).
// Plays the role of the reference's frame loop (src/main.cpp:197-312 + LV/PresentQueue.h:85-166): per frame it maps the
// shader memory pool, lets SSVGIRenderer::RenderFrame declare the passes and runs RenderGraph::Execute on the stream.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lgcu_harness.h"
#include "legit_cuda/ExternalImage.h" // compiled here so that the Vulkan hand-back wrappers stay buildable (no caller in the headless harness)
#include "legit_cuda/InterleaveBuilder.h"
#include "legit_cuda/SSVGIRenderer.h"

using namespace legit_cuda;

namespace {

thread_local char g_error[512] = "";

int setError(int status, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return status;
}

} // namespace

struct lgh_renderer {
  uint32_t width = 0, height = 0;
  cudaStream_t stream = nullptr;
  bool ownsStream = false;
  std::unique_ptr<Core> core;
  std::unique_ptr<SSVGIRenderer> renderer;
  ShaderMemoryPool memoryPool;
  CpuProfiler cpuProfiler;
  GpuProfiler gpuProfiler;
  bool profiled = false;
  int lastPassCount = 0;

  std::unique_ptr<Buffer> fragments, objects, lightDepth;
  uint32_t objectCapacity = 0, objectCount = 0, lightDepthSize = 0;
  std::unique_ptr<Scene> scene;
  // mesh form of the scene (lgh_upload_mesh): device copies of the reference's vertex / index buffers, draw list, per-object constants
  std::unique_ptr<Buffer> meshVertices, meshIndices, meshDraws, meshObjects, rasterScratch;
  lgcu_mesh_scene meshDesc{};
  bool useMesh = false;
  bool debugOverlay = false;

  std::unique_ptr<ImageData> swapchainImage;         // the renderer's own image (kept while an external one is in use: peers may map it)
  std::unique_ptr<ImageData> externalSwapchainImage; // non-owning wrapper of the image lgh_set_external_swapchain put behind the proxy
  std::unique_ptr<ImageView> swapchainView;
  RenderGraph::ImageViewProxyUnique swapchainProxy;

  Camera camera = DefaultCamera(), light = DefaultLight();

  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graphExec = nullptr;
  int graphKernels = 0;

  uint64_t fragmentPitch() const { return uint64_t(width) * sizeof(lgcu_fragment); }

  void ensureScene() {
    if (!lightDepth) {
      lightDepthSize = 1024;
      lightDepth.reset(new Buffer(size_t(lightDepthSize) * lightDepthSize * 4));
    }
    if (!objects) {
      objectCapacity = 4096;
      objects.reset(new Buffer(size_t(objectCapacity) * sizeof(lgcu_draw_call_data)));
    }
    if (!scene) scene.reset(new Scene(core->GetRenderGraph(), fragments.get(), fragmentPitch(), objects.get(), objectCount, lightDepth.get(), lightDepthSize));
    scene->objectsCount = objectCount;
    scene->lightDepthSize = lightDepthSize;
    if (useMesh)
      scene->SetMesh(core->GetRenderGraph(), meshDesc, rasterScratch.get());
    else
      scene->ClearMesh();
  }

  static void fit(std::unique_ptr<Buffer> &b, size_t bytes) {
    if (!b || b->GetSize() < bytes) b.reset(new Buffer(bytes + bytes / 4));
  }

  void declareAndExecute(uint32_t mode, int32_t denoiserRadius, uint32_t giFlags, const lgcu_rows *rows, bool profile, uint32_t stages = FrameOptions::StageAll) {
    ensureScene();
    memoryPool.MapBuffer();
    FrameInfo frameInfo;
    frameInfo.memoryPool = &memoryPool;
    frameInfo.swapchainImageViewProxyId = swapchainProxy->Id();
    FrameOptions options;
    options.mode = mode == LGH_MODE_PASS_GRANULAR ? FrameOptions::Mode::PassGranular : FrameOptions::Mode::Fused;
    options.denoiserRadius = denoiserRadius;
    options.giFlags = giFlags;
    options.stages = stages;
    options.debugOverlay = debugOverlay;
    if (rows) {
      options.useRows = true;
      options.rows = *rows;
    }
    renderer->RenderFrame(frameInfo, camera, light, scene.get(), options);
    lastPassCount = int(core->GetRenderGraph()->GetPendingPassCount());
    core->GetRenderGraph()->Execute(stream, profile ? &cpuProfiler : nullptr, profile ? &gpuProfiler : nullptr);
    profiled = profile;
  }

  ImageView *findImage(const std::string &name) {
    auto *res = renderer->GetViewportResources();
    RenderGraph *g = core->GetRenderGraph();
    if (name == "swapchain") return swapchainView.get();
    if (name == "albedo") return g->GetResolvedImageView(res->albedo.imageViewProxy->Id());
    if (name == "emissive") return g->GetResolvedImageView(res->emissive.imageViewProxy->Id());
    if (name == "normal") return g->GetResolvedImageView(res->normal.imageViewProxy->Id());
    if (name == "depthMoments") return g->GetResolvedImageView(res->depthMoments.imageViewProxy->Id());
    if (name == "blurredDepthMoments") return g->GetResolvedImageView(res->blurredDepthMoments.imageViewProxy->Id());
    if (name == "depthStencil") return g->GetResolvedImageView(res->depthStencil.imageViewProxy->Id());
    if (name == "directLight") return g->GetResolvedImageView(res->directLight.imageViewProxy->Id());
    if (name == "blurredDirectLight") return g->GetResolvedImageView(res->blurredDirectLight.imageViewProxy->Id());
    if (name == "shadowMap") return g->GetResolvedImageView(res->shadowMap.imageViewProxy->Id());
    if (name == "indirectLight") return g->GetResolvedImageView(res->indirectLight.imageViewProxy->Id());
    if (name == "denoisedIndirectLight") return g->GetResolvedImageView(res->denoisedIndirectLight.imageViewProxy->Id());
    return nullptr;
  }
};

#define LGH_TRY(body)                                                  \
  try {                                                                \
    body;                                                              \
    return LGCU_OK;                                                    \
  } catch (const std::exception &e) {                                  \
    return setError(LGCU_ERR_CUDA, "%s", e.what());                    \
  }

extern "C" {

const char *lgh_last_error(void) { return g_error; }

lgh_renderer *lgh_create(uint32_t width, uint32_t height, void *stream) {
  try {
    if (width == 0 || height == 0) {
      setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_create: %ux%u", width, height);
      return nullptr;
    }
    std::unique_ptr<lgh_renderer> r(new lgh_renderer());
    r->width = width;
    r->height = height;
    if (stream) {
      r->stream = static_cast<cudaStream_t>(stream);
    } else {
      CudaCheck(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking), "cudaStreamCreate");
      r->ownsStream = true;
    }
    r->core.reset(new Core(r->stream));
    r->renderer.reset(new SSVGIRenderer(r->core.get()));
    r->renderer->RecreateSwapchainResources(vk::Extent2D(width, height), 1); // main.cpp:197-205
    r->fragments.reset(new Buffer(size_t(r->fragmentPitch()) * height));
    r->swapchainImage.reset(new ImageData(vk::Format::eB8G8R8A8Srgb, glm::uvec2(width, height), 1)); // LV/Swapchain.h:108
    r->swapchainView.reset(new ImageView(r->swapchainImage.get(), 0, 1));
    r->swapchainProxy = r->core->GetRenderGraph()->AddExternalImageView(r->swapchainView.get(), ImageUsageTypes::Present); // LV/PresentQueue.h:106-110
    return r.release();
  } catch (const std::exception &e) {
    setError(LGCU_ERR_CUDA, "lgh_create: %s", e.what());
    return nullptr;
  }
}

void lgh_destroy(lgh_renderer *r) {
  if (!r) return;
  cudaStreamSynchronize(r->stream);
  if (r->graphExec) cudaGraphExecDestroy(r->graphExec);
  if (r->graph) cudaGraphDestroy(r->graph);
  r->scene.reset();
  r->swapchainProxy.Reset();
  r->renderer.reset();
  r->core.reset();
  if (r->ownsStream) cudaStreamDestroy(r->stream);
  delete r;
}

int lgh_set_camera(lgh_renderer *r, const float camPos[3], float camVertAngle, float camHorAngle, const float lightPos[3], float lightVertAngle,
                   float lightHorAngle) {
  if (!r || !camPos || !lightPos) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_set_camera: null argument");
  r->camera.pos = vec3{camPos[0], camPos[1], camPos[2]};
  r->camera.vertAngle = camVertAngle;
  r->camera.horAngle = camHorAngle;
  r->light.pos = vec3{lightPos[0], lightPos[1], lightPos[2]};
  r->light.vertAngle = lightVertAngle;
  r->light.horAngle = lightHorAngle;
  return LGCU_OK;
}

int lgh_upload_fragments(lgh_renderer *r, const lgcu_fragment *hostFragments, uint64_t hostPitchBytes, uint32_t rowBegin, uint32_t rowEnd) {
  if (!r || !hostFragments || rowEnd > r->height || rowBegin > rowEnd || hostPitchBytes < r->fragmentPitch())
    return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_upload_fragments: bad arguments");
  LGH_TRY({
    const uint64_t pitch = r->fragmentPitch();
    uint8_t *dst = static_cast<uint8_t *>(r->fragments->GetHandle()) + pitch * rowBegin;
    const uint8_t *src = reinterpret_cast<const uint8_t *>(hostFragments) + hostPitchBytes * rowBegin;
    if (hostPitchBytes == pitch)
      CudaCheck(cudaMemcpyAsync(dst, src, pitch * (rowEnd - rowBegin), cudaMemcpyHostToDevice, r->stream), "upload fragments");
    else
      CudaCheck(cudaMemcpy2DAsync(dst, pitch, src, hostPitchBytes, pitch, rowEnd - rowBegin, cudaMemcpyHostToDevice, r->stream), "upload fragments");
  })
}

int lgh_upload_objects(lgh_renderer *r, const lgcu_draw_call_data *hostObjects, uint32_t count) {
  if (!r || (!hostObjects && count)) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_upload_objects: null argument");
  LGH_TRY({
    r->ensureScene();
    if (count > r->objectCapacity) throw std::runtime_error("lgh_upload_objects: more than 4096 draw calls");
    CudaCheck(cudaMemcpyAsync(r->objects->GetHandle(), hostObjects, size_t(count) * sizeof(lgcu_draw_call_data), cudaMemcpyHostToDevice, r->stream), "upload objects");
    r->objectCount = count;
  })
}

int lgh_upload_light_depth(lgh_renderer *r, const float *hostDepth, uint32_t size) {
  if (!r || !hostDepth || size != 1024) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_upload_light_depth: the shadow map is 1024x1024 (SSVGIRenderer.h:402)");
  LGH_TRY({
    r->ensureScene();
    CudaCheck(cudaMemcpyAsync(r->lightDepth->GetHandle(), hostDepth, size_t(size) * size * 4, cudaMemcpyHostToDevice, r->stream), "upload light depth");
  })
}

int lgh_upload_mesh(lgh_renderer *r, const lgcu_vertex *hostVertices, uint32_t nVertices, const uint32_t *hostIndices, uint32_t nIndices,
                    const lgcu_draw *hostDraws, uint32_t nDraws, const lgcu_draw_call_data *hostObjects, uint32_t nObjects) {
  if (!r || !hostVertices || !hostIndices || !hostDraws || !hostObjects || !nVertices || !nIndices || !nDraws || !nObjects)
    return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_upload_mesh: null / empty argument");
  uint64_t triangles = 0;
  for (uint32_t i = 0; i < nDraws; i++) {
    const lgcu_draw &d = hostDraws[i];
    if (d.indexCount % 3 != 0 || uint64_t(d.firstIndex) + d.indexCount > nIndices || d.objectId >= nObjects || d.firstTriangle != triangles)
      return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_upload_mesh: draw %u is out of range or its firstTriangle is not filled (lgcu_raster_prepare_draws)", i);
    triangles += d.indexCount / 3;
  }
  if (triangles == 0 || triangles > 0x7fffffffull) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_upload_mesh: %llu triangles", (unsigned long long)triangles);
  LGH_TRY({
    // growing a buffer invalidates a captured frame (it holds the old addresses); the caller re-captures after a scene change
    lgh_renderer::fit(r->meshVertices, size_t(nVertices) * sizeof(lgcu_vertex));
    lgh_renderer::fit(r->meshIndices, size_t(nIndices) * 4);
    lgh_renderer::fit(r->meshDraws, size_t(nDraws) * sizeof(lgcu_draw));
    lgh_renderer::fit(r->meshObjects, size_t(nObjects) * sizeof(lgcu_draw_call_data));
    const uint64_t frameScratch = lgcu_raster_scratch_bytes(uint32_t(triangles), r->width, r->height);
    const uint64_t shadowScratch = lgcu_raster_scratch_bytes(uint32_t(triangles), 1024, 1024); // SSVGIRenderer.h:402
    lgh_renderer::fit(r->rasterScratch, size_t(frameScratch > shadowScratch ? frameScratch : shadowScratch));
    CudaCheck(cudaMemcpyAsync(r->meshVertices->GetHandle(), hostVertices, size_t(nVertices) * sizeof(lgcu_vertex), cudaMemcpyHostToDevice, r->stream), "upload vertices");
    CudaCheck(cudaMemcpyAsync(r->meshIndices->GetHandle(), hostIndices, size_t(nIndices) * 4, cudaMemcpyHostToDevice, r->stream), "upload indices");
    CudaCheck(cudaMemcpyAsync(r->meshDraws->GetHandle(), hostDraws, size_t(nDraws) * sizeof(lgcu_draw), cudaMemcpyHostToDevice, r->stream), "upload draws");
    CudaCheck(cudaMemcpyAsync(r->meshObjects->GetHandle(), hostObjects, size_t(nObjects) * sizeof(lgcu_draw_call_data), cudaMemcpyHostToDevice, r->stream), "upload objects");
    r->meshDesc.vertices = static_cast<const lgcu_vertex *>(r->meshVertices->GetHandle());
    r->meshDesc.indices = static_cast<const uint32_t *>(r->meshIndices->GetHandle());
    r->meshDesc.draws = static_cast<const lgcu_draw *>(r->meshDraws->GetHandle());
    r->meshDesc.objects = static_cast<const lgcu_draw_call_data *>(r->meshObjects->GetHandle());
    r->meshDesc.nVertices = nVertices;
    r->meshDesc.nIndices = nIndices;
    r->meshDesc.nDraws = nDraws;
    r->meshDesc.nObjects = nObjects;
    r->meshDesc.nTriangles = uint32_t(triangles);
    r->useMesh = true;
  })
}

int lgh_set_debug_overlay(lgh_renderer *r, uint32_t enable) {
  if (!r) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_set_debug_overlay: null renderer");
  r->debugOverlay = enable != 0;
  return LGCU_OK;
}

int lgh_set_external_swapchain(lgh_renderer *r, void *deviceBase) {
  if (!r) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_set_external_swapchain: null renderer");
  LGH_TRY({
    // the swapchain view is an EXTERNAL image-view proxy, as in the reference (LV/PresentQueue.h:106-110): swap the image behind it
    lgcu_image desc = r->swapchainImage->GetDesc();
    r->swapchainProxy.Reset();
    r->swapchainView.reset();
    r->externalSwapchainImage.reset();
    if (deviceBase) {
      desc.base = deviceBase;
      r->externalSwapchainImage.reset(new ImageData(desc)); // non-owning: e.g. the presenting GPU's swapchain image through a peer mapping
    }
    r->swapchainView.reset(new ImageView(deviceBase ? r->externalSwapchainImage.get() : r->swapchainImage.get(), 0, 1));
    r->swapchainProxy = r->core->GetRenderGraph()->AddExternalImageView(r->swapchainView.get(), ImageUsageTypes::Present);
    return LGCU_OK;
  })
}

int lgh_run_interleave(lgh_renderer *r, const char *srcName, uint32_t gridX, uint32_t gridY, void *hostDeinterleaved, void *hostRoundTrip, uint64_t hostPitchBytes) {
  if (!r || !srcName || !hostDeinterleaved || !hostRoundTrip) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_run_interleave: null argument");
  ImageView *src = r->findImage(srcName);
  if (!src || src->GetDesc()->mipCount != 1) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_run_interleave: '%s' is not a resolved single-level image", srcName);
  try {
    // InterleaveBuilder on the rendergraph, as the reference's interleaved renderers use it (LSGIRenderer.h:83-85): two transient
    // images of the source's format and size, Deinterleave(src -> a), Interleave(a -> b), then both are read back
    RenderGraph *graph = r->core->GetRenderGraph();
    InterleaveBuilder builder(r->core.get());
    const lgcu_image *d = src->GetDesc();
    const glm::uvec2 size(d->width, d->height);
    auto srcProxy = graph->AddExternalImageView(src);
    auto imgA = graph->AddImage(vk::Format(d->format), 1, 1, size, colorImageUsage);
    auto imgB = graph->AddImage(vk::Format(d->format), 1, 1, size, colorImageUsage);
    auto viewA = graph->AddImageView(imgA->Id(), 0, 1, 0, 1);
    auto viewB = graph->AddImageView(imgB->Id(), 0, 1, 0, 1);
    r->memoryPool.MapBuffer();
    builder.Deinterleave(graph, &r->memoryPool, srcProxy->Id(), viewA->Id(), glm::uvec2(gridX, gridY));
    builder.Interleave(graph, &r->memoryPool, viewA->Id(), viewB->Id(), glm::uvec2(gridX, gridY));
    graph->Execute(r->stream, nullptr, nullptr);
    const uint64_t rowBytes = uint64_t(d->width) * lgcu_format_texel_size(d->format);
    if (hostPitchBytes < rowBytes) throw std::runtime_error("lgh_run_interleave: host pitch");
    void *hosts[2] = {hostDeinterleaved, hostRoundTrip};
    ImageView *views[2] = {graph->GetResolvedImageView(viewA->Id()), graph->GetResolvedImageView(viewB->Id())};
    for (int i = 0; i < 2; i++) {
      const lgcu_image *v = views[i]->GetDesc();
      CudaCheck(cudaMemcpy2DAsync(hosts[i], hostPitchBytes, static_cast<const uint8_t *>(v->base) + v->levelOffset[0], v->levelPitch[0], rowBytes, d->height,
                                  cudaMemcpyDeviceToHost, r->stream),
                "download interleave result");
    }
    CudaCheck(cudaStreamSynchronize(r->stream), "cudaStreamSynchronize");
    return LGCU_OK;
  } catch (const std::exception &e) {
    return setError(LGCU_ERR_CUDA, "lgh_run_interleave: %s", e.what());
  }
}

int lgh_use_mesh(lgh_renderer *r, uint32_t enable) {
  if (!r) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_use_mesh: null renderer");
  if (enable && !r->meshVertices) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_use_mesh: no mesh uploaded");
  r->useMesh = enable != 0;
  return LGCU_OK;
}

int lgh_render_frame(lgh_renderer *r, uint32_t mode, int32_t denoiserRadius, uint32_t giFlags, const lgcu_rows *rows, uint32_t profile) {
  if (!r) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_render_frame: null renderer");
  LGH_TRY(r->declareAndExecute(mode, denoiserRadius, giFlags, rows, profile != 0))
}

int lgh_render_stages(lgh_renderer *r, uint32_t mode, int32_t denoiserRadius, uint32_t giFlags, const lgcu_rows *rows, uint32_t stages) {
  if (!r) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_render_stages: null renderer");
  if (mode != LGH_MODE_FUSED || (stages & ~uint32_t(LGH_STAGE_ALL)) != 0) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_render_stages: fused mode and stage bits 1..8 only");
  LGH_TRY(r->declareAndExecute(mode, denoiserRadius, giFlags, rows, false, stages))
}

// ---- CUDA IPC plumbing for the peer-to-peer strip exchange (one process per GPU) -------------------------------------------
int lgh_ipc_export_image(lgh_renderer *r, const char *name, unsigned char handle[64], uint64_t *bytes) {
  if (!r || !name || !handle) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_ipc_export_image: null argument");
  ImageView *view = r->findImage(name);
  if (!view) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_ipc_export_image: unknown or unresolved image '%s'", name);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, view->GetImageData()->GetDesc().base);
  if (e != cudaSuccess) return setError(LGCU_ERR_CUDA, "cudaIpcGetMemHandle(%s): %s", name, cudaGetErrorString(e));
  std::memcpy(handle, &h, 64);
  if (bytes) *bytes = view->GetImageData()->GetByteSize();
  return LGCU_OK;
}
int lgh_ipc_export_ptr(void *devicePtr, unsigned char handle[64]) {
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, devicePtr);
  if (e != cudaSuccess) return setError(LGCU_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  std::memcpy(handle, &h, 64);
  return LGCU_OK;
}
int lgh_ipc_open(const unsigned char handle[64], void **devicePtr) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  const cudaError_t e = cudaIpcOpenMemHandle(devicePtr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return setError(LGCU_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return LGCU_OK;
}
int lgh_ipc_close(void *devicePtr) {
  const cudaError_t e = cudaIpcCloseMemHandle(devicePtr);
  if (e != cudaSuccess) return setError(LGCU_ERR_CUDA, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
  return LGCU_OK;
}
int lgh_device_alloc_zeroed(uint64_t bytes, void **devicePtr) {
  cudaError_t e = cudaMalloc(devicePtr, bytes ? bytes : 1);
  if (e == cudaSuccess) e = cudaMemset(*devicePtr, 0, bytes);
  if (e != cudaSuccess) return setError(LGCU_ERR_CUDA, "lgh_device_alloc_zeroed: %s", cudaGetErrorString(e));
  return LGCU_OK;
}
int lgh_device_free(void *devicePtr) {
  cudaFree(devicePtr);
  return LGCU_OK;
}

int lgh_capture_frame(lgh_renderer *r, uint32_t mode, int32_t denoiserRadius, uint32_t giFlags, const lgcu_rows *rows) {
  if (!r) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_capture_frame: null renderer");
  try {
    if (r->graphExec) {
      cudaGraphExecDestroy(r->graphExec);
      r->graphExec = nullptr;
    }
    if (r->graph) {
      cudaGraphDestroy(r->graph);
      r->graph = nullptr;
    }
    CudaCheck(cudaStreamBeginCapture(r->stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
    try {
      r->declareAndExecute(mode, denoiserRadius, giFlags, rows, false);
    } catch (...) {
      cudaGraph_t dead = nullptr;
      cudaStreamEndCapture(r->stream, &dead);
      if (dead) cudaGraphDestroy(dead);
      throw;
    }
    CudaCheck(cudaStreamEndCapture(r->stream, &r->graph), "cudaStreamEndCapture");
    CudaCheck(cudaGraphInstantiate(&r->graphExec, r->graph, 0), "cudaGraphInstantiate");
    size_t nodes = 0;
    CudaCheck(cudaGraphGetNodes(r->graph, nullptr, &nodes), "cudaGraphGetNodes");
    std::vector<cudaGraphNode_t> list(nodes);
    if (nodes) CudaCheck(cudaGraphGetNodes(r->graph, list.data(), &nodes), "cudaGraphGetNodes");
    r->graphKernels = 0;
    for (size_t i = 0; i < nodes; i++) {
      cudaGraphNodeType type;
      CudaCheck(cudaGraphNodeGetType(list[i], &type), "cudaGraphNodeGetType");
      if (type == cudaGraphNodeTypeKernel) r->graphKernels++;
    }
    return LGCU_OK;
  } catch (const std::exception &e) {
    return setError(LGCU_ERR_CUDA, "lgh_capture_frame: %s", e.what());
  }
}

int lgh_replay_frame(lgh_renderer *r) {
  if (!r || !r->graphExec) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_replay_frame: no captured frame");
  LGH_TRY(CudaCheck(cudaGraphLaunch(r->graphExec, r->stream), "cudaGraphLaunch"))
}

int lgh_captured_kernel_count(lgh_renderer *r) { return r ? r->graphKernels : 0; }
int lgh_last_pass_count(lgh_renderer *r) { return r ? r->lastPassCount : 0; }

int lgh_image_desc(lgh_renderer *r, const char *name, lgcu_image *out) {
  if (!r || !name || !out) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_image_desc: null argument");
  ImageView *view = r->findImage(name);
  if (!view) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_image_desc: unknown or unresolved image '%s' (render a frame first)", name);
  *out = *view->GetDesc();
  return LGCU_OK;
}

int lgh_download_image(lgh_renderer *r, const char *name, uint32_t level, void *host, uint64_t hostPitchBytes, uint32_t rowBegin, uint32_t rowEnd) {
  if (!r || !name || !host) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_download_image: null argument");
  ImageView *view = r->findImage(name);
  if (!view) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_download_image: unknown or unresolved image '%s'", name);
  const lgcu_image *d = view->GetDesc();
  if (level >= d->imageMipCount) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_download_image: level %u", level);
  const uint32_t w = d->width >> level, h = d->height >> level;
  const uint64_t rowBytes = uint64_t(w) * lgcu_format_texel_size(d->format);
  if (rowEnd > h || rowBegin > rowEnd || hostPitchBytes < rowBytes) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_download_image: rows / pitch");
  LGH_TRY({
    const uint8_t *src = static_cast<const uint8_t *>(d->base) + d->levelOffset[level] + uint64_t(d->levelPitch[level]) * rowBegin;
    uint8_t *dst = static_cast<uint8_t *>(host) + hostPitchBytes * rowBegin;
    if (rowEnd > rowBegin)
      CudaCheck(cudaMemcpy2DAsync(dst, hostPitchBytes, src, d->levelPitch[level], rowBytes, rowEnd - rowBegin, cudaMemcpyDeviceToHost, r->stream), "download image");
  })
}

int lgh_sync(lgh_renderer *r) {
  if (!r) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_sync: null renderer");
  LGH_TRY(CudaCheck(cudaStreamSynchronize(r->stream), "cudaStreamSynchronize"))
}

int lgh_get_profile(lgh_renderer *r, char *nameBuf, uint64_t nameBufBytes, float *ms, uint32_t maxPasses) {
  if (!r || !r->profiled) return setError(LGCU_ERR_INVALID_ARGUMENT, "lgh_get_profile: the last frame was not profiled");
  std::vector<ProfilerTask> tasks = r->gpuProfiler.GatherTasks();
  std::string names;
  uint32_t n = 0;
  for (const ProfilerTask &t : tasks) {
    if (n >= maxPasses) break;
    if (ms) ms[n] = float(t.GetLength() * 1e3);
    names += t.name;
    names += '\n';
    n++;
  }
  if (nameBuf && nameBufBytes) {
    std::strncpy(nameBuf, names.c_str(), nameBufBytes - 1);
    nameBuf[nameBufBytes - 1] = 0;
  }
  return int(n);
}

uint64_t lgh_allocated_bytes(lgh_renderer *r) { return r ? r->core->GetRenderGraph()->GetAllocatedBytes() : 0; }

} // extern "C"
