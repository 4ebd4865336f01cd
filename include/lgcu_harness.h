/*
 * lgcu_harness.h — C ABI of the headless frame harness (liblegit_cuda.so).
 *
 * The harness stands where the reference's application shell stands (src/main.cpp:104-320: create the renderer, recreate
 * the viewport resources, then per frame BeginFrame -> renderer->RenderFrame -> EndFrame -> RenderGraph::Execute): it owns
 * a legit_cuda::Core + SSVGIRenderer (the C++ mirror of the reference's rendergraph / renderer, host/legit_cuda/), the
 * rasterised scene buffers and a B8G8R8A8_SRGB "swapchain" image, and runs one frame per call on a CUDA stream.
 * Python (tests, bench.py) drives it through ctypes; a C++ application would use the legit_cuda headers directly.
 *
 * All calls enqueue on the renderer's stream and return; lgh_sync() waits. Host pointers given to the upload/download
 * calls should be page-locked for the copies to be asynchronous. Return value: 0 or a negative lgcu_status; the message
 * is available from lgh_last_error().
 */
#ifndef LGCU_HARNESS_H
#define LGCU_HARNESS_H

#include "lgcu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lgh_renderer lgh_renderer;

enum { LGH_MODE_PASS_GRANULAR = 0, LGH_MODE_FUSED = 1 };

const char *lgh_last_error(void);

/* stream: cudaStream_t to run on (NULL = a stream owned by the harness). The device is the calling thread's current one. */
lgh_renderer *lgh_create(uint32_t width, uint32_t height, void *stream);
void lgh_destroy(lgh_renderer *r);

/* camera / light as (position, vertAngle, horAngle) — src/Scene/Scene.h:20-34; defaults are the reference's (main.cpp:166-172) */
int lgh_set_camera(lgh_renderer *r, const float camPos[3], float camVertAngle, float camHorAngle, const float lightPos[3],
                   float lightVertAngle, float lightHorAngle);

/* scene upload (host -> device, async). rows [rowBegin,rowEnd) of the fragment buffer; hostFragments points at row 0. */
int lgh_upload_fragments(lgh_renderer *r, const lgcu_fragment *hostFragments, uint64_t hostPitchBytes, uint32_t rowBegin, uint32_t rowEnd);
int lgh_upload_objects(lgh_renderer *r, const lgcu_draw_call_data *hostObjects, uint32_t count);
int lgh_upload_light_depth(lgh_renderer *r, const float *hostDepth, uint32_t size);

/* Scene as the reference's Scene holds it (src/Scene/Scene.h:141-147, src/Scene/Mesh.h:244-262): vertex / index buffers, the draw
 * list (one entry per Scene::IterateObjects callback, firstTriangle filled by lgcu_raster_prepare_draws) and the per-object
 * constants. HOST arrays, copied asynchronously on the renderer's stream. After this call frames start from the mesh: "ShadowPass"
 * and "GBufferRasterPass" rasterise it on the device (lgcu_raster_shadow_map / lgcu_raster_gbuffer) and the fragment / light-depth
 * uploads are not used. A frame captured before the first lgh_upload_mesh, or before one that grew the scene, must be re-captured;
 * re-uploading a scene of the same size (per-frame object constants, animated vertices) keeps the captured frame valid.
 * lgh_use_mesh(r, 0) switches back to the pre-rasterised inputs. */
int lgh_upload_mesh(lgh_renderer *r, const lgcu_vertex *hostVertices, uint32_t nVertices, const uint32_t *hostIndices, uint32_t nIndices,
                    const lgcu_draw *hostDraws, uint32_t nDraws, const lgcu_draw_call_data *hostObjects, uint32_t nObjects);
int lgh_use_mesh(lgh_renderer *r, uint32_t enable);
/* DebugRenderer::RenderImageViews over the finished frame (SSVGIRenderer.h:344-350): off by default — the reference always draws
 * it, the parity tests and the bench look at the frame before it. Takes effect with the next lgh_render_frame / lgh_capture_frame. */
int lgh_set_debug_overlay(lgh_renderer *r, uint32_t enable);
/* Puts another B8G8R8A8_SRGB image of the viewport's size and the canonical layout behind the swapchain proxy — the reference's
 * swapchain view is an external image-view proxy too (LV/PresentQueue.h:106-110). deviceBase may be memory of a PEER GPU (CUDA IPC
 * mapping): a strip-sharded frame then composites by writing every strip straight into the presenting GPU's image. NULL restores the
 * renderer's own image. Not allowed while a captured frame exists (re-capture afterwards). */
int lgh_set_external_swapchain(lgh_renderer *r, void *deviceBase);

/* One frame: SSVGIRenderer::RenderFrame + RenderGraph::Execute. rows may be NULL (whole frame; required for pass-granular).
 * profile != 0 records per-pass GPU events (read them with lgh_get_profile after lgh_sync). */
int lgh_render_frame(lgh_renderer *r, uint32_t mode, int32_t denoiserRadius, uint32_t giFlags, const lgcu_rows *rows, uint32_t profile);

/* Part of a fused frame: only the passes of the selected stages are declared and run. A multi-GPU strip renderer runs the stages
 * one at a time and exchanges halo rows between them (legitengine_b200/multigpu.py). */
enum { LGH_STAGE_FRONT = 1, LGH_STAGE_CHAINS = 2, LGH_STAGE_GATHER = 4, LGH_STAGE_FINAL = 8, LGH_STAGE_ALL = 15 };
int lgh_render_stages(lgh_renderer *r, uint32_t mode, int32_t denoiserRadius, uint32_t giFlags, const lgcu_rows *rows, uint32_t stages);

/* CUDA-IPC plumbing of the peer-to-peer strip exchange (one process per GPU): export an image's allocation (valid after the first
 * frame), open / close a peer's handle in this process, and small zero-initialised device allocations for the flag words. */
int lgh_ipc_export_image(lgh_renderer *r, const char *name, unsigned char handle[64], uint64_t *bytes);
int lgh_ipc_export_ptr(void *devicePtr, unsigned char handle[64]);
int lgh_ipc_open(const unsigned char handle[64], void **devicePtr);
int lgh_ipc_close(void *devicePtr);
int lgh_device_alloc_zeroed(uint64_t bytes, void **devicePtr);
int lgh_device_free(void *devicePtr);

/* Capture the same frame into a CUDA graph once (after at least one lgh_render_frame with the same arguments has
 * allocated the images), then replay it with a single launch per frame. */
int lgh_capture_frame(lgh_renderer *r, uint32_t mode, int32_t denoiserRadius, uint32_t giFlags, const lgcu_rows *rows);
int lgh_replay_frame(lgh_renderer *r);
/* kernel nodes in the captured graph / passes declared by the last lgh_render_frame */
int lgh_captured_kernel_count(lgh_renderer *r);
int lgh_last_pass_count(lgh_renderer *r);

/* images by name: albedo emissive normal depthMoments blurredDepthMoments depthStencil directLight blurredDirectLight
 * shadowMap indirectLight denoisedIndirectLight swapchain. Valid after the first frame. */
int lgh_image_desc(lgh_renderer *r, const char *name, lgcu_image *out);
int lgh_download_image(lgh_renderer *r, const char *name, uint32_t level, void *host, uint64_t hostPitchBytes, uint32_t rowBegin, uint32_t rowEnd);

/* InterleaveBuilder on the rendergraph (src/Render/Common/InterleaveBuilder.h:14-80): Deinterleave(image -> a) then Interleave(a -> b)
 * on two transient images of the source's format and size; both results are copied to host memory (tightly packed rows of
 * hostPitchBytes) and the call waits for them. `srcName` is a single-level image of the last frame (see lgh_image_desc). */
int lgh_run_interleave(lgh_renderer *r, const char *srcName, uint32_t gridX, uint32_t gridY, void *hostDeinterleaved, void *hostRoundTrip,
                       uint64_t hostPitchBytes);

int lgh_sync(lgh_renderer *r);

/* Per-pass GPU times of the last profiled frame: names are written '\n'-separated into nameBuf, durations (ms) into ms.
 * Returns the number of passes (or a negative status). */
int lgh_get_profile(lgh_renderer *r, char *nameBuf, uint64_t nameBufBytes, float *ms, uint32_t maxPasses);

uint64_t lgh_allocated_bytes(lgh_renderer *r);

#ifdef __cplusplus
}
#endif
#endif
