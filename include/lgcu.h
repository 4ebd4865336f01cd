/*
 * lgcu.h — C ABI of the B200-native (sm_100a) replacement for LegitEngine's per-pixel lighting and
 * screen-space GI passes (the SSVGIRenderer hot path).
 *
 * Every entry point below replaces one SPIR-V full-screen pass (or a fused group of them) that the
 * reference dispatches through legit::RenderGraph::AddPass. Citations are reference paths relative to
 * the LegitEngine repository root (LV/ = dependencies/LegitVulkan/LegitVulkan/, SH/ = bin/data/Shaders/glsl/).
 *
 * Conventions
 *  - Plain C, plain pointers and sizes. No torch / C++ types cross this boundary.
 *  - All image pointers are DEVICE pointers. Parameter blocks (the reference's UBO structs) are HOST
 *    pointers and are consumed (copied into kernel arguments) before the call returns.
 *  - Every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *    stream). No host synchronisation, no allocation, no global state; safe to capture in a CUDA graph.
 *  - Return value: LGCU_OK (0) or a negative lgcu_status. Nothing throws across the boundary.
 *  - There is no CPU fallback: if no CUDA device / kernel image is usable the call returns
 *    LGCU_ERR_CUDA and lgcu_last_error() describes why.
 */
#ifndef LGCU_H
#define LGCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGCU_ABI_VERSION 1
#define LGCU_MAX_MIPS 16

typedef enum lgcu_status {
  LGCU_OK = 0,
  LGCU_ERR_INVALID_ARGUMENT = -1, /* null pointer, size mismatch between bound images, bad radius ... */
  LGCU_ERR_UNSUPPORTED_FORMAT = -2,
  LGCU_ERR_CUDA = -3,             /* launch / runtime failure; see lgcu_last_error()                  */
  LGCU_ERR_UNSUPPORTED = -4       /* a mode of the reference shader that has no live caller            */
} lgcu_status;

/* Formats: numeric values are the VkFormat enumerants the reference allocates its images with
 * (src/Render/Renderers/SSVGIRenderer.h:393-404, LV/Swapchain.h:108). */
typedef enum lgcu_format {
  LGCU_FORMAT_UNDEFINED = 0,
  LGCU_FORMAT_B8G8R8A8_SRGB = 50,
  LGCU_FORMAT_R16G16B16A16_SFLOAT = 97,
  LGCU_FORMAT_R32G32_SFLOAT = 103,
  LGCU_FORMAT_R32G32B32A32_SFLOAT = 109, /* extension: fp32 render target for un-quantised parity checks */
  LGCU_FORMAT_D32_SFLOAT = 126
} lgcu_format;

/* A resolved image view: what legit::RenderGraph::PassContext::GetImageView(id) hands to a pass
 * (LV/RenderGraph.h:421-451, LV/ImageView.h:21-31, LV/Image.h:90-108), flattened to a POD.
 * Memory layout: linear row-major texels, one contiguous block per mip level.
 *   texel (x, y) of IMAGE level l lives at  base + levelOffset[l] + y * levelPitch[l] + x * texelSize.
 * Level l has size (width >> l, height >> l) (integer floor; LV/RenderGraph.h:373-374).
 * The view covers image levels [baseMip, baseMip + mipCount); shader "lod 0" == image level baseMip. */
typedef struct lgcu_image {
  void *base;
  uint32_t format;        /* lgcu_format */
  uint32_t width, height; /* size of IMAGE level 0 */
  uint32_t imageMipCount; /* levels allocated in the image */
  uint32_t baseMip;       /* view sub-range */
  uint32_t mipCount;
  uint32_t reserved0;
  uint32_t reserved1;
  uint64_t levelOffset[LGCU_MAX_MIPS]; /* bytes from base */
  uint32_t levelPitch[LGCU_MAX_MIPS];  /* bytes per row   */
} lgcu_image;

/* Row range of the frame that this call computes: rows [y0, y1) of the BASE resolution (full-frame
 * coordinates are kept everywhere: pattern index, uv, clamp-to-edge all use the full image size).
 * NULL means the whole image. For passes that run on mip level l the range is scaled to
 * [y0 >> l, ceil(y1 / 2^l)) clipped to the level. Used by the multi-GPU strip sharding. */
typedef struct lgcu_rows {
  uint32_t y0, y1;
} lgcu_rows;

/* ---- parameter blocks: byte-identical to the reference's tightly packed UBO structs ------------------ */
#pragma pack(push, 1)
/* column-major 4x4, like glm::mat4: m[c*4 + r] */
typedef struct lgcu_mat4 { float m[16]; } lgcu_mat4;

/* SSVGIRenderer.h:425-431 (GBufferBuilderShader::DataBuffer), SH/Common/gBufferBuilder.frag:4-10 */
typedef struct lgcu_gbuffer_builder_data { lgcu_mat4 viewMatrix, projMatrix; float time, bla; } lgcu_gbuffer_builder_data;
/* SSVGIRenderer.h:33-38 (DrawCallDataBuffer), SH/Common/gBufferBuilder.frag:12-17 */
typedef struct lgcu_draw_call_data { lgcu_mat4 modelMatrix; float albedoColor[4]; float emissiveColor[4]; } lgcu_draw_call_data;
/* SSVGIRenderer.h:457-464, SH/Common/directLighting.frag:4-11 */
typedef struct lgcu_direct_lighting_data { lgcu_mat4 viewMatrix, projMatrix, lightViewMatrix, lightProjMatrix; float time; } lgcu_direct_lighting_data;
/* MipBuilder.h:250-253, SH/Common/mipLevelBuilder.frag:4-7 */
typedef struct lgcu_mip_level_builder_data { float filterType; } lgcu_mip_level_builder_data;
/* BlurBuilder.h:63-67, SH/Common/blurLayerBuilder.frag:4-8 */
typedef struct lgcu_blur_layer_builder_data { int32_t size[4]; int32_t radius; } lgcu_blur_layer_builder_data;
/* SSVGIRenderer.h:476-481, SH/SSVGI/indirectLighting.frag:4-9 */
typedef struct lgcu_indirect_lighting_data { lgcu_mat4 viewMatrix, projMatrix; float viewportExtent[4]; } lgcu_indirect_lighting_data;
/* SSVGIRenderer.h:492-498, SH/Common/denoiser.frag:4-10 */
typedef struct lgcu_denoiser_data { lgcu_mat4 viewMatrix, projMatrix; float viewportExtent[4]; int32_t radius; } lgcu_denoiser_data;
/* SSVGIRenderer.h:509-513, SH/Common/finalGatherer.frag:4-8 */
typedef struct lgcu_final_gatherer_data { lgcu_mat4 viewMatrix, projMatrix; } lgcu_final_gatherer_data;
/* InterleaveBuilder.h:99-123 (DeinterleaveShader / InterleaveShader ::ShaderDataBuffer), SH/Common/{interleave,deinterleave}.frag:4-8 */
typedef struct lgcu_interleave_data { int32_t gridSize[4]; int32_t viewportSize[4]; } lgcu_interleave_data;
/* DebugRenderer.h:84-87 (DebugRendererShader::QuadData), SH/Common/debugRenderer.vert:9-12: (min.x, min.y, max.x, max.y) in [0,1] screen units */
typedef struct lgcu_debug_quad_data { float minmax[4]; } lgcu_debug_quad_data;
#pragma pack(pop)

/* Per-pixel fragment attributes: the CUDA-side form of what the rasteriser hands to
 * SH/Common/gBufferBuilder.frag (interpolated vertWorldPos / vertWorldNormal from gBufferBuilder.vert:32-40,
 * the draw call the surviving fragment belongs to, and the depth-tested NDC z). 32 bytes per pixel,
 * row-major, `attribPitch` bytes per row. objectId == LGCU_NO_OBJECT marks an uncovered pixel, which keeps
 * the attachment clear values (LV/RenderGraph.h:466-469, 484-487). */
typedef struct lgcu_fragment {
  float worldPos[3];
  float worldNormal[3];
  uint32_t objectId;
  float ndcDepth;
} lgcu_fragment;
#define LGCU_NO_OBJECT 0xFFFFFFFFu

/* Clear values of the G-buffer pass (default colour clear (1, .5, 0, 1), depth 1; LV/RenderGraph.h:469, 487) */
typedef struct lgcu_clear_values { float color[4]; float depth; } lgcu_clear_values;

/* ---- library ------------------------------------------------------------------------------------------ */
int lgcu_abi_version(void);
/* Human-readable description of the last error raised on the calling thread ("" if none). */
const char *lgcu_last_error(void);
/* Bytes per texel of a format, 0 if unsupported. */
uint32_t lgcu_format_texel_size(uint32_t format);
/* Fills width/height/levelOffset/levelPitch for a `mips`-level image in the library's canonical layout
 * (pitch rounded up to 128 B, level blocks to 256 B) and returns the allocation size in bytes. `base` and the
 * view range are left for the caller. Host-only helper; does not touch the device. */
uint64_t lgcu_image_layout(lgcu_image *img, uint32_t format, uint32_t width, uint32_t height, uint32_t mips);

/* ---- one entry point per reference pass ------------------------------------------------------------- */

/* K1 "GBufferPass" fragment stage: SH/Common/gBufferBuilder.frag:28-38, pass SSVGIRenderer.h:107-158.
 * objects: DEVICE array of nObjects DrawCallData blocks (one per draw call, indexed by lgcu_fragment.objectId).
 * Writes albedo, emissive, normal (RGBA16F), depthMoments level 0 (RG32F) and depthStencil (D32F). */
int lgcu_gbuffer_resolve(const lgcu_gbuffer_builder_data *params, const lgcu_draw_call_data *objects, uint32_t nObjects,
                         const lgcu_fragment *fragments, uint64_t fragmentPitchBytes, const lgcu_clear_values *clear,
                         const lgcu_image *albedo, const lgcu_image *emissive, const lgcu_image *normal,
                         const lgcu_image *depthMoments, const lgcu_image *depthStencil, const lgcu_rows *rows, void *stream);

/* K2 "LightPass": SH/Common/directLighting.frag:45-83, pass SSVGIRenderer.h:161-205. */
int lgcu_direct_light(const lgcu_direct_lighting_data *params, const lgcu_image *albedo, const lgcu_image *emissive,
                      const lgcu_image *normal, const lgcu_image *depthStencil, const lgcu_image *shadowMap,
                      const lgcu_image *directLight, const lgcu_rows *rows, void *stream);

/* K3 "MipBuilderPass" (one level): SH/Common/mipLevelBuilder.frag:17-43, driver MipBuilder.h:142-181.
 * src and dst are single-level views (level l-1 and l). filterType < 0.5 = Avg (:23-28, the live path);
 * filterType >= 0.5 = Depth (:29-42; MipBuilder::FilterTypes::Depth, MipBuilder.h:136-141, 168): per 2x2 block
 * (min of .x, max of .y, sum((y - x) * z) / (max - min) / 4, 0). */
int lgcu_mip_level(const lgcu_mip_level_builder_data *params, const lgcu_image *srcLevel, const lgcu_image *dstLevel,
                   const lgcu_rows *rows, void *stream);

/* K4 "BlurPass" (one level): SH/Common/blurLayerBuilder.frag:17-35, driver BlurBuilder.h:14-46.
 * radius 0 = copy, radius 2 = asymmetric 4x4 box (-2..+1) with clamp-to-edge. Other radii: generic loop. */
int lgcu_blur_level(const lgcu_blur_layer_builder_data *params, const lgcu_image *srcLevel, const lgcu_image *dstLevel,
                    const lgcu_rows *rows, void *stream);

/* K5 "IndirectLightPass": SH/SSVGI/indirectLighting.frag:114-272, pass SSVGIRenderer.h:224-263.
 * flags: LGCU_GI_* below. */
int lgcu_gi_gather(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight,
                   const lgcu_image *blurredDepthMoments, const lgcu_image *normal, const lgcu_image *depthStencil,
                   const lgcu_image *indirectLight, uint32_t flags, const lgcu_rows *rows, void *stream);
#define LGCU_GI_DEFAULT 0u
#define LGCU_GI_STRICT 1u /* shader-order arithmetic with libm-grade sin/cos/atan/pow/log (parity variant) */

/* K5 with the quad-packed depth pyramid (a private acceleration structure of this library):
 *   bytes   = lgcu_gather_scratch_bytes(width, height, mips)      size of the scratch buffer for a pyramid of that shape
 *   lgcu_gi_gather_pack(...)    builds the scratch from blurredDepthMoments (after the blur passes, before the gather)
 *   lgcu_gi_gather_packed(...)  the same pass as lgcu_gi_gather(LGCU_GI_DEFAULT), reading depth through the scratch
 * scratch is DEVICE memory owned by the caller (16-byte aligned); the image arguments are those of lgcu_gi_gather. With a
 * row strip, pack rebuilds the rows the strip's march can reach. */
uint64_t lgcu_gather_scratch_bytes(uint32_t width, uint32_t height, uint32_t mips);
int lgcu_gi_gather_pack(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight,
                        const lgcu_image *blurredDepthMoments, const lgcu_image *normal, const lgcu_image *depthStencil,
                        const lgcu_image *indirectLight, void *scratch, uint64_t scratchBytes, const lgcu_rows *rows, void *stream);
int lgcu_gi_gather_packed(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight,
                          const lgcu_image *blurredDepthMoments, const lgcu_image *normal, const lgcu_image *depthStencil,
                          const lgcu_image *indirectLight, const void *scratch, uint64_t scratchBytes, const lgcu_rows *rows,
                          void *stream);

/* K6 "DenoiserPass": SH/Common/denoiser.frag:72-185, pass SSVGIRenderer.h:266-302.
 * `depthMoments` is what the reference binds to the shader's depthStencilSampler (SSVGIRenderer.h:293). */
int lgcu_denoise(const lgcu_denoiser_data *params, const lgcu_image *noisy, const lgcu_image *normal,
                 const lgcu_image *depthMoments, const lgcu_image *denoised, const lgcu_rows *rows, void *stream);

/* K7 "GatheringPass": SH/Common/finalGatherer.frag:42-60, pass SSVGIRenderer.h:305-342. */
int lgcu_final_gather(const lgcu_final_gatherer_data *params, const lgcu_image *directLight,
                      const lgcu_image *blurredDirectLight, const lgcu_image *albedo, const lgcu_image *indirectLight,
                      const lgcu_image *swapchain, const lgcu_rows *rows, void *stream);

/* ---- fused groups (same outputs as the passes they replace, fewer trips through HBM) ------------------ */

/* K1+K2: G-buffer resolve and direct lighting in one pass over the fragments. Writes everything K1 writes
 * plus directLight level 0. */
int lgcu_gbuffer_direct_light(const lgcu_gbuffer_builder_data *gparams, const lgcu_direct_lighting_data *lparams,
                              const lgcu_draw_call_data *objects, uint32_t nObjects, const lgcu_fragment *fragments,
                              uint64_t fragmentPitchBytes, const lgcu_clear_values *clear, const lgcu_image *albedo,
                              const lgcu_image *emissive, const lgcu_image *normal, const lgcu_image *depthMoments,
                              const lgcu_image *depthStencil, const lgcu_image *shadowMap, const lgcu_image *directLight,
                              const lgcu_rows *rows, void *stream);

/* K3+K4 for one chain: builds levels 1..mips-1 of `chain` from its level 0 (MipBuilder::BuildMips, Avg filter)
 * and writes blurred[l] = blur(chain[l]) with radius 0 at level 0 and `radius` at levels >= 1
 * (SSVGIRenderer.h:207-221). `chain` and `blurred` are whole-image views with identical format and size. */
int lgcu_mip_blur_chain(const lgcu_image *chain, const lgcu_image *blurred, int32_t radius, const lgcu_rows *rows,
                        void *stream);

/* The fused frame front: K1 + K2 + the radius-0 blur passes of level 0 + mip levels 1..LGCU_FRONT_MIP_LEVELS of both chains, in
 * one pass over the fragments. Writes everything lgcu_gbuffer_direct_light writes, plus blurredDirectLight / blurredDepthMoments
 * level 0 (= directLight / depthMoments level 0) and levels 1..4 of directLight and depthMoments. The four chain arguments are
 * whole-image views (MippedProxy::imageViewProxy). A row strip must start on a multiple of 16 rows and end on one (or at the image
 * bottom). Follow it with lgcu_frame_chains. */
#define LGCU_FRONT_MIP_LEVELS 4
int lgcu_frame_front(const lgcu_gbuffer_builder_data *gparams, const lgcu_direct_lighting_data *lparams,
                     const lgcu_draw_call_data *objects, uint32_t nObjects, const lgcu_fragment *fragments,
                     uint64_t fragmentPitchBytes, const lgcu_clear_values *clear, const lgcu_image *albedo,
                     const lgcu_image *emissive, const lgcu_image *normal, const lgcu_image *depthMoments,
                     const lgcu_image *depthStencil, const lgcu_image *shadowMap, const lgcu_image *directLight,
                     const lgcu_image *blurredDirectLight, const lgcu_image *blurredDepthMoments, const lgcu_rows *rows,
                     void *stream);

/* The rest of K3 + K4 after lgcu_frame_front, in one launch: blur (radius) of levels >= 1 of both chains and mip levels above
 * LGCU_FRONT_MIP_LEVELS (built and blurred by one thread-block cluster per chain, in the same launch). With a row strip, the blur reads up to `radius` rows of each
 * level outside the strip (they must be present) and the levels above LGCU_FRONT_MIP_LEVELS are built whole from level 4. */
int lgcu_frame_chains(const lgcu_image *directLight, const lgcu_image *blurredDirectLight, const lgcu_image *depthMoments,
                      const lgcu_image *blurredDepthMoments, int32_t radius, const lgcu_rows *rows, void *stream);

/* K6(radius 0)+K7: denoised = noisy; swapchain = directLight + denoised * albedo. */
int lgcu_denoise_final_gather(const lgcu_denoiser_data *dparams, const lgcu_final_gatherer_data *fparams,
                              const lgcu_image *noisy, const lgcu_image *normal, const lgcu_image *depthMoments,
                              const lgcu_image *denoised, const lgcu_image *directLight,
                              const lgcu_image *blurredDirectLight, const lgcu_image *albedo,
                              const lgcu_image *swapchain, const lgcu_rows *rows, void *stream);

/* ---- the passes either side of the hot path ---------------------------------------------------------
 * Interleaved rendering (InterleaveBuilder.h:14-80): a gridSize.x x gridSize.y pattern of pixels is regrouped into gridSize
 * contiguous sub-images of size viewportSize / gridSize (integer division) and back. Both are pure index permutations of
 * texels (bit-exact); source and destination must have the same format (RGBA16F, RG32F or RGBA32F) and the same size =
 * viewportSize, with viewportSize >= gridSize >= 1 per axis. rows = destination rows.
 *   lgcu_deinterleave: SH/Common/deinterleave.frag:17-29  dst(p) = src(InterleavePixel(p)),
 *                      InterleavePixel(p) = (p % (viewport / grid)) * grid + p / (viewport / grid)
 *   lgcu_interleave:   SH/Common/interleave.frag:16-28    dst(p) = src(DeinterleavePixel(p)),
 *                      DeinterleavePixel(p) = (p % grid) * (viewport / grid) + p / grid
 * (The reference's InterleaveBuilder::Interleave binds the de-interleave program by mistake, InterleaveBuilder.h:60-65; this
 * entry point is the shipped interleave.frag.spv, i.e. what the builder means to run.) */
int lgcu_deinterleave(const lgcu_interleave_data *params, const lgcu_image *interleaved, const lgcu_image *deinterleaved,
                      const lgcu_rows *rows, void *stream);
int lgcu_interleave(const lgcu_interleave_data *params, const lgcu_image *deinterleaved, const lgcu_image *interleaved,
                    const lgcu_rows *rows, void *stream);

/* "DebugInfoPass" (DebugRenderer.h:13-63, SH/Common/debugRenderer.vert:15-24, debugRenderer.frag:11-15): one textured quad
 * covering [minmax.xy, minmax.zw] of the target (screen units), loadOp eLoad, opaque: every target pixel whose centre lies in
 * the quad receives texture(src, t), t = (pixel centre / target size - min) / (max - min), sampled with the pass's sampler
 * (clamp-to-edge, linear min/mag, nearest mip; single-level source views as on the live path, SSVGIRenderer.h:344-350) and
 * written with the target's format conversion (B8G8R8A8_SRGB swapchain, or RGBA16F / RGBA32F). Pixels outside the quad keep
 * their contents. The renderer issues one call per debug view (tile layout DebugRenderer.h:27-57). rows = target rows. */
int lgcu_debug_overlay(const lgcu_debug_quad_data *params, const lgcu_image *src, const lgcu_image *target, const lgcu_rows *rows,
                       void *stream);

/* ---- rasterisation front end: "ShadowPass" and the raster half of "GBufferPass"  --------------------
 * The reference draws every scene object with drawIndexed through the fixed-function rasteriser (SSVGIRenderer.h:63-104 shadow
 * map, :107-158 G-buffer; vertex stages SH/Common/shadowmapBuilder.vert:32-40 and SH/Common/gBufferBuilder.vert:32-40; pipeline
 * state LV/Pipeline.h:6-12, 148-178: fill, no culling, depth test LESS + write, depth clamp off, one sample per pixel; viewport =
 * render area, depth range [0,1], LV/RenderPassCache.h:94-101). These entry points do the same on the CUDA device and hand the
 * result to the fragment stage: the shadow map image, and the lgcu_fragment buffer lgcu_gbuffer_resolve / lgcu_frame_front consume.
 * The rasterisation rule (pixel-centre sampling, top-left fill rule, perspective-correct interpolation, near/far clipping per
 * fragment, first-drawn-wins on equal depth) is stated in DESIGN.md §8. */
#pragma pack(push, 1)
/* src/Scene/Mesh.h:209-216 (MeshData::Vertex, 32 bytes; vertex declaration :263-271) */
typedef struct lgcu_vertex { float pos[3]; float normal[3]; float uv[2]; } lgcu_vertex;
/* SSVGIRenderer.h:439-447 (ShadowmapBuilderShader::DataBuffer), SH/Common/shadowmapBuilder.vert:8-12 */
typedef struct lgcu_shadowmap_builder_data { lgcu_mat4 lightViewMatrix, lightProjMatrix; } lgcu_shadowmap_builder_data;
#pragma pack(pop)
/* One drawIndexed(indexCount, 1, firstIndex, vertexOffset, 0) with DrawCallData = objects[objectId] bound
 * (Scene::IterateObjects callback, SSVGIRenderer.h:84-102, 138-156). firstTriangle = number of triangles drawn before this call
 * in the frame (draw order decides equal-depth fragments); lgcu_raster_prepare_draws fills it. */
typedef struct lgcu_draw { uint32_t firstIndex, indexCount, vertexOffset, objectId, firstTriangle, reserved[3]; } lgcu_draw;
/* The scene as the reference's Scene holds it: vertex / index buffers, the per-object constants and the draw list. All four arrays
 * are DEVICE memory; they are read by the kernels, never by the host. */
typedef struct lgcu_mesh_scene {
  const lgcu_vertex *vertices;
  const uint32_t *indices;
  const lgcu_draw *draws;
  const lgcu_draw_call_data *objects;
  uint32_t nVertices, nIndices, nDraws, nObjects;
  uint32_t nTriangles; /* sum of indexCount / 3 over the draws (= what lgcu_raster_prepare_draws returned) */
} lgcu_mesh_scene;
/* HOST helper: fills draws[i].firstTriangle (draws in HOST memory, before upload) and returns the total triangle count. */
uint32_t lgcu_raster_prepare_draws(lgcu_draw *hostDraws, uint32_t nDraws);
/* Bytes of DEVICE scratch (256-byte aligned) the raster calls need for a scene of nTriangles on a width x height target. */
uint64_t lgcu_raster_scratch_bytes(uint32_t nTriangles, uint32_t width, uint32_t height);
/* K0 "ShadowPass": depth-only rasterisation from the light into shadowMap (D32F, cleared to 1). */
int lgcu_raster_shadow_map(const lgcu_shadowmap_builder_data *params, const lgcu_mesh_scene *scene, void *scratch,
                           uint64_t scratchBytes, const lgcu_image *shadowMap, void *stream);
/* Raster half of K1 "GBufferPass": visibility + interpolated vertWorldPos / vertWorldNormal per pixel, written as the
 * lgcu_fragment buffer (width x height, fragmentPitchBytes per row, DEVICE memory) that the fragment-stage entries read. */
int lgcu_raster_gbuffer(const lgcu_gbuffer_builder_data *params, const lgcu_mesh_scene *scene, void *scratch,
                        uint64_t scratchBytes, uint32_t width, uint32_t height, lgcu_fragment *fragments,
                        uint64_t fragmentPitchBytes, const lgcu_rows *rows, void *stream);

/* ---- peer-to-peer row exchange for the strip-sharded frame (no counterpart in the single-GPU reference) -------
 * All pointers are DEVICE addresses valid in the calling process; a peer GPU's memory is addressed through a CUDA-IPC mapping, and
 * loads / stores on it travel over NVLink. Everything is enqueue-only and CUDA-graph capturable. */
typedef struct lgcu_row_copy { const void *src; void *dst; uint64_t bytes; } lgcu_row_copy; /* 16-byte aligned, multiple of 16 */
/* Copies `count` contiguous slabs in one kernel launch per 64 slabs (rows of an image level are contiguous, so a halo is a slab). */
int lgcu_copy_rows(const lgcu_row_copy *copies, uint32_t count, void *stream);
/* *frameCounter += 1 on the stream (device memory; keeps signal / wait values correct under graph replay). */
int lgcu_frame_counter_bump(uint32_t *frameCounter, void *stream);
/* *flags[i] = *frameCounter for i < count (<= 32), after a system-scope fence: "my work enqueued before this call is done". */
int lgcu_signal_flags(uint32_t *const *flags, uint32_t count, const uint32_t *frameCounter, void *stream);
/* Blocks the stream (a one-warp spinning kernel) until *flags[i] >= *frameCounter - lag for every i < count. */
int lgcu_wait_flags(uint32_t *const *flags, uint32_t count, const uint32_t *frameCounter, uint32_t lag, void *stream);
/* One exchange step of the strip protocol in ONE launch — signal, wait, pull and acknowledge fused with the copy over peer memory:
 *   bump != 0        : *frameCounter += 1 first (then no copies are allowed: the step opens a frame)
 *   signalBefore     : flags set to the frame number before anything else ("my previous stage is done")
 *   wait, lag        : every CTA spins until the flags reach frame - lag, then
 *   copies           : the slabs are pulled (or pushed) by the whole grid, and
 *   signalAfter      : the LAST CTA to finish sets these flags ("I am done reading your rows"); doneCounter is a zero-initialised
 *                      device word owned by the caller that the kernel uses to find that CTA (it leaves it at zero).
 * Lists hold at most 32 flags / 64 slabs. Replaces a sequence of up to four of the calls above (and their launch gaps). */
typedef struct lgcu_exchange_desc {
  uint32_t *const *signalBefore; uint32_t signalBeforeCount;
  uint32_t *const *wait; uint32_t waitCount; uint32_t lag;
  const lgcu_row_copy *copies; uint32_t copyCount;
  uint32_t *const *signalAfter; uint32_t signalAfterCount;
  uint32_t *frameCounter; uint32_t *doneCounter; uint32_t bump;
} lgcu_exchange_desc;
int lgcu_exchange(const lgcu_exchange_desc *desc, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LGCU_H */
