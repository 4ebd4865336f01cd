// ref_pass.cpp — TEST INFRASTRUCTURE (oracle/_ref reference arm; never linked into the product path).
//
// Compiled once per hot-path pass with -DREF_PASS=<n> -DREF_PASS_NAME=<name> -DREF_GEN_FILE="<generated .cpp>".
// Each variant includes the C++ that the reference's vendored SPIRV-Cross emitted from the reference's shipped
// SPIR-V module and wraps it in a per-pixel driver that plays the role of the rasteriser / ROP:
// it sets gl_FragCoord and the interpolated quad coordinate, binds the UBO and the image views exactly as the
// reference's record lambdas do (file:line cited per pass), invokes the module and stores the outputs with the
// render-target conversions of SURVEY.md Appendix B. Signatures mirror include/lgcu.h minus the stream.
#include "ref_prelude.hpp"

#include REF_GEN_FILE

#include <omp.h>

using spirv_cross::RefSampler2D;

static inline void ref_store(const lgcu_image *img, int x, int y, const glm::vec4 &v) {
  const float f[4] = {v.x, v.y, v.z, v.w};
  orc_store_texel(img, 0, x, y, f);
}

#if REF_PASS == 1
// GBufferPass, fragment stage: bin/data/Shaders/spirv/Common/gBufferBuilder.frag.spv
// bindings as in src/Render/Renderers/SSVGIRenderer.h:107-158 (set 0 = GBufferBuilderData, set 1 = DrawCallData);
// attachment order :108-114; clear values LV/RenderGraph.h:466-469, 484-487; depth test/write LV/Pipeline.h.
extern "C" int ref_gbuffer_resolve(const lgcu_gbuffer_builder_data *params, const lgcu_draw_call_data *objects,
                                   uint32_t nObjects, const lgcu_fragment *fragments, uint64_t fragmentPitchBytes,
                                   const lgcu_clear_values *clear, const lgcu_image *albedoImg, const lgcu_image *emissiveImg,
                                   const lgcu_image *normalImg, const lgcu_image *depthMomentsImg,
                                   const lgcu_image *depthStencilImg, const lgcu_rows *rows) {
  int w, h, y0, y1;
  orc_level_size(albedoImg, 0, &w, &h);
  ref_row_range(rows, 0, h, &y0, &y1);
  int bad = 0;
#pragma omp parallel
  {
    RefShaderInstance s;
    glm::vec3 worldPos, worldNormal;
    glm::vec2 uv(0.0f);
    glm::vec4 oAlbedo, oEmissive, oNormal, oDepth;
    s.resource(0, 0, (void *)params);
    s.input(0, &worldPos, sizeof(worldPos));
    s.input(1, &worldNormal, sizeof(worldNormal));
    s.input(2, &uv, sizeof(uv));
    s.output(0, &oAlbedo, sizeof(oAlbedo));
    s.output(1, &oEmissive, sizeof(oEmissive));
    s.output(2, &oNormal, sizeof(oNormal));
    s.output(3, &oDepth, sizeof(oDepth));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++) {
      const lgcu_fragment *row = (const lgcu_fragment *)((const uint8_t *)fragments + (uint64_t)y * fragmentPitchBytes);
      for (int x = 0; x < w; x++) {
        const lgcu_fragment &f = row[x];
        if (f.objectId == LGCU_NO_OBJECT) {
          glm::vec4 c(clear->color[0], clear->color[1], clear->color[2], clear->color[3]);
          ref_store(albedoImg, x, y, c);
          ref_store(emissiveImg, x, y, c);
          ref_store(normalImg, x, y, c);
          ref_store(depthMomentsImg, x, y, c);
          ref_store(depthStencilImg, x, y, glm::vec4(clear->depth));
          continue;
        }
        if (f.objectId >= nObjects) {
          bad = 1;
          continue;
        }
        s.resource(1, 0, (void *)&objects[f.objectId]);
        worldPos = glm::vec3(f.worldPos[0], f.worldPos[1], f.worldPos[2]);
        worldNormal = glm::vec3(f.worldNormal[0], f.worldNormal[1], f.worldNormal[2]);
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(albedoImg, x, y, oAlbedo);
        ref_store(emissiveImg, x, y, oEmissive);
        ref_store(normalImg, x, y, oNormal);
        ref_store(depthMomentsImg, x, y, oDepth);
        ref_store(depthStencilImg, x, y, glm::vec4(f.ndcDepth));
      }
    }
  }
  return bad ? LGCU_ERR_INVALID_ARGUMENT : LGCU_OK;
}
#endif

#if REF_PASS == 2
// LightPass: spirv/Common/directLighting.frag.spv; bindings SSVGIRenderer.h:174-203
// (binding 1 albedoImg, 2 emissiveImg, 3 normalImg, 4 depthStencilImg, 5 shadowmap — directLighting.frag:13-17).
extern "C" int ref_direct_light(const lgcu_direct_lighting_data *params, const lgcu_image *albedoImg,
                                const lgcu_image *emissiveImg, const lgcu_image *normalImg, const lgcu_image *depthStencilImg,
                                const lgcu_image *shadowMapImg, const lgcu_image *directLightImg, const lgcu_rows *rows) {
  int w, h, y0, y1;
  orc_level_size(directLightImg, 0, &w, &h);
  ref_row_range(rows, 0, h, &y0, &y1);
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sAlbedo(albedoImg), sEmissive(emissiveImg), sNormal(normalImg), sDepth(depthStencilImg);
    spirv_cross::sampler2DShadow sShadow{shadowMapImg};
    glm::vec4 out;
    s.resource(0, 0, (void *)params);
    s.resource(0, 1, &sAlbedo);
    s.resource(0, 2, &sEmissive);
    s.resource(0, 3, &sNormal);
    s.resource(0, 4, &sDepth);
    s.resource(0, 5, &sShadow);
    s.bindScreenCoord(0);
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++)
      for (int x = 0; x < w; x++) {
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(directLightImg, x, y, out);
      }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 3
// MipBuilderPass (one level): spirv/Common/mipLevelBuilder.frag.spv; MipBuilder.h:142-181
// (render area = destination level size :148-149, 158; source = previous level view :153, 175).
extern "C" int ref_mip_level(const lgcu_mip_level_builder_data *params, const lgcu_image *srcLevelImg,
                             const lgcu_image *dstLevelImg, const lgcu_rows *rows) {
  int w, h, y0, y1;
  orc_level_size(dstLevelImg, 0, &w, &h);
  ref_row_range(rows, dstLevelImg->baseMip, h, &y0, &y1);
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sSrc(srcLevelImg);
    glm::vec4 out;
    s.resource(0, 0, (void *)params);
    s.resource(0, 1, &sSrc);
    s.bindScreenCoord(0);
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++)
      for (int x = 0; x < w; x++) {
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(dstLevelImg, x, y, out);
      }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 4
// BlurPass (one level): spirv/Common/blurLayerBuilder.frag.spv; BlurBuilder.h:14-46.
extern "C" int ref_blur_level(const lgcu_blur_layer_builder_data *params, const lgcu_image *srcLevelImg,
                              const lgcu_image *dstLevelImg, const lgcu_rows *rows) {
  int w, h, y0, y1;
  orc_level_size(dstLevelImg, 0, &w, &h);
  ref_row_range(rows, dstLevelImg->baseMip, h, &y0, &y1);
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sSrc(srcLevelImg);
    glm::vec4 out;
    s.resource(0, 0, (void *)params);
    s.resource(0, 1, &sSrc);
    s.bindScreenCoord(0);
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++)
      for (int x = 0; x < w; x++) {
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(dstLevelImg, x, y, out);
      }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 5
// IndirectLightPass: spirv/SSVGI/indirectLighting.frag.spv; bindings SSVGIRenderer.h:235-261
// (1 blurredDirectLightImg, 2 blurredDepthMomentsImg, 3 normalImg, 4 depthStencilImg — indirectLighting.frag:11-14).
// `flags` is accepted for signature parity with lgcu_gi_gather and ignored.
extern "C" int ref_gi_gather(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLightImg,
                             const lgcu_image *blurredDepthMomentsImg, const lgcu_image *normalImg,
                             const lgcu_image *depthStencilImg, const lgcu_image *indirectLightImg, uint32_t flags,
                             const lgcu_rows *rows) {
  (void)flags;
  int w, h, y0, y1;
  orc_level_size(indirectLightImg, 0, &w, &h);
  ref_row_range(rows, 0, h, &y0, &y1);
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sLight(blurredDirectLightImg), sMoments(blurredDepthMomentsImg), sNormal(normalImg), sDepth(depthStencilImg);
    glm::vec4 out;
    s.resource(0, 0, (void *)params);
    s.resource(0, 1, &sLight);
    s.resource(0, 2, &sMoments);
    s.resource(0, 3, &sNormal);
    s.resource(0, 4, &sDepth);
    s.bindScreenCoord(0);
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 2)
    for (int y = y0; y < y1; y++)
      for (int x = 0; x < w; x++) {
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(indirectLightImg, x, y, out);
      }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 6
// DenoiserPass: spirv/Common/denoiser.frag.spv; bindings SSVGIRenderer.h:276-300
// (1 noisyImg = indirectLightImg, 2 normalImg, 3 "depthStencilSampler" = depthMomentsImg :293 — denoiser.frag:12-14).
extern "C" int ref_denoise(const lgcu_denoiser_data *params, const lgcu_image *noisyImg, const lgcu_image *normalImg,
                           const lgcu_image *depthMomentsImg, const lgcu_image *denoisedImg, const lgcu_rows *rows) {
  int w, h, y0, y1;
  orc_level_size(denoisedImg, 0, &w, &h);
  ref_row_range(rows, 0, h, &y0, &y1);
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sNoisy(noisyImg), sNormal(normalImg), sDepth(depthMomentsImg);
    glm::vec4 out(0.0f);
    s.resource(0, 0, (void *)params);
    s.resource(0, 1, &sNoisy);
    s.resource(0, 2, &sNormal);
    s.resource(0, 3, &sDepth);
    s.bindScreenCoord(0);
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++)
      for (int x = 0; x < w; x++) {
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(denoisedImg, x, y, out);
      }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 7
// GatheringPass: spirv/Common/finalGatherer.frag.spv; bindings SSVGIRenderer.h:316-340
// (1 directLightImg, 2 blurredDirectLightImg, 3 albedoImg, 4 indirectLightImg = denoisedIndirectLight — finalGatherer.frag:10-13).
extern "C" int ref_final_gather(const lgcu_final_gatherer_data *params, const lgcu_image *directLightImg,
                                const lgcu_image *blurredDirectLightImg, const lgcu_image *albedoImg,
                                const lgcu_image *indirectLightImg, const lgcu_image *swapchainImg, const lgcu_rows *rows) {
  int w, h, y0, y1;
  orc_level_size(swapchainImg, 0, &w, &h);
  ref_row_range(rows, 0, h, &y0, &y1);
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sDirect(directLightImg), sBlurred(blurredDirectLightImg), sAlbedo(albedoImg), sIndirect(indirectLightImg);
    glm::vec4 out;
    s.resource(0, 0, (void *)params);
    s.resource(0, 1, &sDirect);
    s.resource(0, 2, &sBlurred);
    s.resource(0, 3, &sAlbedo);
    s.resource(0, 4, &sIndirect);
    s.bindScreenCoord(0);
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++)
      for (int x = 0; x < w; x++) {
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(swapchainImg, x, y, out);
      }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 8 || REF_PASS == 9
// Vertex stage of GBufferPass (8: spirv/Common/gBufferBuilder.vert.spv) and ShadowPass (9: spirv/Common/shadowmapBuilder.vert.spv):
// set 0 binding 0 = GBufferBuilderData / ShadowmapBuilderData, set 1 binding 0 = DrawCallData (SSVGIRenderer.h:72-92, 125-146);
// inputs = the vertex declaration of src/Scene/Mesh.h:263-271. out = n x {vertWorldPos[3], vertWorldNormal[3], gl_Position[4]}.
// Used to pin the vertex stage of oracle/raster_oracle.c.
#if REF_PASS == 8
extern "C" int ref_gbuffer_vertex_stage(const lgcu_draw_call_data *drawCall, const void *shaderData, const lgcu_vertex *vertices, uint32_t n, float *out) {
#else
extern "C" int ref_shadowmap_vertex_stage(const lgcu_draw_call_data *drawCall, const void *shaderData, const lgcu_vertex *vertices, uint32_t n, float *out) {
#endif
  const spirv_cross_interface *iface = spirv_cross_get_interface();
  spirv_cross_shader_t *sh = iface->construct();
  glm::vec4 position;
  glm::vec3 pos, normal, worldPos, worldNormal;
  glm::vec2 uv, outUv;
  void *p0 = (void *)shaderData, *p1 = (void *)drawCall;
  spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_POSITION, &position, sizeof(position));
  spirv_cross_set_resource(sh, 0, 0, &p0, sizeof(p0));
  spirv_cross_set_resource(sh, 1, 0, &p1, sizeof(p1));
  spirv_cross_set_stage_input(sh, 0, &pos, sizeof(pos));
  spirv_cross_set_stage_input(sh, 1, &normal, sizeof(normal));
  spirv_cross_set_stage_input(sh, 2, &uv, sizeof(uv));
  spirv_cross_set_stage_output(sh, 0, &worldPos, sizeof(worldPos));
  spirv_cross_set_stage_output(sh, 1, &worldNormal, sizeof(worldNormal));
  spirv_cross_set_stage_output(sh, 2, &outUv, sizeof(outUv));
  for (uint32_t i = 0; i < n; i++) {
    pos = glm::vec3(vertices[i].pos[0], vertices[i].pos[1], vertices[i].pos[2]);
    normal = glm::vec3(vertices[i].normal[0], vertices[i].normal[1], vertices[i].normal[2]);
    uv = glm::vec2(vertices[i].uv[0], vertices[i].uv[1]);
    iface->invoke(sh);
    float *o = out + 10 * i;
    o[0] = worldPos.x; o[1] = worldPos.y; o[2] = worldPos.z;
    o[3] = worldNormal.x; o[4] = worldNormal.y; o[5] = worldNormal.z;
    o[6] = position.x; o[7] = position.y; o[8] = position.z; o[9] = position.w;
  }
  iface->destruct(sh);
  return LGCU_OK;
}
#endif

#if REF_PASS == 10 || REF_PASS == 11
// Interleaved rendering: 10 = spirv/Common/deinterleave.frag.spv (InterleaveBuilder::Deinterleave, InterleaveBuilder.h:14-45),
// 11 = spirv/Common/interleave.frag.spv (what InterleaveBuilder::Interleave means to run, :47-80). Set 0: binding 0 = the UBO
// (gridSize, viewportSize), binding 1 = the source image; render area = the view's size (:16, :49).
#if REF_PASS == 10
extern "C" int ref_deinterleave(const lgcu_interleave_data *params, const lgcu_image *srcImg, const lgcu_image *dstImg, const lgcu_rows *rows) {
#else
extern "C" int ref_interleave(const lgcu_interleave_data *params, const lgcu_image *srcImg, const lgcu_image *dstImg, const lgcu_rows *rows) {
#endif
  int w, h, sw, sh, y0, y1;
  orc_level_size(dstImg, 0, &w, &h);
  orc_level_size(srcImg, 0, &sw, &sh);
  // same argument contract as include/lgcu.h (keeps every texelFetch of the module in bounds)
  if (sw != w || sh != h || srcImg->format != dstImg->format || params->viewportSize[0] != w || params->viewportSize[1] != h || params->gridSize[0] < 1 ||
      params->gridSize[1] < 1 || params->gridSize[0] > w || params->gridSize[1] > h)
    return LGCU_ERR_INVALID_ARGUMENT;
  ref_row_range(rows, 0, h, &y0, &y1);
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sSrc(srcImg);
    glm::vec4 out;
    s.resource(0, 0, (void *)params);
    s.resource(0, 1, &sSrc);
    s.bindScreenCoord(0);
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++)
      for (int x = 0; x < w; x++) {
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(dstImg, x, y, out);
      }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 12
// One quad of DebugInfoPass: fragment stage = spirv/Common/debugRenderer.frag.spv (set 1 binding 0 = srcSampler, location 0 =
// fragTexCoord; DebugRenderer.h:44-50). The vendored C++ backend has no gl_VertexIndex builtin, so the 4-vertex stage
// (debugRenderer.vert:15-24) and the rasterisation of the axis-aligned quad are restated ("rule D", see oracle/ssvgi_oracle.c).
extern "C" int ref_debug_overlay(const lgcu_debug_quad_data *params, const lgcu_image *srcImg, const lgcu_image *targetImg, const lgcu_rows *rows) {
  int w, h, y0, y1;
  orc_level_size(targetImg, 0, &w, &h);
  ref_row_range(rows, 0, h, &y0, &y1);
  const glm::vec4 minmax(params->minmax[0], params->minmax[1], params->minmax[2], params->minmax[3]);
  const glm::vec2 p0 = (glm::vec2(minmax.x, minmax.y) + glm::vec2(0.0f, 0.0f) * (glm::vec2(minmax.z, minmax.w) - glm::vec2(minmax.x, minmax.y))) * 2.0f - glm::vec2(1.0f);
  const glm::vec2 p1 = (glm::vec2(minmax.x, minmax.y) + glm::vec2(1.0f, 1.0f) * (glm::vec2(minmax.z, minmax.w) - glm::vec2(minmax.x, minmax.y))) * 2.0f - glm::vec2(1.0f);
  const float wx0 = (p0.x + 1.0f) * (float(w) / 2.0f), wx1 = (p1.x + 1.0f) * (float(w) / 2.0f);
  const float wy0 = (p0.y + 1.0f) * (float(h) / 2.0f), wy1 = (p1.y + 1.0f) * (float(h) / 2.0f);
  if (!(wx1 > wx0) || !(wy1 > wy0)) return LGCU_OK;
#pragma omp parallel
  {
    RefShaderInstance s;
    RefSampler2D sSrc(srcImg);
    glm::vec2 texCoord;
    glm::vec4 out;
    s.resource(1, 0, &sSrc);
    s.input(0, &texCoord, sizeof(texCoord));
    s.output(0, &out, sizeof(out));
#pragma omp for schedule(dynamic, 8)
    for (int y = y0; y < y1; y++) {
      const float cy = float(y) + 0.5f;
      if (!(wy0 <= cy && cy < wy1)) continue;
      for (int x = 0; x < w; x++) {
        const float cx = float(x) + 0.5f;
        if (!(wx0 <= cx && cx < wx1)) continue;
        texCoord = glm::vec2((cx - wx0) / (wx1 - wx0), (cy - wy0) / (wy1 - wy0));
        s.setPixel(x, y, w, h);
        s.invoke();
        ref_store(targetImg, x, y, out);
      }
    }
  }
  return LGCU_OK;
}
#endif

#if REF_PASS == 1
extern "C" int ref_num_threads(void) { return omp_get_max_threads(); }
extern "C" void ref_set_num_threads(int n) { omp_set_num_threads(n); }
#endif
