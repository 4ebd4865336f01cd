// lgcu_shading.cuh — per-pixel device functions of the G-buffer resolve (K1) and direct lighting (K2), shared by the
// pass-granular kernels (k_streaming.cu) and the fused frame-front kernel (k_front.cu). Include only from translation units
// compiled with -fmad=false: every expression is evaluated in the reference shader's order.
#pragma once

#include "lgcu_kernels.h"

namespace lgcu {
namespace shading {

constexpr uint32_t F16 = LGCU_FORMAT_R16G16B16A16_SFLOAT, RG32 = LGCU_FORMAT_R32G32_SFLOAT, D32 = LGCU_FORMAT_D32_SFLOAT;

// pow(c, 2.2f) for the per-draw-call colours (gBufferBuilder.frag:34,36). Evaluated in double and rounded once:
// within a rounding of the correctly rounded float, which is what the CPU libm returns in all but ~1e-3 of cases.
__device__ __forceinline__ float pow22(float c) { return (float)pow((double)c, (double)2.2f); }

// ---------------------------------------------------------------------------------------------------- K1 (+K2)
constexpr int kMaxSharedObjects = 1024;

struct ObjectColors { // RGBA16F-packed pow(albedo, 2.2), pow(emissive, 2.2)
  uint2 albedo, emissive;
};

__device__ __forceinline__ ObjectColors objectColors(const lgcu_draw_call_data &o) {
  ObjectColors c;
  c.albedo = Texel<F16>::pack(make_float4(pow22(o.albedoColor[0]), pow22(o.albedoColor[1]), pow22(o.albedoColor[2]), pow22(o.albedoColor[3])));
  c.emissive = Texel<F16>::pack(make_float4(pow22(o.emissiveColor[0]), pow22(o.emissiveColor[1]), pow22(o.emissiveColor[2]), pow22(o.emissiveColor[3])));
  return c;
}

struct ResolvedTexel {
  uint2 albedo, emissive, normal; // RGBA16F bit patterns
  float2 moments;
  float depth;
};

// SH/Common/gBufferBuilder.frag:28-38 for one fragment (or the clear values for an uncovered pixel)
__device__ __forceinline__ ResolvedTexel resolveFragment(const GBufferArgs &a, const ObjectColors *table, int x, int y) {
  const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(a.fragments) + (size_t)y * a.fragmentPitch) + 2 * x;
  const float4 f0 = __ldg(src), f1 = __ldg(src + 1); // worldPos.xyz, normal.x | normal.yz, objectId, ndcDepth
  const uint32_t objectId = __float_as_uint(f1.z);
  ResolvedTexel r;
  if (objectId >= a.nObjects) { // LGCU_NO_OBJECT (or an out-of-range id): attachment clear values
    const uint2 c = Texel<F16>::pack(make_float4(a.clear.color[0], a.clear.color[1], a.clear.color[2], a.clear.color[3]));
    r.albedo = r.emissive = r.normal = c;
    r.moments = make_float2(a.clear.color[0], a.clear.color[1]);
    r.depth = a.clear.depth;
    return r;
  }
  ObjectColors oc;
  if (table)
    oc = table[objectId];
  else
    oc = objectColors(a.objects[objectId]);
  const V3 delta = v3(f0.x, f0.y, f0.z) - v3(a.cam[0], a.cam[1], a.cam[2]);
  const float len = sqrtf(dot3(delta, delta)); // :31
  r.albedo = oc.albedo;
  r.emissive = oc.emissive;
  r.normal = Texel<F16>::pack(make_float4(f0.w, f1.x, f1.y, 1.0f)); // :35
  r.moments = make_float2(len, len * len);                            // :37
  r.depth = f1.w;
  return r;
}

__device__ __forceinline__ const ObjectColors *stageObjectTable(const GBufferArgs &a, ObjectColors *smem) {
  if (a.nObjects > kMaxSharedObjects) return nullptr;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
  for (uint32_t i = tid; i < a.nObjects; i += nthreads) smem[i] = objectColors(a.objects[i]);
  __syncthreads();
  return smem;
}

__device__ __forceinline__ void storeResolved(const GBufferArgs &a, int x, int y, const ResolvedTexel &r) {
  reinterpret_cast<uint2 *>(a.albedo.ptr + (size_t)y * a.albedo.pitch)[x] = r.albedo;
  reinterpret_cast<uint2 *>(a.emissive.ptr + (size_t)y * a.emissive.pitch)[x] = r.emissive;
  reinterpret_cast<uint2 *>(a.normal.ptr + (size_t)y * a.normal.pitch)[x] = r.normal;
  reinterpret_cast<float2 *>(a.depthMoments.ptr + (size_t)y * a.depthMoments.pitch)[x] = r.moments;
  reinterpret_cast<float *>(a.depthStencil.ptr + (size_t)y * a.depthStencil.pitch)[x] = r.depth;
}

// SH/Common/directLighting.frag:45-83 given the four centre samples (fragScreenCoord at a pixel centre of a 1:1
// full-screen pass addresses exactly that texel: SURVEY.md Appendix B "centre-tap shortcut").
__device__ __forceinline__ float4 shadeDirect(const DirectLightArgs &a, int x, int y, float4 albedo, float4 emissive, float4 normal, float depth) {
  const float u = ((float)x + 0.5f) / (float)a.directLight.w, v = ((float)y + 0.5f) / (float)a.directLight.h;
  const V3 worldNormal = v3(normal.x, normal.y, normal.z);
  const V3 worldPos = unproject(u, v, depth, a.invViewProj);                                   // :56
  const V3 lightVec = worldPos - v3(a.lightPos[0], a.lightPos[1], a.lightPos[2]);              // :57
  const float diffuse = glmMax(0.0f, -dot3(normalize3(lightVec), worldNormal));                // :58
  const float4 ndc = mulMat4(a.lightViewProj, worldPos.x, worldPos.y, worldPos.z, 1.0f);       // Project() :36-43
  const float scx = (ndc.x / ndc.w) * 0.5f + 0.5f, scy = (ndc.y / ndc.w) * 0.5f + 0.5f, scz = ndc.z / ndc.w;
  const float4 lightViewPos = mulMat4(a.lightView, worldPos.x, worldPos.y, worldPos.z, 1.0f);  // :62
  const float dx = scx - 0.5f, dy = scy - 0.5f;
  const float radius = sqrtf(dx * dx + dy * dy) * 2.0f;                                        // :65
  const float t = glmMin(glmMax(1.0f - saturatef((radius - 0.6f) / 0.4f), 0.0f), 1.0f);        // smoothstep(0,1,.) :66
  const float penumbra = (t * t * (3.0f - 2.0f * t)) * (lightViewPos.z > 0.0f ? 1.0f : 0.0f);
  float intensity = 5.0f * penumbra;                                                           // :64, :67
  // texture(sampler2DShadow): 2x2 PCF, compare LESS_OR_EQUAL, clamp-to-edge (SSVGIRenderer.h:18)           :73
  const float ref = scz - 0.0002f;                                                             // :69-70
  const BilinearTaps tp = bilinearTaps(a.shadowMap, scx, scy);
  const float c00 = ref <= Texel<D32>::load(a.shadowMap, tp.x0, tp.y0).x ? 1.0f : 0.0f;
  const float c10 = ref <= Texel<D32>::load(a.shadowMap, tp.x1, tp.y0).x ? 1.0f : 0.0f;
  const float c01 = ref <= Texel<D32>::load(a.shadowMap, tp.x0, tp.y1).x ? 1.0f : 0.0f;
  const float c11 = ref <= Texel<D32>::load(a.shadowMap, tp.x1, tp.y1).x ? 1.0f : 0.0f;
  const float shadow = lerpExact(lerpExact(c00, c10, tp.a), lerpExact(c01, c11, tp.a), tp.b);
  float shadowPow = shadow; // pow(x, 2.2) is exact at 0 and 1, which is every pixel off a shadow edge
  if (shadow != 0.0f && shadow != 1.0f) shadowPow = powf(shadow, 2.2f);
  intensity = intensity * shadowPow;
  return make_float4((albedo.x * diffuse) * intensity + emissive.x, (albedo.y * diffuse) * intensity + emissive.y,
                     (albedo.z * diffuse) * intensity + emissive.z, 1.0f);                      // :79
}


} // namespace shading
} // namespace lgcu
