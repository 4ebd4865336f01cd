// lgcu_api.cu — the C ABI of liblgcu.so (include/lgcu.h): argument validation, view resolution, hoisting of the
// per-frame constants the reference shaders recompute per fragment, and kernel launches. Host code only; every
// entry point enqueues on the caller's stream and returns. There is no CPU implementation behind any of them.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "lgcu_kernels.h"
#include "lgcu_mat4.h"

using namespace lgcu;

namespace {

thread_local char g_lastError[512] = "";

int fail(int status, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
  va_end(ap);
  return status;
}

} // namespace

// shared with lgcu_interop.cu (same thread-local message buffer behind lgcu_last_error); not part of the public ABI
extern "C" __attribute__((visibility("hidden"))) int lgcu_set_last_error(int status, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
  va_end(ap);
  return status;
}

namespace {

int cudaStatus(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return LGCU_OK;
  return fail(LGCU_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

uint32_t texelSize(uint32_t format) {
  switch (format) {
    case LGCU_FORMAT_B8G8R8A8_SRGB: return 4;
    case LGCU_FORMAT_R16G16B16A16_SFLOAT: return 8;
    case LGCU_FORMAT_R32G32_SFLOAT: return 8;
    case LGCU_FORMAT_R32G32B32A32_SFLOAT: return 16;
    case LGCU_FORMAT_D32_SFLOAT: return 4;
    default: return 0;
  }
}

// VIEW level `lod` of `img` as a device view. Returns false (and sets the error) on a malformed descriptor.
bool resolveLevel(const lgcu_image *img, uint32_t lod, const char *name, LevelView *out, int *status) {
  if (!img || !img->base) {
    *status = fail(LGCU_ERR_INVALID_ARGUMENT, "%s: null image", name);
    return false;
  }
  const uint32_t level = img->baseMip + lod;
  if (img->imageMipCount == 0 || img->imageMipCount > LGCU_MAX_MIPS || lod >= img->mipCount || level >= img->imageMipCount) {
    *status = fail(LGCU_ERR_INVALID_ARGUMENT, "%s: level %u outside view [%u,+%u) of a %u-level image", name, lod, img->baseMip,
                   img->mipCount, img->imageMipCount);
    return false;
  }
  const uint32_t ts = texelSize(img->format);
  if (!ts) {
    *status = fail(LGCU_ERR_UNSUPPORTED_FORMAT, "%s: unsupported format %u", name, img->format);
    return false;
  }
  out->w = (int)(img->width >> level);
  out->h = (int)(img->height >> level);
  out->pitch = img->levelPitch[level];
  out->ptr = static_cast<unsigned char *>(img->base) + img->levelOffset[level];
  if (out->w > 0 && ((uint64_t)out->pitch < (uint64_t)out->w * ts || (out->pitch % ts) != 0 || (reinterpret_cast<uintptr_t>(out->ptr) % 16) != 0 ||
                     (out->pitch % 16) != 0)) {
    *status = fail(LGCU_ERR_INVALID_ARGUMENT, "%s: level %u pitch %u / alignment unusable for %d texels of %u bytes", name, level, out->pitch,
                   out->w, ts);
    return false;
  }
  return true;
}

bool expectFormat(const lgcu_image *img, uint32_t format, const char *name, int *status) {
  if (img && img->format == format) return true;
  *status = fail(LGCU_ERR_UNSUPPORTED_FORMAT, "%s: format %u, expected %u", name, img ? img->format : 0u, format);
  return false;
}

bool sameSize(const LevelView &a, const LevelView &b, const char *an, const char *bn, int *status) {
  if (a.w == b.w && a.h == b.h) return true;
  *status = fail(LGCU_ERR_INVALID_ARGUMENT, "%s is %dx%d but %s is %dx%d", an, a.w, a.h, bn, b.w, b.h);
  return false;
}

RowRange rowRange(const lgcu_rows *rows, uint32_t level, int h) {
  if (!rows) return RowRange{0, h};
  const uint32_t a = rows->y0 >> level, b = (rows->y1 + ((1u << level) - 1u)) >> level;
  return RowRange{(int)a < h ? (int)a : h, (int)b < h ? (int)b : h};
}

Mat4 toMat4(const lgcu_mat4 &m) {
  Mat4 r;
  std::memcpy(r.m, m.m, sizeof(r.m));
  return r;
}

void originOf(const lgcu_mat4 &inverseOfView, float out[3]) { // (inverse(view) * vec4(0,0,0,1)).xyz
  const float o[4] = {0.0f, 0.0f, 0.0f, 1.0f};
  float r[4];
  lgcu_mat4_mul_vec4(&inverseOfView, o, r);
  out[0] = r[0];
  out[1] = r[1];
  out[2] = r[2];
}

bool resolvePyramid(const lgcu_image *img, const char *name, PyramidView *out, int *status) {
  if (!img) {
    *status = fail(LGCU_ERR_INVALID_ARGUMENT, "%s: null image", name);
    return false;
  }
  if (img->mipCount == 0 || img->mipCount > (uint32_t)kMaxGatherLevels) {
    *status = fail(LGCU_ERR_INVALID_ARGUMENT, "%s: view has %u levels, supported 1..%d", name, img->mipCount, kMaxGatherLevels);
    return false;
  }
  out->count = (int)img->mipCount;
  for (int l = 0; l < out->count; l++)
    if (!resolveLevel(img, (uint32_t)l, name, &out->lv[l], status)) return false;
  // levels whose size degenerated to 0 (tiny images) are never sampled meaningfully; keep them addressable
  for (int l = 0; l < out->count; l++) {
    if (out->lv[l].w < 1) out->lv[l].w = 1;
    if (out->lv[l].h < 1) out->lv[l].h = 1;
  }
  return true;
}

// ---- GI gather tables (SH/SSVGI/indirectLighting.frag:155-178, :212, :217, :234-235) --------------------------
float gatherStepQuotient(float path, float nearStep) { return std::log(path / nearStep) / 0.944197714328765869140625f; }

bool buildGatherTables(float viewportX, float viewportY, int levels, GatherTables *t, int *status) {
  const float nearStep = viewportX / 1000.0f; // :167
  if (!(viewportX >= 1.0f) || !(viewportY >= 1.0f)) {
    *status = fail(LGCU_ERR_INVALID_ARGUMENT, "gi_gather: viewportExtent %g x %g", viewportX, viewportY);
    return false;
  }
  const float q = gatherStepQuotient(viewportX + viewportY, nearStep);
  const int maxSteps = (int)q + 1;
  if (maxSteps > kGatherMaxSteps) {
    *status = fail(LGCU_ERR_UNSUPPORTED, "gi_gather: %d march steps exceed the table capacity %d", maxSteps, kGatherMaxSteps);
    return false;
  }
  t->maxSteps = maxSteps < 1 ? 1 : maxSteps;
  for (int idx = 0; idx < 16; idx++) {
    uint32_t b = ((uint32_t)idx << 16) | ((uint32_t)idx >> 16); // HammersleyNorm :102-112
    b = ((b & 0x55555555u) << 1) | ((b & 0xAAAAAAAAu) >> 1);
    b = ((b & 0x33333333u) << 2) | ((b & 0xCCCCCCCCu) >> 2);
    b = ((b & 0x0F0F0F0Fu) << 4) | ((b & 0xF0F0F0F0u) >> 4);
    b = ((b & 0x00FF00FFu) << 8) | ((b & 0xFF00FF00u) >> 8);
    const float angOffset = (float)idx / 16.0f, linOffset = (float)b / 4294967296.0f;
    const float pixelAngOffset = 1.57075f * angOffset; // :170 (pi = 3.1415f)
    for (int d = 0; d < kGatherDirs; d++) {
      const float screenAng = pixelAngOffset + (1.57075f * (float)d); // :177
      t->dirX[idx][d] = std::cos(screenAng);
      t->dirY[idx][d] = std::sin(screenAng);
    }
    for (int k = 0; k < kGatherMaxSteps; k++) {
      const float off = ((nearStep * std::pow(2.57075f, (float)k + linOffset)) + 1.0f) - nearStep; // :217
      const float arg = (1.57075f * (off - 1.0f)) * 0.5f;
      float lod = (std::log(0.0f < arg ? arg : 0.0f) / 0.693147182464599609375f) + -2.0f;       // :234-235
      const float last = (float)(levels - 1);
      if (!(lod > 0.0f)) lod = 0.0f; // sampler minLod 0 (also -inf)
      if (lod > last) lod = last;    // view level count
      t->pixelOffset[idx][k] = off;
      t->lod[idx][k] = lod;
    }
  }
  // iterationsCount(path) = int(q(path)) + 1 > n  <=>  n == 0 ? q > -1 : q >= n ; q is monotone in path, so find the
  // smallest float that satisfies it by bisection over the (ordered) positive float bit patterns.
  for (int n = 0; n < kGatherMaxSteps; n++) {
    uint32_t lo = 0x00000001u, hi = 0x7F7FFFFFu; // smallest subnormal .. FLT_MAX
    auto reached = [&](uint32_t bits) {
      float path;
      std::memcpy(&path, &bits, 4);
      const float qq = gatherStepQuotient(path, nearStep);
      return n == 0 ? (qq > -1.0f) : (qq >= (float)n);
    };
    if (!reached(hi)) {
      lo = 0x7F800000u; // +inf: never
    } else {
      while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (reached(mid))
          hi = mid;
        else
          lo = mid + 1;
      }
    }
    std::memcpy(&t->iterThreshold[n], &lo, 4);
  }
  return true;
}

} // namespace

extern "C" {

int lgcu_abi_version(void) { return LGCU_ABI_VERSION; }
const char *lgcu_last_error(void) { return g_lastError; }
uint32_t lgcu_format_texel_size(uint32_t format) { return texelSize(format); }

uint64_t lgcu_image_layout(lgcu_image *img, uint32_t format, uint32_t width, uint32_t height, uint32_t mips) {
  const uint32_t ts = texelSize(format);
  if (!img || !ts || mips == 0 || mips > LGCU_MAX_MIPS) return 0;
  img->format = format;
  img->width = width;
  img->height = height;
  img->imageMipCount = mips;
  img->baseMip = 0;
  img->mipCount = mips;
  img->reserved0 = img->reserved1 = 0;
  uint64_t offset = 0;
  for (uint32_t l = 0; l < LGCU_MAX_MIPS; l++) {
    img->levelOffset[l] = 0;
    img->levelPitch[l] = 0;
  }
  for (uint32_t l = 0; l < mips; l++) {
    const uint64_t w = (width >> l) ? (width >> l) : 1, h = (height >> l) ? (height >> l) : 1;
    const uint64_t pitch = (w * ts + 127u) / 128u * 128u;
    img->levelOffset[l] = offset;
    img->levelPitch[l] = (uint32_t)pitch;
    offset += (pitch * h + 255u) / 256u * 256u;
  }
  return offset;
}

// ------------------------------------------------------------------------------------------------------- K1
static bool fillGBufferArgs(const lgcu_gbuffer_builder_data *params, const lgcu_draw_call_data *objects, uint32_t nObjects,
                            const lgcu_fragment *fragments, uint64_t fragmentPitchBytes, const lgcu_clear_values *clear,
                            const lgcu_image *albedo, const lgcu_image *emissive, const lgcu_image *normal, const lgcu_image *depthMoments,
                            const lgcu_image *depthStencil, const lgcu_rows *rows, GBufferArgs *a, int *st) {
  if (!params || !fragments || !clear || (!objects && nObjects)) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "gbuffer_resolve: null argument");
    return false;
  }
  if (!expectFormat(albedo, LGCU_FORMAT_R16G16B16A16_SFLOAT, "albedo", st) || !expectFormat(emissive, LGCU_FORMAT_R16G16B16A16_SFLOAT, "emissive", st) ||
      !expectFormat(normal, LGCU_FORMAT_R16G16B16A16_SFLOAT, "normal", st) || !expectFormat(depthMoments, LGCU_FORMAT_R32G32_SFLOAT, "depthMoments", st) ||
      !expectFormat(depthStencil, LGCU_FORMAT_D32_SFLOAT, "depthStencil", st))
    return false;
  if (!resolveLevel(albedo, 0, "albedo", &a->albedo, st) || !resolveLevel(emissive, 0, "emissive", &a->emissive, st) ||
      !resolveLevel(normal, 0, "normal", &a->normal, st) || !resolveLevel(depthMoments, 0, "depthMoments", &a->depthMoments, st) ||
      !resolveLevel(depthStencil, 0, "depthStencil", &a->depthStencil, st))
    return false;
  if (!sameSize(a->albedo, a->emissive, "albedo", "emissive", st) || !sameSize(a->albedo, a->normal, "albedo", "normal", st) ||
      !sameSize(a->albedo, a->depthMoments, "albedo", "depthMoments", st) || !sameSize(a->albedo, a->depthStencil, "albedo", "depthStencil", st))
    return false;
  if (fragmentPitchBytes < (uint64_t)a->albedo.w * sizeof(lgcu_fragment) || (fragmentPitchBytes % 16) != 0 ||
      (reinterpret_cast<uintptr_t>(fragments) % 16) != 0) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "gbuffer_resolve: fragment pitch %llu / alignment unusable for width %d",
               (unsigned long long)fragmentPitchBytes, a->albedo.w);
    return false;
  }
  a->fragments = fragments;
  a->fragmentPitch = fragmentPitchBytes;
  a->objects = objects;
  a->nObjects = nObjects;
  const lgcu_mat4 invView = lgcu_mat4_inverse(&params->viewMatrix); // gBufferBuilder.frag:30
  originOf(invView, a->cam);
  std::memcpy(a->clear.color, clear->color, sizeof(a->clear.color));
  a->clear.depth = clear->depth;
  a->rows = rowRange(rows, 0, a->albedo.h);
  return true;
}

int lgcu_gbuffer_resolve(const lgcu_gbuffer_builder_data *params, const lgcu_draw_call_data *objects, uint32_t nObjects,
                         const lgcu_fragment *fragments, uint64_t fragmentPitchBytes, const lgcu_clear_values *clear,
                         const lgcu_image *albedo, const lgcu_image *emissive, const lgcu_image *normal, const lgcu_image *depthMoments,
                         const lgcu_image *depthStencil, const lgcu_rows *rows, void *stream) {
  GBufferArgs a;
  int st = LGCU_OK;
  if (!fillGBufferArgs(params, objects, nObjects, fragments, fragmentPitchBytes, clear, albedo, emissive, normal, depthMoments, depthStencil, rows, &a, &st))
    return st;
  return cudaStatus(launchGBufferResolve(a, static_cast<cudaStream_t>(stream)), "gbuffer_resolve");
}

// ------------------------------------------------------------------------------------------------------- K2
static bool fillDirectLightArgs(const lgcu_direct_lighting_data *params, const lgcu_image *albedo, const lgcu_image *emissive,
                                const lgcu_image *normal, const lgcu_image *depthStencil, const lgcu_image *shadowMap,
                                const lgcu_image *directLight, const lgcu_rows *rows, DirectLightArgs *a, int *st) {
  if (!params) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "direct_light: null params");
    return false;
  }
  if (!expectFormat(albedo, LGCU_FORMAT_R16G16B16A16_SFLOAT, "albedo", st) || !expectFormat(emissive, LGCU_FORMAT_R16G16B16A16_SFLOAT, "emissive", st) ||
      !expectFormat(normal, LGCU_FORMAT_R16G16B16A16_SFLOAT, "normal", st) || !expectFormat(depthStencil, LGCU_FORMAT_D32_SFLOAT, "depthStencil", st) ||
      !expectFormat(shadowMap, LGCU_FORMAT_D32_SFLOAT, "shadowMap", st) || !expectFormat(directLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, "directLight", st))
    return false;
  if (!resolveLevel(albedo, 0, "albedo", &a->albedo, st) || !resolveLevel(emissive, 0, "emissive", &a->emissive, st) ||
      !resolveLevel(normal, 0, "normal", &a->normal, st) || !resolveLevel(depthStencil, 0, "depthStencil", &a->depthStencil, st) ||
      !resolveLevel(shadowMap, 0, "shadowMap", &a->shadowMap, st) || !resolveLevel(directLight, 0, "directLight", &a->directLight, st))
    return false;
  if (!sameSize(a->directLight, a->albedo, "directLight", "albedo", st) || !sameSize(a->directLight, a->emissive, "directLight", "emissive", st) ||
      !sameSize(a->directLight, a->normal, "directLight", "normal", st) || !sameSize(a->directLight, a->depthStencil, "directLight", "depthStencil", st))
    return false;
  const lgcu_mat4 viewProj = lgcu_mat4_mul(&params->projMatrix, &params->viewMatrix);                // :50
  a->invViewProj = toMat4(lgcu_mat4_inverse(&viewProj));                                            // :52
  const lgcu_mat4 invLightView = lgcu_mat4_inverse(&params->lightViewMatrix);                       // :54
  originOf(invLightView, a->lightPos);
  a->lightViewProj = toMat4(lgcu_mat4_mul(&params->lightProjMatrix, &params->lightViewMatrix));     // :60
  a->lightView = toMat4(params->lightViewMatrix);
  a->rows = rowRange(rows, 0, a->directLight.h);
  return true;
}

int lgcu_direct_light(const lgcu_direct_lighting_data *params, const lgcu_image *albedo, const lgcu_image *emissive, const lgcu_image *normal,
                      const lgcu_image *depthStencil, const lgcu_image *shadowMap, const lgcu_image *directLight, const lgcu_rows *rows,
                      void *stream) {
  DirectLightArgs a;
  int st = LGCU_OK;
  if (!fillDirectLightArgs(params, albedo, emissive, normal, depthStencil, shadowMap, directLight, rows, &a, &st)) return st;
  return cudaStatus(launchDirectLight(a, static_cast<cudaStream_t>(stream)), "direct_light");
}

int lgcu_gbuffer_direct_light(const lgcu_gbuffer_builder_data *gparams, const lgcu_direct_lighting_data *lparams,
                              const lgcu_draw_call_data *objects, uint32_t nObjects, const lgcu_fragment *fragments, uint64_t fragmentPitchBytes,
                              const lgcu_clear_values *clear, const lgcu_image *albedo, const lgcu_image *emissive, const lgcu_image *normal,
                              const lgcu_image *depthMoments, const lgcu_image *depthStencil, const lgcu_image *shadowMap,
                              const lgcu_image *directLight, const lgcu_rows *rows, void *stream) {
  GBufferLightArgs a;
  int st = LGCU_OK;
  if (!fillGBufferArgs(gparams, objects, nObjects, fragments, fragmentPitchBytes, clear, albedo, emissive, normal, depthMoments, depthStencil, rows, &a.g, &st))
    return st;
  if (!fillDirectLightArgs(lparams, albedo, emissive, normal, depthStencil, shadowMap, directLight, rows, &a.l, &st)) return st;
  return cudaStatus(launchGBufferDirectLight(a, static_cast<cudaStream_t>(stream)), "gbuffer_direct_light");
}

// ------------------------------------------------------------------------------------------------------- K3 / K4
static bool chainFormat(uint32_t f) { return f == LGCU_FORMAT_R16G16B16A16_SFLOAT || f == LGCU_FORMAT_R32G32_SFLOAT; }

int lgcu_mip_level(const lgcu_mip_level_builder_data *params, const lgcu_image *srcLevel, const lgcu_image *dstLevel, const lgcu_rows *rows,
                   void *stream) {
  int st = LGCU_OK;
  if (!params || !srcLevel || !dstLevel) return fail(LGCU_ERR_INVALID_ARGUMENT, "mip_level: null argument");
  if (!chainFormat(srcLevel->format) || srcLevel->format != dstLevel->format)
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "mip_level: formats %u -> %u", srcLevel->format, dstLevel->format);
  MipLevelArgs a;
  a.format = srcLevel->format;
  a.depthFilter = params->filterType < 0.5f ? 0 : 1; // mipLevelBuilder.frag:22
  if (!resolveLevel(srcLevel, 0, "mip src", &a.src, &st) || !resolveLevel(dstLevel, 0, "mip dst", &a.dst, &st)) return st;
  if (a.dst.w * 2 > a.src.w || a.dst.h * 2 > a.src.h)
    return fail(LGCU_ERR_INVALID_ARGUMENT, "mip_level: dst %dx%d is not a half-size level of src %dx%d", a.dst.w, a.dst.h, a.src.w, a.src.h);
  a.rows = rowRange(rows, dstLevel->baseMip, a.dst.h);
  return cudaStatus(launchMipLevel(a, static_cast<cudaStream_t>(stream)), "mip_level");
}

int lgcu_blur_level(const lgcu_blur_layer_builder_data *params, const lgcu_image *srcLevel, const lgcu_image *dstLevel, const lgcu_rows *rows,
                    void *stream) {
  int st = LGCU_OK;
  if (!params || !srcLevel || !dstLevel) return fail(LGCU_ERR_INVALID_ARGUMENT, "blur_level: null argument");
  if (!chainFormat(srcLevel->format) || srcLevel->format != dstLevel->format)
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "blur_level: formats %u -> %u", srcLevel->format, dstLevel->format);
  BlurLevelArgs a;
  a.format = srcLevel->format;
  if (!resolveLevel(srcLevel, 0, "blur src", &a.src, &st) || !resolveLevel(dstLevel, 0, "blur dst", &a.dst, &st)) return st;
  if (!sameSize(a.src, a.dst, "blur src", "blur dst", &st)) return st;
  a.sizeX = params->size[0];
  a.sizeY = params->size[1];
  a.radius = params->radius;
  if (a.radius < 0 || a.radius > 64 || a.sizeX < 1 || a.sizeY < 1 || a.sizeX > a.src.w || a.sizeY > a.src.h)
    return fail(LGCU_ERR_INVALID_ARGUMENT, "blur_level: radius %d / size %dx%d on a %dx%d level", a.radius, a.sizeX, a.sizeY, a.src.w, a.src.h);
  a.rows = rowRange(rows, dstLevel->baseMip, a.dst.h);
  return cudaStatus(launchBlurLevel(a, static_cast<cudaStream_t>(stream)), "blur_level");
}

// ---- interleaved rendering and the debug overlay ---------------------------------------------------
static int fillInterleaveArgs(const char *what, const lgcu_interleave_data *params, const lgcu_image *interleaved, const lgcu_image *deinterleaved,
                              const lgcu_image *dst, const lgcu_rows *rows, InterleaveArgs *a) {
  int st = LGCU_OK;
  if (!params || !interleaved || !deinterleaved) return fail(LGCU_ERR_INVALID_ARGUMENT, "%s: null argument", what);
  const uint32_t f = interleaved->format;
  if (f != deinterleaved->format || (f != LGCU_FORMAT_R16G16B16A16_SFLOAT && f != LGCU_FORMAT_R32G32_SFLOAT && f != LGCU_FORMAT_R32G32B32A32_SFLOAT))
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "%s: formats %u / %u (one of RGBA16F, RG32F, RGBA32F on both sides)", what, interleaved->format, deinterleaved->format);
  if (!resolveLevel(interleaved, 0, "interleaved", &a->interleaved, &st) || !resolveLevel(deinterleaved, 0, "deinterleaved", &a->deinterleaved, &st)) return st;
  if (!sameSize(a->interleaved, a->deinterleaved, "interleaved", "deinterleaved", &st)) return st;
  const int w = a->interleaved.w, h = a->interleaved.h;
  if (params->viewportSize[0] != w || params->viewportSize[1] != h || params->gridSize[0] < 1 || params->gridSize[1] < 1 || params->gridSize[0] > w ||
      params->gridSize[1] > h)
    return fail(LGCU_ERR_INVALID_ARGUMENT, "%s: viewportSize %dx%d / gridSize %dx%d on %dx%d images", what, params->viewportSize[0], params->viewportSize[1],
                params->gridSize[0], params->gridSize[1], w, h);
  a->texelBytes = (int)texelSize(f);
  a->gridX = params->gridSize[0];
  a->gridY = params->gridSize[1];
  a->cellsX = w / a->gridX; // deinterleave.frag:19, interleave.frag:18
  a->cellsY = h / a->gridY;
  a->rows = rowRange(rows, dst->baseMip, h);
  return LGCU_OK;
}

int lgcu_deinterleave(const lgcu_interleave_data *params, const lgcu_image *interleaved, const lgcu_image *deinterleaved, const lgcu_rows *rows, void *stream) {
  InterleaveArgs a;
  const int st = fillInterleaveArgs("deinterleave", params, interleaved, deinterleaved, deinterleaved, rows, &a);
  if (st != LGCU_OK) return st;
  return cudaStatus(launchDeinterleave(a, static_cast<cudaStream_t>(stream)), "deinterleave");
}

int lgcu_interleave(const lgcu_interleave_data *params, const lgcu_image *deinterleaved, const lgcu_image *interleaved, const lgcu_rows *rows, void *stream) {
  InterleaveArgs a;
  const int st = fillInterleaveArgs("interleave", params, interleaved, deinterleaved, interleaved, rows, &a);
  if (st != LGCU_OK) return st;
  return cudaStatus(launchInterleave(a, static_cast<cudaStream_t>(stream)), "interleave");
}

int lgcu_debug_overlay(const lgcu_debug_quad_data *params, const lgcu_image *src, const lgcu_image *target, const lgcu_rows *rows, void *stream) {
  int st = LGCU_OK;
  if (!params || !src || !target) return fail(LGCU_ERR_INVALID_ARGUMENT, "debug_overlay: null argument");
  if (src->format != LGCU_FORMAT_R16G16B16A16_SFLOAT && src->format != LGCU_FORMAT_R32G32_SFLOAT && src->format != LGCU_FORMAT_R32G32B32A32_SFLOAT)
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "debug_overlay: source format %u", src->format);
  if (target->format != LGCU_FORMAT_B8G8R8A8_SRGB && target->format != LGCU_FORMAT_R16G16B16A16_SFLOAT && target->format != LGCU_FORMAT_R32G32B32A32_SFLOAT)
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "debug_overlay: target format %u", target->format);
  if (src->mipCount != 1) return fail(LGCU_ERR_UNSUPPORTED, "debug_overlay: %u-level source view (the live path shows single-level views, SSVGIRenderer.h:344-350)", src->mipCount);
  DebugOverlayArgs a;
  a.srcFormat = src->format;
  a.targetFormat = target->format;
  if (!resolveLevel(src, 0, "debug src", &a.src, &st) || !resolveLevel(target, 0, "debug target", &a.target, &st)) return st;
  const float *mm = params->minmax;
  const float w = (float)a.target.w, h = (float)a.target.h;
  // debugRenderer.vert:20 for corners (0,0) and (1,1), then the viewport transform (this unit is compiled with -fmad=false)
  const float ndcX0 = (mm[0] + 0.0f * (mm[2] - mm[0])) * 2.0f - 1.0f, ndcX1 = (mm[0] + 1.0f * (mm[2] - mm[0])) * 2.0f - 1.0f;
  const float ndcY0 = (mm[1] + 0.0f * (mm[3] - mm[1])) * 2.0f - 1.0f, ndcY1 = (mm[1] + 1.0f * (mm[3] - mm[1])) * 2.0f - 1.0f;
  a.wx0 = (ndcX0 + 1.0f) * (w / 2.0f);
  a.wx1 = (ndcX1 + 1.0f) * (w / 2.0f);
  a.wy0 = (ndcY0 + 1.0f) * (h / 2.0f);
  a.wy1 = (ndcY1 + 1.0f) * (h / 2.0f);
  if (!(a.wx1 > a.wx0) || !(a.wy1 > a.wy0)) return LGCU_OK; // degenerate or non-finite quad covers nothing
  const RowRange rr = rowRange(rows, 0, a.target.h);
  auto lo = [](float v, int limit) { return v <= 0.0f ? 0 : (v >= (float)limit ? limit : (int)v); };
  a.x0 = lo(a.wx0 - 1.0f, a.target.w);
  a.x1 = lo(a.wx1 + 1.0f, a.target.w);
  a.y0 = lo(a.wy0 - 1.0f, a.target.h);
  a.y1 = lo(a.wy1 + 1.0f, a.target.h);
  if (a.y0 < rr.y0) a.y0 = rr.y0;
  if (a.y1 > rr.y1) a.y1 = rr.y1;
  return cudaStatus(launchDebugOverlay(a, static_cast<cudaStream_t>(stream)), "debug_overlay");
}

int lgcu_mip_blur_chain(const lgcu_image *chain, const lgcu_image *blurred, int32_t radius, const lgcu_rows *rows, void *stream) {
  int st = LGCU_OK;
  if (!chain || !blurred) return fail(LGCU_ERR_INVALID_ARGUMENT, "mip_blur_chain: null argument");
  if (!chainFormat(chain->format) || chain->format != blurred->format)
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "mip_blur_chain: formats %u / %u", chain->format, blurred->format);
  if (chain->mipCount != blurred->mipCount || chain->width != blurred->width || chain->height != blurred->height || chain->baseMip != 0 ||
      blurred->baseMip != 0)
    return fail(LGCU_ERR_INVALID_ARGUMENT, "mip_blur_chain: chain and blurred must be whole-image views of identical size");
  if (radius < 1 || radius > 2) return fail(LGCU_ERR_UNSUPPORTED, "mip_blur_chain: radius %d (the reference uses 2)", radius);
  ChainArgs a;
  a.format = chain->format;
  a.radius = radius;
  if (!resolvePyramid(chain, "chain", &a.chain, &st) || !resolvePyramid(blurred, "blurred", &a.blurred, &st)) return st;
  // MipBuilder::BuildMips stops at the first level with a zero dimension (MipBuilder.h:151-152)
  int levels = 1;
  for (int l = 1; l < a.chain.count; l++) {
    if ((chain->width >> l) == 0 || (chain->height >> l) == 0) break;
    levels++;
  }
  a.levels = levels;
  a.rows = rowRange(rows, 0, (int)chain->height);
  return cudaStatus(launchMipBlurChain(a, static_cast<cudaStream_t>(stream)), "mip_blur_chain");
}

// ------------------------------------------------------------------------------------------------------- frame front / chains
static int smCountOfCurrentDevice() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  return sms;
}

static int builtLevels(const lgcu_image *img) { // MipBuilder::BuildMips stops at the first level with a zero dimension (MipBuilder.h:151-152)
  int levels = 1;
  for (uint32_t l = 1; l < img->mipCount; l++) {
    if ((img->width >> l) == 0 || (img->height >> l) == 0) break;
    levels++;
  }
  return levels;
}

static bool wholeChain(const lgcu_image *img, uint32_t format, const lgcu_image *like, const char *name, int *st) {
  if (!expectFormat(img, format, name, st)) return false;
  if (img->baseMip != 0 || img->mipCount == 0 || img->mipCount > (uint32_t)kMaxGatherLevels || (like && (img->width != like->width || img->height != like->height || img->mipCount != like->mipCount))) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "%s: must be a whole-image view of the chain (same size and level count as its partner)", name);
    return false;
  }
  return true;
}

int lgcu_frame_front(const lgcu_gbuffer_builder_data *gparams, const lgcu_direct_lighting_data *lparams, const lgcu_draw_call_data *objects,
                     uint32_t nObjects, const lgcu_fragment *fragments, uint64_t fragmentPitchBytes, const lgcu_clear_values *clear,
                     const lgcu_image *albedo, const lgcu_image *emissive, const lgcu_image *normal, const lgcu_image *depthMoments,
                     const lgcu_image *depthStencil, const lgcu_image *shadowMap, const lgcu_image *directLight, const lgcu_image *blurredDirectLight,
                     const lgcu_image *blurredDepthMoments, const lgcu_rows *rows, void *stream) {
  FrontArgs a;
  int st = LGCU_OK;
  if (!wholeChain(directLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, nullptr, "directLight", &st) ||
      !wholeChain(blurredDirectLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, directLight, "blurredDirectLight", &st) ||
      !wholeChain(depthMoments, LGCU_FORMAT_R32G32_SFLOAT, directLight, "depthMoments", &st) ||
      !wholeChain(blurredDepthMoments, LGCU_FORMAT_R32G32_SFLOAT, directLight, "blurredDepthMoments", &st))
    return st;
  if (!fillGBufferArgs(gparams, objects, nObjects, fragments, fragmentPitchBytes, clear, albedo, emissive, normal, depthMoments, depthStencil, rows, &a.g, &st))
    return st;
  if (!fillDirectLightArgs(lparams, albedo, emissive, normal, depthStencil, shadowMap, directLight, rows, &a.l, &st)) return st;
  if (!resolveLevel(blurredDirectLight, 0, "blurredDirectLight", &a.blurLight0, &st) || !resolveLevel(blurredDepthMoments, 0, "blurredDepthMoments", &a.blurMoments0, &st))
    return st;
  if (!sameSize(a.g.albedo, a.blurLight0, "albedo", "blurredDirectLight", &st)) return st;
  if ((a.g.rows.y0 % 16) != 0 || ((a.g.rows.y1 % 16) != 0 && a.g.rows.y1 != a.g.albedo.h))
    return fail(LGCU_ERR_INVALID_ARGUMENT, "frame_front: row strip [%d,%d) must start and end on multiples of 16 rows", a.g.rows.y0, a.g.rows.y1);
  const int levels = builtLevels(directLight);
  a.mipLevels = levels - 1 < kFrontMipLevels ? levels - 1 : kFrontMipLevels;
  for (int l = 1; l <= kFrontMipLevels; l++) {
    const uint32_t lod = l <= a.mipLevels ? (uint32_t)l : 0u;
    if (!resolveLevel(directLight, lod, "directLight", &a.lightMip[l - 1], &st) || !resolveLevel(depthMoments, lod, "depthMoments", &a.momentsMip[l - 1], &st)) return st;
  }
  return cudaStatus(launchFrameFront(a, smCountOfCurrentDevice(), static_cast<cudaStream_t>(stream)), "frame_front");
}

int lgcu_frame_chains(const lgcu_image *directLight, const lgcu_image *blurredDirectLight, const lgcu_image *depthMoments,
                      const lgcu_image *blurredDepthMoments, int32_t radius, const lgcu_rows *rows, void *stream) {
  ChainsArgs a;
  int st = LGCU_OK;
  if (!wholeChain(directLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, nullptr, "directLight", &st) ||
      !wholeChain(blurredDirectLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, directLight, "blurredDirectLight", &st) ||
      !wholeChain(depthMoments, LGCU_FORMAT_R32G32_SFLOAT, directLight, "depthMoments", &st) ||
      !wholeChain(blurredDepthMoments, LGCU_FORMAT_R32G32_SFLOAT, directLight, "blurredDepthMoments", &st))
    return st;
  if (radius < 1 || radius > 2) return fail(LGCU_ERR_UNSUPPORTED, "frame_chains: radius %d (the reference uses 2)", radius);
  if (!resolvePyramid(directLight, "directLight", &a.light, &st) || !resolvePyramid(blurredDirectLight, "blurredDirectLight", &a.blurredLight, &st) ||
      !resolvePyramid(depthMoments, "depthMoments", &a.moments, &st) || !resolvePyramid(blurredDepthMoments, "blurredDepthMoments", &a.blurredMoments, &st))
    return st;
  for (int l = 0; l < a.light.count; l++)
    if (a.light.lv[l].pitch != a.blurredLight.lv[l].pitch || a.moments.lv[l].pitch != a.blurredMoments.lv[l].pitch || a.light.lv[l].w != a.moments.lv[l].w)
      return fail(LGCU_ERR_INVALID_ARGUMENT, "frame_chains: the four chains must share one layout");
  a.levels = builtLevels(directLight);
  a.radius = radius;
  a.gridLevels = a.levels - 1 < kFrontMipLevels ? a.levels - 1 : kFrontMipLevels;
  a.rows = rowRange(rows, 0, (int)directLight->height);
  return cudaStatus(launchFrameChains(a, static_cast<cudaStream_t>(stream)), "frame_chains");
}

// ------------------------------------------------------------------------------------------------------- rasterisation front end
uint32_t lgcu_raster_prepare_draws(lgcu_draw *hostDraws, uint32_t nDraws) {
  uint32_t tri = 0;
  for (uint32_t i = 0; hostDraws && i < nDraws; i++) {
    hostDraws[i].firstTriangle = tri;
    tri += hostDraws[i].indexCount / 3;
  }
  return tri;
}

uint64_t lgcu_raster_scratch_bytes(uint32_t nTriangles, uint32_t width, uint32_t height) { return rasterScratchBytes(nTriangles, width, height); }

static bool fillRasterArgs(const lgcu_mesh_scene *scene, const lgcu_mat4 &view, const lgcu_mat4 &proj, void *scratch, uint64_t scratchBytes, uint32_t width,
                           uint32_t height, RasterArgs *a, int *st) {
  if (!scene || !scratch) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "raster: null scene / scratch");
    return false;
  }
  if (scene->nTriangles && (!scene->vertices || !scene->indices || !scene->draws || !scene->objects || !scene->nDraws)) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "raster: scene with %u triangles but a null vertex / index / draw / object array", scene->nTriangles);
    return false;
  }
  if ((reinterpret_cast<uintptr_t>(scene->vertices) % 16) != 0 || (reinterpret_cast<uintptr_t>(scratch) % 256) != 0) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "raster: vertices must be 16-byte aligned and scratch 256-byte aligned");
    return false;
  }
  if (width == 0 || height == 0 || width > 32768 || height > 32768 || scratchBytes < rasterScratchBytes(scene->nTriangles, width, height)) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "raster: %ux%u target needs %llu scratch bytes, got %llu", width, height,
               (unsigned long long)rasterScratchBytes(scene->nTriangles, width, height), (unsigned long long)scratchBytes);
    return false;
  }
  a->scene = *scene;
  const lgcu_mat4 vp = lgcu_mat4_mul(&proj, &view); // gl_Position = projMatrix * viewMatrix * v is (proj * view) * v   gBufferBuilder.vert:36
  a->viewProj = toMat4(vp);
  a->width = (int)width;
  a->height = (int)height;
  a->scratch = scratch;
  a->fragments = nullptr;
  a->fragmentPitch = 0;
  return true;
}

int lgcu_raster_shadow_map(const lgcu_shadowmap_builder_data *params, const lgcu_mesh_scene *scene, void *scratch, uint64_t scratchBytes,
                           const lgcu_image *shadowMap, void *stream) {
  int st = LGCU_OK;
  RasterArgs a;
  if (!params) return fail(LGCU_ERR_INVALID_ARGUMENT, "raster_shadow_map: null params");
  if (!expectFormat(shadowMap, LGCU_FORMAT_D32_SFLOAT, "shadowMap", &st) || !resolveLevel(shadowMap, 0, "shadowMap", &a.depth, &st)) return st;
  if (!fillRasterArgs(scene, params->lightViewMatrix, params->lightProjMatrix, scratch, scratchBytes, (uint32_t)a.depth.w, (uint32_t)a.depth.h, &a, &st)) return st;
  a.rows = RowRange{0, a.height};
  return cudaStatus(launchRaster(a, smCountOfCurrentDevice(), static_cast<cudaStream_t>(stream)), "raster_shadow_map");
}

int lgcu_raster_gbuffer(const lgcu_gbuffer_builder_data *params, const lgcu_mesh_scene *scene, void *scratch, uint64_t scratchBytes, uint32_t width,
                        uint32_t height, lgcu_fragment *fragments, uint64_t fragmentPitchBytes, const lgcu_rows *rows, void *stream) {
  int st = LGCU_OK;
  RasterArgs a;
  if (!params || !fragments) return fail(LGCU_ERR_INVALID_ARGUMENT, "raster_gbuffer: null params / fragments");
  if (fragmentPitchBytes < (uint64_t)width * sizeof(lgcu_fragment) || (fragmentPitchBytes % 16) != 0 || (reinterpret_cast<uintptr_t>(fragments) % 16) != 0)
    return fail(LGCU_ERR_INVALID_ARGUMENT, "raster_gbuffer: fragment pitch %llu / alignment unusable for %u pixels", (unsigned long long)fragmentPitchBytes, width);
  if (!fillRasterArgs(scene, params->viewMatrix, params->projMatrix, scratch, scratchBytes, width, height, &a, &st)) return st;
  a.rows = rowRange(rows, 0, (int)height);
  a.fragments = fragments;
  a.fragmentPitch = fragmentPitchBytes;
  a.depth = LevelView{nullptr, 0, 0, 0};
  return cudaStatus(launchRaster(a, smCountOfCurrentDevice(), static_cast<cudaStream_t>(stream)), "raster_gbuffer");
}

// ------------------------------------------------------------------------------------------------------- peer-to-peer rows
int lgcu_copy_rows(const lgcu_row_copy *copies, uint32_t count, void *stream) {
  if (!copies && count) return fail(LGCU_ERR_INVALID_ARGUMENT, "copy_rows: null list");
  const int sms = smCountOfCurrentDevice();
  for (uint32_t next = 0; next < count;) { // chunks of kMaxCopies non-empty slabs; `next` = first entry not consumed yet
    RowCopyArgs a;
    a.count = 0;
    uint64_t units = 0;
    for (; next < count && a.count < kMaxCopies; next++) {
      const uint32_t i = next;
      const lgcu_row_copy &c = copies[i];
      if (!c.bytes) continue;
      if (!c.src || !c.dst || (c.bytes % 16) != 0 || (reinterpret_cast<uintptr_t>(c.src) % 16) != 0 || (reinterpret_cast<uintptr_t>(c.dst) % 16) != 0)
        return fail(LGCU_ERR_INVALID_ARGUMENT, "copy_rows: slab %u must be non-null, 16-byte aligned and a multiple of 16 bytes", i);
      units += c.bytes / 16;
      a.src[a.count] = c.src;
      a.dst[a.count] = c.dst;
      a.unitEnd[a.count] = units;
      a.count++;
    }
    if (a.count == 0) continue;
    const int st = cudaStatus(launchRowCopies(a, sms, static_cast<cudaStream_t>(stream)), "copy_rows");
    if (st != LGCU_OK) return st;
  }
  return LGCU_OK;
}

static bool fillFlags(uint32_t *const *flags, uint32_t count, FlagArgs *a) {
  if (count > (uint32_t)kMaxFlags || (!flags && count)) return false;
  a->count = (int)count;
  for (uint32_t i = 0; i < count; i++) {
    if (!flags[i]) return false;
    a->flags[i] = flags[i];
  }
  return true;
}

int lgcu_exchange(const lgcu_exchange_desc *d, void *stream) {
  if (!d || !d->frameCounter) return fail(LGCU_ERR_INVALID_ARGUMENT, "exchange: null descriptor / frame counter");
  ExchangeArgs a;
  if (!fillFlags(d->signalBefore, d->signalBeforeCount, &a.signalBefore) || !fillFlags(d->wait, d->waitCount, &a.wait) || !fillFlags(d->signalAfter, d->signalAfterCount, &a.signalAfter))
    return fail(LGCU_ERR_INVALID_ARGUMENT, "exchange: bad flag list (max %d)", kMaxFlags);
  if (d->copyCount > (uint32_t)kMaxCopies || (!d->copies && d->copyCount)) return fail(LGCU_ERR_INVALID_ARGUMENT, "exchange: at most %d slabs per step", kMaxCopies);
  if (d->bump && d->copyCount) return fail(LGCU_ERR_INVALID_ARGUMENT, "exchange: the step that bumps the frame counter cannot copy (single CTA)");
  if (d->signalAfterCount && !d->doneCounter) return fail(LGCU_ERR_INVALID_ARGUMENT, "exchange: signalAfter needs a doneCounter");
  a.copies.count = 0;
  uint64_t units = 0;
  for (uint32_t i = 0; i < d->copyCount; i++) {
    const lgcu_row_copy &c = d->copies[i];
    if (!c.bytes) continue;
    if (!c.src || !c.dst || (c.bytes % 16) != 0 || (reinterpret_cast<uintptr_t>(c.src) % 16) != 0 || (reinterpret_cast<uintptr_t>(c.dst) % 16) != 0)
      return fail(LGCU_ERR_INVALID_ARGUMENT, "exchange: slab %u must be non-null, 16-byte aligned and a multiple of 16 bytes", i);
    units += c.bytes / 16;
    a.copies.src[a.copies.count] = c.src;
    a.copies.dst[a.copies.count] = c.dst;
    a.copies.unitEnd[a.copies.count] = units;
    a.copies.count++;
  }
  a.frame = d->frameCounter;
  a.done = d->doneCounter;
  a.lag = (int)d->lag;
  a.bump = d->bump ? 1 : 0;
  return cudaStatus(launchExchange(a, smCountOfCurrentDevice(), static_cast<cudaStream_t>(stream)), "exchange");
}

int lgcu_frame_counter_bump(uint32_t *frameCounter, void *stream) {
  if (!frameCounter) return fail(LGCU_ERR_INVALID_ARGUMENT, "frame_counter_bump: null counter");
  return cudaStatus(launchBumpFrame(frameCounter, static_cast<cudaStream_t>(stream)), "frame_counter_bump");
}

int lgcu_signal_flags(uint32_t *const *flags, uint32_t count, const uint32_t *frameCounter, void *stream) {
  FlagArgs a;
  if (!frameCounter || !fillFlags(flags, count, &a)) return fail(LGCU_ERR_INVALID_ARGUMENT, "signal_flags: bad flag list (max %d)", kMaxFlags);
  return cudaStatus(launchSignal(a, frameCounter, static_cast<cudaStream_t>(stream)), "signal_flags");
}

int lgcu_wait_flags(uint32_t *const *flags, uint32_t count, const uint32_t *frameCounter, uint32_t lag, void *stream) {
  FlagArgs a;
  if (!frameCounter || !fillFlags(flags, count, &a)) return fail(LGCU_ERR_INVALID_ARGUMENT, "wait_flags: bad flag list (max %d)", kMaxFlags);
  return cudaStatus(launchWait(a, frameCounter, (int)lag, static_cast<cudaStream_t>(stream)), "wait_flags");
}

// ------------------------------------------------------------------------------------------------------- K5
static bool fillGatherArgs(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight, const lgcu_image *blurredDepthMoments,
                           const lgcu_image *normal, const lgcu_image *depthStencil, const lgcu_image *indirectLight, const lgcu_rows *rows, GatherArgs *a,
                           int *st) {
  if (!params) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "gi_gather: null params");
    return false;
  }
  if (!expectFormat(blurredDirectLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, "blurredDirectLight", st) ||
      !expectFormat(blurredDepthMoments, LGCU_FORMAT_R32G32_SFLOAT, "blurredDepthMoments", st) ||
      !expectFormat(normal, LGCU_FORMAT_R16G16B16A16_SFLOAT, "normal", st) || !expectFormat(depthStencil, LGCU_FORMAT_D32_SFLOAT, "depthStencil", st))
    return false;
  if (!indirectLight || (indirectLight->format != LGCU_FORMAT_R16G16B16A16_SFLOAT && indirectLight->format != LGCU_FORMAT_R32G32B32A32_SFLOAT)) {
    *st = fail(LGCU_ERR_UNSUPPORTED_FORMAT, "indirectLight: format %u", indirectLight ? indirectLight->format : 0u);
    return false;
  }
  a->outFormat = indirectLight->format;
  if (!resolvePyramid(blurredDirectLight, "blurredDirectLight", &a->light, st) || !resolvePyramid(blurredDepthMoments, "blurredDepthMoments", &a->moments, st))
    return false;
  if (a->light.count != a->moments.count) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "gi_gather: pyramids have %d and %d levels", a->light.count, a->moments.count);
    return false;
  }
  if (!resolveLevel(normal, 0, "normal", &a->normal, st) || !resolveLevel(depthStencil, 0, "depthStencil", &a->depthStencil, st) ||
      !resolveLevel(indirectLight, 0, "indirectLight", &a->indirect, st))
    return false;
  if (!sameSize(a->indirect, a->normal, "indirectLight", "normal", st) || !sameSize(a->indirect, a->depthStencil, "indirectLight", "depthStencil", st) ||
      !sameSize(a->light.lv[0], a->moments.lv[0], "blurredDirectLight", "blurredDepthMoments", st))
    return false;
  const lgcu_mat4 viewProj = lgcu_mat4_mul(&params->projMatrix, &params->viewMatrix); // :122
  a->invViewProj = toMat4(lgcu_mat4_inverse(&viewProj));                             // :123
  const lgcu_mat4 invView = lgcu_mat4_inverse(&params->viewMatrix);                  // :124
  originOf(invView, a->cam);                                                         // :127
  a->viewport[0] = params->viewportExtent[0];
  a->viewport[1] = params->viewportExtent[1];
  a->rows = rowRange(rows, 0, a->indirect.h);
  return true;
}

static bool checkScratch(const GatherArgs &a, const lgcu_image *moments, const void *scratch, uint64_t scratchBytes, int *st) {
  const uint64_t need = gatherScratchBytes(moments->width >> moments->baseMip, moments->height >> moments->baseMip, (uint32_t)a.moments.count);
  if (!scratch || (reinterpret_cast<uintptr_t>(scratch) % 16) != 0 || scratchBytes < need) {
    *st = fail(LGCU_ERR_INVALID_ARGUMENT, "gi_gather: scratch %p / %llu bytes, need %llu bytes, 16-byte aligned", scratch, (unsigned long long)scratchBytes,
               (unsigned long long)need);
    return false;
  }
  return true;
}

int lgcu_gi_gather(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight, const lgcu_image *blurredDepthMoments,
                   const lgcu_image *normal, const lgcu_image *depthStencil, const lgcu_image *indirectLight, uint32_t flags, const lgcu_rows *rows,
                   void *stream) {
  int st = LGCU_OK;
  GatherArgs a;
  if (!fillGatherArgs(params, blurredDirectLight, blurredDepthMoments, normal, depthStencil, indirectLight, rows, &a, &st)) return st;
  GatherTables tables;
  if (!buildGatherTables(a.viewport[0], a.viewport[1], a.light.count, &tables, &st)) return st;
  if (flags & LGCU_GI_STRICT) return cudaStatus(launchGatherStrict(a, tables, static_cast<cudaStream_t>(stream)), "gi_gather(strict)");
  return cudaStatus(launchGatherFast(a, tables, nullptr, static_cast<cudaStream_t>(stream)), "gi_gather");
}

uint64_t lgcu_gather_scratch_bytes(uint32_t width, uint32_t height, uint32_t mips) { return gatherScratchBytes(width, height, mips); }

int lgcu_gi_gather_pack(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight, const lgcu_image *blurredDepthMoments,
                        const lgcu_image *normal, const lgcu_image *depthStencil, const lgcu_image *indirectLight, void *scratch, uint64_t scratchBytes,
                        const lgcu_rows *rows, void *stream) {
  int st = LGCU_OK;
  GatherArgs a;
  if (!fillGatherArgs(params, blurredDirectLight, blurredDepthMoments, normal, depthStencil, indirectLight, rows, &a, &st)) return st;
  if (!checkScratch(a, blurredDepthMoments, scratch, scratchBytes, &st)) return st;
  const cudaError_t e = launchGatherPack(a, scratch, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorInvalidValue) return fail(LGCU_ERR_INVALID_ARGUMENT, "gi_gather_pack: the two pyramids must share one memory layout");
  return cudaStatus(e, "gi_gather_pack");
}

int lgcu_gi_gather_packed(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight, const lgcu_image *blurredDepthMoments,
                          const lgcu_image *normal, const lgcu_image *depthStencil, const lgcu_image *indirectLight, const void *scratch,
                          uint64_t scratchBytes, const lgcu_rows *rows, void *stream) {
  int st = LGCU_OK;
  GatherArgs a;
  if (!fillGatherArgs(params, blurredDirectLight, blurredDepthMoments, normal, depthStencil, indirectLight, rows, &a, &st)) return st;
  if (!checkScratch(a, blurredDepthMoments, scratch, scratchBytes, &st)) return st;
  GatherTables tables;
  if (!buildGatherTables(a.viewport[0], a.viewport[1], a.light.count, &tables, &st)) return st;
  return cudaStatus(launchGatherFast(a, tables, scratch, static_cast<cudaStream_t>(stream)), "gi_gather_packed");
}

// ------------------------------------------------------------------------------------------------------- K6 / K7
static bool indirectFormat(const lgcu_image *img) {
  return img && (img->format == LGCU_FORMAT_R16G16B16A16_SFLOAT || img->format == LGCU_FORMAT_R32G32B32A32_SFLOAT);
}

int lgcu_denoise(const lgcu_denoiser_data *params, const lgcu_image *noisy, const lgcu_image *normal, const lgcu_image *depthMoments,
                 const lgcu_image *denoised, const lgcu_rows *rows, void *stream) {
  int st = LGCU_OK;
  (void)normal; // bound by the reference (SSVGIRenderer.h:294) and sampled by the shader, but it never reaches the output
  if (!params) return fail(LGCU_ERR_INVALID_ARGUMENT, "denoise: null params");
  if (!indirectFormat(noisy) || !denoised || noisy->format != denoised->format)
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "denoise: noisy/denoised formats %u / %u", noisy ? noisy->format : 0u, denoised ? denoised->format : 0u);
  if (params->radius != 0 && params->radius != 2)
    return fail(LGCU_ERR_UNSUPPORTED, "denoise: radius %d (the shader hard-codes a 4x4 window for any non-zero radius; the reference passes 0 or 2)", params->radius);
  DenoiseArgs a;
  a.format = noisy->format;
  a.radius = params->radius;
  a.viewport[0] = params->viewportExtent[0];
  a.viewport[1] = params->viewportExtent[1];
  if (!resolveLevel(noisy, 0, "noisy", &a.noisy, &st) || !resolveLevel(denoised, 0, "denoised", &a.denoised, &st)) return st;
  if (!sameSize(a.noisy, a.denoised, "noisy", "denoised", &st)) return st;
  if (a.radius != 0) {
    if (!expectFormat(depthMoments, LGCU_FORMAT_R32G32_SFLOAT, "depthMoments", &st) || !resolveLevel(depthMoments, 0, "depthMoments", &a.depthMoments, &st))
      return st;
    if (!sameSize(a.noisy, a.depthMoments, "noisy", "depthMoments", &st)) return st;
  } else {
    a.depthMoments = a.noisy;
  }
  a.rows = rowRange(rows, 0, a.denoised.h);
  return cudaStatus(launchDenoise(a, static_cast<cudaStream_t>(stream)), "denoise");
}

int lgcu_final_gather(const lgcu_final_gatherer_data *params, const lgcu_image *directLight, const lgcu_image *blurredDirectLight,
                      const lgcu_image *albedo, const lgcu_image *indirectLight, const lgcu_image *swapchain, const lgcu_rows *rows, void *stream) {
  int st = LGCU_OK;
  (void)params;             // view/proj are uploaded by the reference (SSVGIRenderer.h:321-324) but unused by the shader
  (void)blurredDirectLight; // sampled with weight 0 (finalGatherer.frag:52-57)
  if (!expectFormat(directLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, "directLight", &st) || !expectFormat(albedo, LGCU_FORMAT_R16G16B16A16_SFLOAT, "albedo", &st) ||
      !expectFormat(swapchain, LGCU_FORMAT_B8G8R8A8_SRGB, "swapchain", &st))
    return st;
  if (!indirectFormat(indirectLight)) return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "indirectLight: format %u", indirectLight ? indirectLight->format : 0u);
  FinalGatherArgs a;
  a.indirectFormat = indirectLight->format;
  if (!resolveLevel(directLight, 0, "directLight", &a.directLight, &st) || !resolveLevel(albedo, 0, "albedo", &a.albedo, &st) ||
      !resolveLevel(indirectLight, 0, "indirectLight", &a.indirect, &st) || !resolveLevel(swapchain, 0, "swapchain", &a.swapchain, &st))
    return st;
  if (!sameSize(a.swapchain, a.directLight, "swapchain", "directLight", &st) || !sameSize(a.swapchain, a.albedo, "swapchain", "albedo", &st) ||
      !sameSize(a.swapchain, a.indirect, "swapchain", "indirectLight", &st))
    return st;
  a.rows = rowRange(rows, 0, a.swapchain.h);
  return cudaStatus(launchFinalGather(a, static_cast<cudaStream_t>(stream)), "final_gather");
}

int lgcu_denoise_final_gather(const lgcu_denoiser_data *dparams, const lgcu_final_gatherer_data *fparams, const lgcu_image *noisy,
                              const lgcu_image *normal, const lgcu_image *depthMoments, const lgcu_image *denoised, const lgcu_image *directLight,
                              const lgcu_image *blurredDirectLight, const lgcu_image *albedo, const lgcu_image *swapchain, const lgcu_rows *rows,
                              void *stream) {
  int st = LGCU_OK;
  if (!dparams) return fail(LGCU_ERR_INVALID_ARGUMENT, "denoise_final_gather: null params");
  (void)fparams;            // view/proj are uploaded by the reference (SSVGIRenderer.h:321-324) but unused by the shader
  (void)blurredDirectLight; // sampled with weight 0 (finalGatherer.frag:52-57)
  (void)normal;             // sampled by the denoiser but unused in its result (denoiser.frag:50, :77)
  if (dparams->radius != 0 && dparams->radius != 2)
    return fail(LGCU_ERR_UNSUPPORTED, "denoise_final_gather: radius %d (the shader hard-codes a 4x4 window for any non-zero radius; the reference passes 0 or 2)", dparams->radius);
  if (!indirectFormat(noisy) || !denoised || noisy->format != denoised->format)
    return fail(LGCU_ERR_UNSUPPORTED_FORMAT, "denoise_final_gather: noisy/denoised formats");
  if (!expectFormat(directLight, LGCU_FORMAT_R16G16B16A16_SFLOAT, "directLight", &st) || !expectFormat(albedo, LGCU_FORMAT_R16G16B16A16_SFLOAT, "albedo", &st) ||
      !expectFormat(swapchain, LGCU_FORMAT_B8G8R8A8_SRGB, "swapchain", &st))
    return st;
  DenoiseFinalArgs a;
  a.indirectFormat = noisy->format;
  a.radius = dparams->radius;
  a.viewport[0] = dparams->viewportExtent[0];
  a.viewport[1] = dparams->viewportExtent[1];
  if (!resolveLevel(noisy, 0, "noisy", &a.noisy, &st) || !resolveLevel(denoised, 0, "denoised", &a.denoised, &st) ||
      !resolveLevel(directLight, 0, "directLight", &a.directLight, &st) || !resolveLevel(albedo, 0, "albedo", &a.albedo, &st) ||
      !resolveLevel(swapchain, 0, "swapchain", &a.swapchain, &st))
    return st;
  if (!sameSize(a.swapchain, a.noisy, "swapchain", "noisy", &st) || !sameSize(a.swapchain, a.denoised, "swapchain", "denoised", &st) ||
      !sameSize(a.swapchain, a.directLight, "swapchain", "directLight", &st) || !sameSize(a.swapchain, a.albedo, "swapchain", "albedo", &st))
    return st;
  if (a.radius != 0) {
    if (!expectFormat(depthMoments, LGCU_FORMAT_R32G32_SFLOAT, "depthMoments", &st) || !resolveLevel(depthMoments, 0, "depthMoments", &a.depthMoments, &st)) return st;
    if (!sameSize(a.noisy, a.depthMoments, "noisy", "depthMoments", &st)) return st;
  } else {
    a.depthMoments = a.noisy;
  }
  a.rows = rowRange(rows, 0, a.swapchain.h);
  return cudaStatus(launchDenoiseFinalGather(a, static_cast<cudaStream_t>(stream)), "denoise_final_gather");
}

} // extern "C"
