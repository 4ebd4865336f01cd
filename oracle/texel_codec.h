/*
 * texel_codec.h — TEST INFRASTRUCTURE (part of the CPU oracle; never linked into the product path).
 *
 * Texel decode/encode for the image formats the SSVGI path uses, and addressing of lgcu_image on HOST memory.
 * Store rules follow SURVEY.md Appendix B: RGBA16F = fp32 -> fp16 round-to-nearest-even (no clamp),
 * RG32F / D32F exact, B8G8R8A8_SRGB = clamp[0,1] -> sRGB OETF on RGB, linear alpha -> round(255 x).
 * Reference formats: src/Render/Renderers/SSVGIRenderer.h:393-404, LV/Swapchain.h:108.
 */
#ifndef LGCU_ORACLE_TEXEL_CODEC_H
#define LGCU_ORACLE_TEXEL_CODEC_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/lgcu.h"

#ifdef __cplusplus
extern "C" {
#endif

static inline uint32_t orc_f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float orc_bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* IEEE binary32 -> binary16, round to nearest even, overflow to inf, NaN preserved (quiet). */
static inline uint16_t orc_f32_to_f16(float f) {
  uint32_t x = orc_f32_bits(f);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t abs = x & 0x7FFFFFFFu;
  if (abs >= 0x7F800000u) { /* inf / nan */
    return (uint16_t)(sign | 0x7C00u | (abs > 0x7F800000u ? (0x0200u | ((abs >> 13) & 0x03FFu)) : 0u));
  }
  if (abs >= 0x477FF000u) { /* >= 65520 rounds to inf */
    return (uint16_t)(sign | 0x7C00u);
  }
  if (abs < 0x38800000u) { /* subnormal half or zero: |f| < 2^-14 */
    if (abs < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 */
    uint32_t exp = abs >> 23;
    uint32_t mant = (abs & 0x007FFFFFu) | 0x00800000u;
    uint32_t shift = 126u - exp; /* 14..24: target unit is 2^-24 */
    uint32_t half = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1u);
    if (rem > halfway || (rem == halfway && (half & 1u))) half++;
    return (uint16_t)(sign | half);
  }
  {
    uint32_t half = ((abs - 0x38000000u) >> 13);
    uint32_t rem = abs & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;
    return (uint16_t)(sign | half);
  }
}

static inline float orc_f16_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu;
  uint32_t mant = h & 0x03FFu;
  if (exp == 0) {
    if (mant == 0) return orc_bits_f32(sign);
    /* subnormal: value = mant * 2^-24 */
    float v = (float)mant * 5.9604644775390625e-08f;
    return orc_bits_f32(sign | orc_f32_bits(v));
  }
  if (exp == 31) return orc_bits_f32(sign | 0x7F800000u | (mant << 13));
  return orc_bits_f32(sign | ((exp + 112u) << 23) | (mant << 13));
}

static inline uint8_t orc_unorm8(float x) { /* x already clamped to [0,1]; NaN -> 0 */
  if (!(x > 0.0f)) return 0;
  if (x > 1.0f) x = 1.0f;
  return (uint8_t)(x * 255.0f + 0.5f);
}

static inline float orc_linear_to_srgb(float c) {
  if (!(c > 0.0f)) c = 0.0f;
  if (c > 1.0f) c = 1.0f;
  return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}

static inline uint32_t orc_texel_size(uint32_t format) {
  switch (format) {
    case LGCU_FORMAT_B8G8R8A8_SRGB: return 4;
    case LGCU_FORMAT_R16G16B16A16_SFLOAT: return 8;
    case LGCU_FORMAT_R32G32_SFLOAT: return 8;
    case LGCU_FORMAT_R32G32B32A32_SFLOAT: return 16;
    case LGCU_FORMAT_D32_SFLOAT: return 4;
    default: return 0;
  }
}

/* size of VIEW level `lod` (image level baseMip + lod) */
static inline void orc_level_size(const lgcu_image *img, uint32_t lod, int *w, int *h) {
  uint32_t l = img->baseMip + lod;
  *w = (int)(img->width >> l);
  *h = (int)(img->height >> l);
}

static inline const uint8_t *orc_texel_ptr(const lgcu_image *img, uint32_t lod, int x, int y) {
  uint32_t l = img->baseMip + lod;
  return (const uint8_t *)img->base + img->levelOffset[l] + (uint64_t)y * img->levelPitch[l] +
         (uint64_t)x * orc_texel_size(img->format);
}

/* texel fetch with GLSL component defaults (missing g,b -> 0, a -> 1) */
static inline void orc_load_texel(const lgcu_image *img, uint32_t lod, int x, int y, float out[4]) {
  const uint8_t *p = orc_texel_ptr(img, lod, x, y);
  switch (img->format) {
    case LGCU_FORMAT_R16G16B16A16_SFLOAT: {
      uint16_t h[4]; memcpy(h, p, 8);
      out[0] = orc_f16_to_f32(h[0]); out[1] = orc_f16_to_f32(h[1]);
      out[2] = orc_f16_to_f32(h[2]); out[3] = orc_f16_to_f32(h[3]);
    } break;
    case LGCU_FORMAT_R32G32_SFLOAT: {
      float f[2]; memcpy(f, p, 8);
      out[0] = f[0]; out[1] = f[1]; out[2] = 0.0f; out[3] = 1.0f;
    } break;
    case LGCU_FORMAT_R32G32B32A32_SFLOAT: memcpy(out, p, 16); break;
    case LGCU_FORMAT_D32_SFLOAT: {
      float f; memcpy(&f, p, 4);
      out[0] = f; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1.0f;
    } break;
    default: out[0] = out[1] = out[2] = 0.0f; out[3] = 1.0f; break;
  }
}

/* render-target store of an fp32 RGBA value */
static inline void orc_store_texel(const lgcu_image *img, uint32_t lod, int x, int y, const float v[4]) {
  uint8_t *p = (uint8_t *)orc_texel_ptr(img, lod, x, y);
  switch (img->format) {
    case LGCU_FORMAT_R16G16B16A16_SFLOAT: {
      uint16_t h[4] = {orc_f32_to_f16(v[0]), orc_f32_to_f16(v[1]), orc_f32_to_f16(v[2]), orc_f32_to_f16(v[3])};
      memcpy(p, h, 8);
    } break;
    case LGCU_FORMAT_R32G32_SFLOAT: memcpy(p, v, 8); break;
    case LGCU_FORMAT_R32G32B32A32_SFLOAT: memcpy(p, v, 16); break;
    case LGCU_FORMAT_D32_SFLOAT: memcpy(p, v, 4); break;
    case LGCU_FORMAT_B8G8R8A8_SRGB: {
      float a = v[3];
      if (!(a > 0.0f)) a = 0.0f;
      if (a > 1.0f) a = 1.0f;
      uint8_t b[4] = {orc_unorm8(orc_linear_to_srgb(v[2])), orc_unorm8(orc_linear_to_srgb(v[1])),
                      orc_unorm8(orc_linear_to_srgb(v[0])), orc_unorm8(a)};
      memcpy(p, b, 4);
    } break;
    default: break;
  }
}

static inline int orc_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Bilinear fetch at view level `lod`: u = uv.x*w - 0.5, i0 = floor(u), alpha = u - i0,
 * clamp-to-edge indices, lerp(lerp(t00,t10,a), lerp(t01,t11,a), b) with lerp(p,q,t) = p + (q - p) * t. */
static inline void orc_bilinear(const lgcu_image *img, uint32_t lod, float u_, float v_, float out[4]) {
  int w, h;
  orc_level_size(img, lod, &w, &h);
  float u = u_ * (float)w - 0.5f, v = v_ * (float)h - 0.5f;
  float fu = floorf(u), fv = floorf(v);
  float a = u - fu, b = v - fv;
  int x0 = orc_clampi((int)fu, 0, w - 1), x1 = orc_clampi((int)fu + 1, 0, w - 1);
  int y0 = orc_clampi((int)fv, 0, h - 1), y1 = orc_clampi((int)fv + 1, 0, h - 1);
  float t00[4], t10[4], t01[4], t11[4];
  orc_load_texel(img, lod, x0, y0, t00);
  orc_load_texel(img, lod, x1, y0, t10);
  orc_load_texel(img, lod, x0, y1, t01);
  orc_load_texel(img, lod, x1, y1, t11);
  for (int c = 0; c < 4; c++) {
    float top = t00[c] + (t10[c] - t00[c]) * a;
    float bot = t01[c] + (t11[c] - t01[c]) * a;
    out[c] = top + (bot - top) * b;
  }
}

/* textureLod with linear mip filter on a mipCount-level view: lambda = clamp(lod, 0, mipCount-1),
 * d = floor(lambda), delta = lambda - d, result = (1-delta)*bilinear(d) + delta*bilinear(min(d+1, last)).
 * -inf clamps to 0; NaN is treated as 0 (cannot occur on the live path, indirectLighting.frag:234-235). */
static inline void orc_texture_lod(const lgcu_image *img, float u, float v, float lod, float out[4]) {
  float last = (float)(img->mipCount - 1);
  float lambda = lod;
  if (!(lambda > 0.0f)) lambda = 0.0f;
  if (lambda > last) lambda = last;
  float fd = floorf(lambda);
  float delta = lambda - fd;
  uint32_t d = (uint32_t)fd;
  uint32_t d1 = d + 1 < img->mipCount ? d + 1 : img->mipCount - 1;
  float lo[4], hi[4];
  orc_bilinear(img, d, u, v, lo);
  orc_bilinear(img, d1, u, v, hi);
  for (int c = 0; c < 4; c++) out[c] = (1.0f - delta) * lo[c] + delta * hi[c];
}

/* sampler2DShadow, compare LESS_OR_EQUAL, linear filter: bilinear blend of the four 0/1 compare results. */
static inline float orc_texture_shadow(const lgcu_image *img, float u_, float v_, float ref) {
  int w, h;
  orc_level_size(img, 0, &w, &h);
  float u = u_ * (float)w - 0.5f, v = v_ * (float)h - 0.5f;
  float fu = floorf(u), fv = floorf(v);
  float a = u - fu, b = v - fv;
  int x0 = orc_clampi((int)fu, 0, w - 1), x1 = orc_clampi((int)fu + 1, 0, w - 1);
  int y0 = orc_clampi((int)fv, 0, h - 1), y1 = orc_clampi((int)fv + 1, 0, h - 1);
  float t[4], c00, c10, c01, c11;
  orc_load_texel(img, 0, x0, y0, t); c00 = ref <= t[0] ? 1.0f : 0.0f;
  orc_load_texel(img, 0, x1, y0, t); c10 = ref <= t[0] ? 1.0f : 0.0f;
  orc_load_texel(img, 0, x0, y1, t); c01 = ref <= t[0] ? 1.0f : 0.0f;
  orc_load_texel(img, 0, x1, y1, t); c11 = ref <= t[0] ? 1.0f : 0.0f;
  float top = c00 + (c10 - c00) * a;
  float bot = c01 + (c11 - c01) * a;
  return top + (bot - top) * b;
}

#ifdef __cplusplus
}
#endif
#endif
