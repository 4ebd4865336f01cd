// lgcu_device.cuh — device-side views, texel codecs and the software texture unit shared by all SSVGI kernels.
//
// Texture-unit semantics follow SURVEY.md Appendix B (clamp-to-edge, unquantised fp32 bilinear / trilinear weights,
// RGBA16F stores round-to-nearest-even). CUDA texture objects are deliberately not used: their 8-bit filter weights
// would break parity with the reference's fp32 shader arithmetic, and the images live in plain linear HBM so that
// 128-bit coalesced accesses work on every pass.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lgcu.h"

namespace lgcu {

constexpr int kMaxGatherLevels = 10; // MippedProxy uses 10 levels (src/Render/Common/MipBuilder.h:21)

// One mip level of an image in linear memory.
struct LevelView {
  unsigned char *ptr;
  uint32_t pitch; // bytes per row
  int w, h;
};

struct PyramidView {
  LevelView lv[kMaxGatherLevels];
  int count;
};

// rows [y0, y1) to process
struct RowRange {
  int y0, y1;
};

// ---- format traits -------------------------------------------------------------------------------------------
template <uint32_t F> struct Texel;

template <> struct Texel<LGCU_FORMAT_R16G16B16A16_SFLOAT> {
  static constexpr int kBytes = 8;
  __device__ __forceinline__ static float4 load(const LevelView &l, int x, int y) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(l.ptr + (size_t)y * l.pitch) + x);
    return unpack(raw);
  }
  __device__ __forceinline__ static float4 unpack(uint2 raw) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
  __device__ __forceinline__ static uint2 pack(float4 v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<const uint32_t *>(&lo);
    raw.y = *reinterpret_cast<const uint32_t *>(&hi);
    return raw;
  }
  __device__ __forceinline__ static void store(const LevelView &l, int x, int y, float4 v) {
    reinterpret_cast<uint2 *>(l.ptr + (size_t)y * l.pitch)[x] = pack(v);
  }
};

template <> struct Texel<LGCU_FORMAT_R32G32_SFLOAT> {
  static constexpr int kBytes = 8;
  __device__ __forceinline__ static float4 load(const LevelView &l, int x, int y) {
    const float2 raw = __ldg(reinterpret_cast<const float2 *>(l.ptr + (size_t)y * l.pitch) + x);
    return make_float4(raw.x, raw.y, 0.0f, 1.0f);
  }
  __device__ __forceinline__ static float4 unpack(uint2 raw) {
    return make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), 0.0f, 1.0f);
  }
  __device__ __forceinline__ static uint2 pack(float4 v) { return make_uint2(__float_as_uint(v.x), __float_as_uint(v.y)); }
  __device__ __forceinline__ static void store(const LevelView &l, int x, int y, float4 v) {
    reinterpret_cast<float2 *>(l.ptr + (size_t)y * l.pitch)[x] = make_float2(v.x, v.y);
  }
};

template <> struct Texel<LGCU_FORMAT_R32G32B32A32_SFLOAT> {
  static constexpr int kBytes = 16;
  __device__ __forceinline__ static float4 load(const LevelView &l, int x, int y) {
    return __ldg(reinterpret_cast<const float4 *>(l.ptr + (size_t)y * l.pitch) + x);
  }
  __device__ __forceinline__ static void store(const LevelView &l, int x, int y, float4 v) {
    reinterpret_cast<float4 *>(l.ptr + (size_t)y * l.pitch)[x] = v;
  }
};

template <> struct Texel<LGCU_FORMAT_D32_SFLOAT> {
  static constexpr int kBytes = 4;
  __device__ __forceinline__ static float4 load(const LevelView &l, int x, int y) {
    const float raw = __ldg(reinterpret_cast<const float *>(l.ptr + (size_t)y * l.pitch) + x);
    return make_float4(raw, 0.0f, 0.0f, 1.0f);
  }
  __device__ __forceinline__ static void store(const LevelView &l, int x, int y, float4 v) {
    reinterpret_cast<float *>(l.ptr + (size_t)y * l.pitch)[x] = v.x;
  }
};

// ---- render-target store of an fp32 RGBA value for run-time selected colour formats ------------------------------
__device__ __forceinline__ void storeColor(uint32_t format, const LevelView &l, int x, int y, float4 v) {
  if (format == LGCU_FORMAT_R16G16B16A16_SFLOAT)
    Texel<LGCU_FORMAT_R16G16B16A16_SFLOAT>::store(l, x, y, v);
  else
    Texel<LGCU_FORMAT_R32G32B32A32_SFLOAT>::store(l, x, y, v);
}
__device__ __forceinline__ float4 loadColor(uint32_t format, const LevelView &l, int x, int y) {
  if (format == LGCU_FORMAT_R16G16B16A16_SFLOAT) return Texel<LGCU_FORMAT_R16G16B16A16_SFLOAT>::load(l, x, y);
  return Texel<LGCU_FORMAT_R32G32B32A32_SFLOAT>::load(l, x, y);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// glm::min / glm::max semantics (NaN handling follows the comparison, like the reference's generated code)
__device__ __forceinline__ float glmMin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float glmMax(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float saturatef(float x) { return glmMin(glmMax(x, 0.0f), 1.0f); }

// ---- render-target conversion to B8G8R8A8_SRGB (LV/Swapchain.h:108): clamp, sRGB OETF on RGB, linear alpha, RTN to 8 bit
__device__ __forceinline__ uint32_t unorm8(float x) {
  if (!(x > 0.0f)) return 0u;
  if (x > 1.0f) x = 1.0f;
  return (uint32_t)(x * 255.0f + 0.5f);
}
__device__ __forceinline__ float linearToSrgb(float c) {
  if (!(c > 0.0f)) c = 0.0f;
  if (c > 1.0f) c = 1.0f;
  return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ uint32_t packBgra8Srgb(float4 v) {
  return unorm8(linearToSrgb(v.z)) | (unorm8(linearToSrgb(v.y)) << 8) | (unorm8(linearToSrgb(v.x)) << 16) | (unorm8(saturatef(v.w)) << 24);
}
// The same conversion with the power function on the SFU (ex2.approx(lg2.approx(c) / 2.4), relative error ~1e-6 = 3e-4 of an 8-bit
// code): the result differs from the libm-grade one only where the encoded value sits within that distance of a rounding boundary,
// and then by one code — inside the +-1 LSB bar of the swapchain. The accurate powf costs ~40 instructions per channel, which makes
// the final composite issue-bound instead of HBM-bound (ncu r01j: issue 73 %, DRAM 41 %).
__device__ __forceinline__ float linearToSrgbFast(float c) { // branch-free; c = 0 -> lg2 = -inf -> ex2 = 0, and the linear segment is selected
  c = fminf(fmaxf(c, 0.0f), 1.0f); // NaN -> 0, like the comparisons of linearToSrgb
  float l2, p;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(c)); // arguments of the power segment are >= 0.0031308: no denormal handling needed
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(l2 * (1.0f / 2.4f)));
  return c <= 0.0031308f ? 12.92f * c : 1.055f * p - 0.055f;
}
__device__ __forceinline__ uint32_t unorm8Encoded(float x) { return (uint32_t)(x * 255.0f + 0.5f); } // x in [0, 1 + ulp]: the encoder's range
template <bool kFast> __device__ __forceinline__ uint32_t packBgra8SrgbT(float4 v) {
  if (!kFast) return packBgra8Srgb(v);
  return unorm8Encoded(linearToSrgbFast(v.z)) | (unorm8Encoded(linearToSrgbFast(v.y)) << 8) | (unorm8Encoded(linearToSrgbFast(v.x)) << 16) |
         (unorm8(saturatef(v.w)) << 24);
}

// ---- exact-order bilinear / trilinear (strict kernels; compile the TU with -fmad=false) ---------------------------
// lerp(p, q, t) = p + (q - p) * t; bilinear = lerp(lerp(t00, t10, a), lerp(t01, t11, a), b).
struct BilinearTaps {
  int x0, x1, y0, y1;
  float a, b;
};

__device__ __forceinline__ BilinearTaps bilinearTaps(const LevelView &l, float u, float v) {
  BilinearTaps t;
  const float fu = u * (float)l.w - 0.5f, fv = v * (float)l.h - 0.5f;
  const float flu = floorf(fu), flv = floorf(fv);
  t.a = fu - flu;
  t.b = fv - flv;
  const int ix = (int)flu, iy = (int)flv;
  t.x0 = clampi(ix, 0, l.w - 1);
  t.x1 = clampi(ix + 1, 0, l.w - 1);
  t.y0 = clampi(iy, 0, l.h - 1);
  t.y1 = clampi(iy + 1, 0, l.h - 1);
  return t;
}

__device__ __forceinline__ float lerpExact(float p, float q, float t) { return p + (q - p) * t; }

template <uint32_t F> __device__ __forceinline__ float4 bilinear(const LevelView &l, float u, float v) {
  const BilinearTaps t = bilinearTaps(l, u, v);
  const float4 t00 = Texel<F>::load(l, t.x0, t.y0), t10 = Texel<F>::load(l, t.x1, t.y0);
  const float4 t01 = Texel<F>::load(l, t.x0, t.y1), t11 = Texel<F>::load(l, t.x1, t.y1);
  float4 r;
  r.x = lerpExact(lerpExact(t00.x, t10.x, t.a), lerpExact(t01.x, t11.x, t.a), t.b);
  r.y = lerpExact(lerpExact(t00.y, t10.y, t.a), lerpExact(t01.y, t11.y, t.a), t.b);
  r.z = lerpExact(lerpExact(t00.z, t10.z, t.a), lerpExact(t01.z, t11.z, t.a), t.b);
  r.w = lerpExact(lerpExact(t00.w, t10.w, t.a), lerpExact(t01.w, t11.w, t.a), t.b);
  return r;
}

// ---- small vector helpers in glm evaluation order -------------------------------------------------------------
struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 x, V3 y) {
  return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
// glm::normalize = v * inversesqrt(dot(v, v)), inversesqrt = 1 / sqrt  (IEEE sqrt and divide)
__device__ __forceinline__ V3 normalize3(V3 a) { return a * (1.0f / sqrtf(dot3(a, a))); }

struct Mat4 {
  float m[16]; // column-major
};
// glm mat4 * vec4: (c0*x + c1*y) + (c2*z + c3*w)
__device__ __forceinline__ float4 mulMat4(const Mat4 &M, float x, float y, float z, float w) {
  float4 r;
  r.x = (M.m[0] * x + M.m[4] * y) + (M.m[8] * z + M.m[12] * w);
  r.y = (M.m[1] * x + M.m[5] * y) + (M.m[9] * z + M.m[13] * w);
  r.z = (M.m[2] * x + M.m[6] * y) + (M.m[10] * z + M.m[14] * w);
  r.w = (M.m[3] * x + M.m[7] * y) + (M.m[11] * z + M.m[15] * w);
  return r;
}
// Unproject(): SH/Common/directLighting.frag:24-29, SH/SSVGI/indirectLighting.frag:20-25
__device__ __forceinline__ V3 unproject(float sx, float sy, float sz, const Mat4 &inv) {
  const float4 v = mulMat4(inv, sx * 2.0f - 1.0f, sy * 2.0f - 1.0f, sz, 1.0f);
  return V3{v.x / v.w, v.y / v.w, v.z / v.w};
}

} // namespace lgcu
