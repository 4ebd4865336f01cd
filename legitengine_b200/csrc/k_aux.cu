// k_aux.cu — the passes either side of the hot path as sm_100a kernels:
//   interleaved rendering (InterleaveBuilder.h:14-80; SH/Common/deinterleave.frag:17-29, SH/Common/interleave.frag:16-28) and
//   one quad of the debug overlay (DebugRenderer.h:13-63; SH/Common/debugRenderer.vert:15-24, debugRenderer.frag:11-15).
//
// (De)interleave is a pure texel permutation, HBM-bound at 2 x texel bytes per pixel. Both kernels walk the INTERLEAVED image in
// 16-byte vectors (two 8-byte texels or one 16-byte texel per thread, x fastest): that side is then read or written with fully
// used 128-bit accesses, and on the de-interleaved side the lanes of a warp that share a pattern index touch consecutive texels
// (gridSize.x runs of 32 / gridSize.x texels per warp), so every 32-byte sector that moves is fully used there too. A scatter by
// source is only a bijection where viewportSize is divisible by gridSize; the de-interleave pass therefore handles the ragged
// right / bottom border (destination pixels at or beyond gridSize * (viewportSize / gridSize)) by a gather per destination
// pixel with the shader's formula. Index arithmetic is int32 like the shaders'.
//
// Compiled with -fmad=false: the overlay's coverage test, texture coordinate and bilinear filter are evaluated in the oracle's order.
#include "lgcu_kernels.h"

namespace lgcu {

namespace {

constexpr int kThreads = 256;
constexpr uint32_t F16 = LGCU_FORMAT_R16G16B16A16_SFLOAT, F32 = LGCU_FORMAT_R32G32B32A32_SFLOAT, RG32 = LGCU_FORMAT_R32G32_SFLOAT;

template <int B> struct RawTexel;
template <> struct RawTexel<8> { using type = uint2; };
template <> struct RawTexel<16> { using type = uint4; };

template <int B> __device__ __forceinline__ typename RawTexel<B>::type loadRaw(const LevelView &l, int x, int y) {
  return __ldg(reinterpret_cast<const typename RawTexel<B>::type *>(l.ptr + (size_t)y * l.pitch) + x);
}
template <int B> __device__ __forceinline__ void storeRaw(const LevelView &l, int x, int y, typename RawTexel<B>::type v) {
  reinterpret_cast<typename RawTexel<B>::type *>(l.ptr + (size_t)y * l.pitch)[x] = v;
}

// interleave.frag:16-21  DeinterleavePixel(interleavedPixel) = (p % grid) * (viewport / grid) + p / grid
__device__ __forceinline__ int deinterleavedCoord(int p, int grid, int cells) { return (p % grid) * cells + p / grid; }
// deinterleave.frag:17-22  InterleavePixel(deinterleavedPixel) = (p % (viewport / grid)) * grid + p / (viewport / grid)
__device__ __forceinline__ int interleavedCoord(int p, int grid, int cells) { return (p % cells) * grid + p / cells; }

// Interleave pass: dst = interleaved image, one thread per 16-byte destination vector; rows = interleaved (destination) rows.
template <int B> __global__ void __launch_bounds__(kThreads) interleaveKernel(const __grid_constant__ InterleaveArgs a) {
  constexpr int kPer = 16 / B;
  const int vx = blockIdx.x * kThreads + threadIdx.x, y = a.rows.y0 + blockIdx.y;
  const int x = vx * kPer;
  if (x >= a.interleaved.w) return;
  const int sy = deinterleavedCoord(y, a.gridY, a.cellsY);
  if (kPer == 1) {
    storeRaw<B>(a.interleaved, x, y, loadRaw<B>(a.deinterleaved, deinterleavedCoord(x, a.gridX, a.cellsX), sy));
  } else {
    const uint2 t0 = loadRaw<8>(a.deinterleaved, deinterleavedCoord(x, a.gridX, a.cellsX), sy);
    if (x + 1 < a.interleaved.w) {
      const uint2 t1 = loadRaw<8>(a.deinterleaved, deinterleavedCoord(x + 1, a.gridX, a.cellsX), sy);
      reinterpret_cast<uint4 *>(a.interleaved.ptr + (size_t)y * a.interleaved.pitch)[vx] = make_uint4(t0.x, t0.y, t1.x, t1.y);
    } else {
      storeRaw<8>(a.interleaved, x, y, t0);
    }
  }
}

// De-interleave pass, divisible part: src = interleaved image, one thread per 16-byte source vector, scattered to the destination
// pixel that reads it. blockIdx.y walks the SOURCE rows whose destination row lies in [rows.y0, rows.y1).
template <int B> __global__ void __launch_bounds__(kThreads) deinterleaveScatterKernel(const __grid_constant__ InterleaveArgs a) {
  constexpr int kPer = 16 / B;
  const int vx = blockIdx.x * kThreads + threadIdx.x, sy = blockIdx.y;
  const int x = vx * kPer;
  const int coveredW = a.gridX * a.cellsX; // source pixels below this are read by exactly one regular destination pixel
  if (x >= coveredW) return;
  const int dy = deinterleavedCoord(sy, a.gridY, a.cellsY); // the inverse of InterleavePixel on the divisible part
  if (dy < a.rows.y0 || dy >= a.rows.y1) return;
  if (kPer == 1) {
    storeRaw<B>(a.deinterleaved, deinterleavedCoord(x, a.gridX, a.cellsX), dy, loadRaw<B>(a.interleaved, x, sy));
  } else if (x + 1 < a.interleaved.w) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.interleaved.ptr + (size_t)sy * a.interleaved.pitch) + vx);
    storeRaw<8>(a.deinterleaved, deinterleavedCoord(x, a.gridX, a.cellsX), dy, make_uint2(v.x, v.y));
    if (x + 1 < coveredW) storeRaw<8>(a.deinterleaved, deinterleavedCoord(x + 1, a.gridX, a.cellsX), dy, make_uint2(v.z, v.w));
  } else {
    storeRaw<8>(a.deinterleaved, deinterleavedCoord(x, a.gridX, a.cellsX), dy, loadRaw<8>(a.interleaved, x, sy));
  }
}

// De-interleave pass, ragged border: destination pixels of the box [bx0, bx1) x [by0, by1), gathered with the shader's formula.
template <int B> __global__ void __launch_bounds__(kThreads) deinterleaveGatherKernel(const __grid_constant__ InterleaveArgs a, int bx0, int bx1, int by0, int by1) {
  const int x = bx0 + blockIdx.x * kThreads + threadIdx.x, y = by0 + blockIdx.y;
  if (x >= bx1 || y >= by1) return;
  storeRaw<B>(a.deinterleaved, x, y, loadRaw<B>(a.interleaved, interleavedCoord(x, a.gridX, a.cellsX), interleavedCoord(y, a.gridY, a.cellsY)));
}

template <int B> cudaError_t launchDeinterleaveT(const InterleaveArgs &a, cudaStream_t s) {
  constexpr int kPer = 16 / B;
  const int coveredW = a.gridX * a.cellsX, coveredH = a.gridY * a.cellsY, w = a.deinterleaved.w;
  const int vectors = (coveredW + kPer - 1) / kPer;
  deinterleaveScatterKernel<B><<<dim3((vectors + kThreads - 1) / kThreads, coveredH), kThreads, 0, s>>>(a);
  // right border: columns [coveredW, w), every row of the range; bottom border: rows [coveredH, h), columns [0, coveredW)
  const int y0 = a.rows.y0, y1 = a.rows.y1;
  if (coveredW < w) deinterleaveGatherKernel<B><<<dim3((w - coveredW + kThreads - 1) / kThreads, y1 - y0), kThreads, 0, s>>>(a, coveredW, w, y0, y1);
  const int b0 = y0 > coveredH ? y0 : coveredH;
  if (b0 < y1) deinterleaveGatherKernel<B><<<dim3((coveredW + kThreads - 1) / kThreads, y1 - b0), kThreads, 0, s>>>(a, 0, coveredW, b0, y1);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ debug overlay
template <uint32_t F> __device__ __forceinline__ float4 sampleLevel0(const LevelView &src, float u, float v) { return bilinear<F>(src, u, v); }

__global__ void __launch_bounds__(kThreads) debugOverlayKernel(const __grid_constant__ DebugOverlayArgs a) {
  const int x = a.x0 + blockIdx.x * 32 + (threadIdx.x & 31), y = a.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.x1 || y >= a.y1) return;
  const float cx = (float)x + 0.5f, cy = (float)y + 0.5f;
  if (!(a.wx0 <= cx && cx < a.wx1) || !(a.wy0 <= cy && cy < a.wy1)) return; // top-left rule on the axis-aligned quad
  const float tu = (cx - a.wx0) / (a.wx1 - a.wx0), tv = (cy - a.wy0) / (a.wy1 - a.wy0); // interpolated fragTexCoord
  float4 c; // debugRenderer.frag:13 texture(srcSampler, fragTexCoord): single-level view -> bilinear at level 0
  if (a.srcFormat == F16)
    c = sampleLevel0<F16>(a.src, tu, tv);
  else if (a.srcFormat == F32)
    c = sampleLevel0<F32>(a.src, tu, tv);
  else
    c = sampleLevel0<RG32>(a.src, tu, tv);
  if (a.targetFormat == LGCU_FORMAT_B8G8R8A8_SRGB)
    reinterpret_cast<uint32_t *>(a.target.ptr + (size_t)y * a.target.pitch)[x] = packBgra8Srgb(c);
  else
    storeColor(a.targetFormat, a.target, x, y, c);
}

} // namespace

cudaError_t launchInterleave(const InterleaveArgs &a, cudaStream_t s) {
  if (a.rows.y1 <= a.rows.y0 || a.interleaved.w <= 0) return cudaSuccess;
  const int per = 16 / a.texelBytes, vectors = (a.interleaved.w + per - 1) / per;
  const dim3 grid((vectors + kThreads - 1) / kThreads, a.rows.y1 - a.rows.y0);
  if (a.texelBytes == 8)
    interleaveKernel<8><<<grid, kThreads, 0, s>>>(a);
  else
    interleaveKernel<16><<<grid, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launchDeinterleave(const InterleaveArgs &a, cudaStream_t s) {
  if (a.rows.y1 <= a.rows.y0 || a.deinterleaved.w <= 0) return cudaSuccess;
  return a.texelBytes == 8 ? launchDeinterleaveT<8>(a, s) : launchDeinterleaveT<16>(a, s);
}

cudaError_t launchDebugOverlay(const DebugOverlayArgs &a, cudaStream_t s) {
  if (a.x1 <= a.x0 || a.y1 <= a.y0) return cudaSuccess;
  debugOverlayKernel<<<dim3((a.x1 - a.x0 + 31) / 32, (a.y1 - a.y0 + 7) / 8), kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

} // namespace lgcu
