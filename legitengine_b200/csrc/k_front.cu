// k_front.cu — the front of the fused SSVGI frame in one pass over the fragments:
//   K1 G-buffer resolve (gBufferBuilder.frag:28-38) + K2 direct lighting (directLighting.frag:45-83)
//   + the two radius-0 "blur" passes of level 0 (blurLayerBuilder.frag:17-35 with radius 0 is a copy; SSVGIRenderer.h:209-221)
//   + mip levels 1..4 of both chains (mipLevelBuilder.frag:17-28; MipBuilder::BuildMips, MipBuilder.h:142-181).
//
// HBM-bound: 32 B/px read, 60 B/px written at level 0 plus 5.3 B/px of mip levels, against 76 + 32 + 2x13.3 B/px for the
// same outputs as separate passes. Design:
//  * persistent CTAs (a small multiple of the SM count) walk 64x16-pixel tiles; the per-draw-call pow(colour, 2.2) table is
//    evaluated once per CTA into shared memory instead of once per tile;
//  * a thread owns a 2x2 pixel quad, a warp a 32x4 block (16x2 quads): every image row is written with 16-byte stores,
//    256 contiguous bytes per half warp;
//  * level 1 is reduced in registers, level 2 with warp shuffles (lanes l, l^1, l^16, l^17), levels 3 and 4 through 1 KiB of
//    shared memory. Every level is rounded to its storage format (fp16 for the light chain) before it feeds the next one
//    and the four taps are summed in the shader's order, so the result is bit-identical to the nine separate passes.
// Compiled with -fmad=false (shader-order fp32, see lgcu_shading.cuh).
#include <cstdlib>

#include "lgcu_shading.cuh"

namespace lgcu {

namespace {

using namespace shading;

constexpr int kTileW = 64, kTileH = 16, kThreads = 256;
constexpr int kFrontBlocksPerSmDefault = 4; // resident CTAs per SM (register budget); measured r01n at 4K: 2 -> 0.220 ms, 3 -> 0.189, 4 -> 0.182

struct MipTexel { // one texel of both chains, in storage form
  uint2 light;    // RGBA16F
  float2 moments; // RG32F
};

// mipLevelBuilder.frag:19-28: (((s(2x,2y) + s(2x+1,2y)) + s(2x,2y+1)) + s(2x+1,2y+1)) / 4, from the STORED values of the level above
__device__ __forceinline__ MipTexel average4(const MipTexel &s00, const MipTexel &s10, const MipTexel &s01, const MipTexel &s11) {
  const float4 a = Texel<F16>::unpack(s00.light), b = Texel<F16>::unpack(s10.light), c = Texel<F16>::unpack(s01.light), d = Texel<F16>::unpack(s11.light);
  float4 sum;
  sum.x = ((((0.0f + a.x) + b.x) + c.x) + d.x) / 4.0f;
  sum.y = ((((0.0f + a.y) + b.y) + c.y) + d.y) / 4.0f;
  sum.z = ((((0.0f + a.z) + b.z) + c.z) + d.z) / 4.0f;
  sum.w = ((((0.0f + a.w) + b.w) + c.w) + d.w) / 4.0f;
  MipTexel r;
  r.light = Texel<F16>::pack(sum);
  r.moments.x = ((((0.0f + s00.moments.x) + s10.moments.x) + s01.moments.x) + s11.moments.x) / 4.0f;
  r.moments.y = ((((0.0f + s00.moments.y) + s10.moments.y) + s01.moments.y) + s11.moments.y) / 4.0f;
  return r;
}

__device__ __forceinline__ MipTexel shuffleXor(const MipTexel &t, int mask) {
  MipTexel r;
  r.light.x = __shfl_xor_sync(0xffffffffu, t.light.x, mask);
  r.light.y = __shfl_xor_sync(0xffffffffu, t.light.y, mask);
  r.moments.x = __shfl_xor_sync(0xffffffffu, t.moments.x, mask);
  r.moments.y = __shfl_xor_sync(0xffffffffu, t.moments.y, mask);
  return r;
}

__device__ __forceinline__ void storeMip(const FrontArgs &a, int level, int x, int y, const MipTexel &t) { // level 1..4
  const LevelView &lv = a.lightMip[level - 1], &mv = a.momentsMip[level - 1];
  reinterpret_cast<uint2 *>(lv.ptr + (size_t)y * lv.pitch)[x] = t.light;
  reinterpret_cast<float2 *>(mv.ptr + (size_t)y * mv.pitch)[x] = t.moments;
}

// 8-byte texels of two horizontally adjacent pixels of one image row (x even): one 16-byte store when both exist
__device__ __forceinline__ void storePair(const LevelView &l, int x, int y, uint2 t0, uint2 t1, bool has1) {
  unsigned char *dst = l.ptr + (size_t)y * l.pitch + (size_t)x * 8;
  if (has1)
    *reinterpret_cast<uint4 *>(dst) = make_uint4(t0.x, t0.y, t1.x, t1.y);
  else
    *reinterpret_cast<uint2 *>(dst) = t0;
}
__device__ __forceinline__ uint2 asBits(float2 v) { return make_uint2(__float_as_uint(v.x), __float_as_uint(v.y)); }

// kMinBlocks = resident CTAs per SM the register allocation aims for: 2 -> 118 registers, 3 -> 78, 4 -> 64 (16 bytes of spills). The
// kernel is bound by latency (a long per-pixel chain behind four 16-byte loads), so more resident warps beat fewer spills.
// (Requesting the thread's second fragment row and its next tile's first row with prefetch.global.L1 ahead of the first row's arithmetic
// was measured: 0.192 ms against 0.182 ms at 4K, r04b — the extra L1 traffic costs more than the hidden latency.)
template <int kMinBlocks> __global__ void __launch_bounds__(kThreads, kMinBlocks) frameFrontKernel(const __grid_constant__ FrontArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ MipTexel s2[kTileH / 4][kTileW / 4]; // level-2 texels of the tile (4 x 16)
  __shared__ MipTexel s3[kTileH / 8][kTileW / 8]; // level-3 texels (2 x 8)
  const ObjectColors *table = stageObjectTable(a.g, reinterpret_cast<ObjectColors *>(smemRaw));

  const int W = a.g.albedo.w, y1 = a.g.rows.y1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int QX = (warp & 1) * 16 + (lane & 15), QY = (warp >> 1) * 2 + (lane >> 4); // quad inside the tile: 32 x 8

  for (int tile = blockIdx.x; tile < a.tileCount; tile += gridDim.x) {
    const int tileX = (tile % a.tilesX) * kTileW, tileY = a.g.rows.y0 + (tile / a.tilesX) * kTileH;
    const int x0 = tileX + 2 * QX, y0 = tileY + 2 * QY;
    // ---- level 0: resolve, light, store (K1, K2, blur radius 0) ----------------------------------------------------------
    MipTexel px[2][2] = {}; // [row][column], storage form of directLight / depthMoments level 0
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int y = y0 + r;
      if (y >= y1 || x0 >= W) continue;
      const bool has1 = x0 + 1 < W;
      ResolvedTexel t0 = resolveFragment(a.g, table, x0, y), t1 = t0;
      if (has1) t1 = resolveFragment(a.g, table, x0 + 1, y);
      const uint2 lit0 = Texel<F16>::pack(shadeDirect(a.l, x0, y, Texel<F16>::unpack(t0.albedo), Texel<F16>::unpack(t0.emissive), Texel<F16>::unpack(t0.normal), t0.depth));
      uint2 lit1 = lit0;
      if (has1) lit1 = Texel<F16>::pack(shadeDirect(a.l, x0 + 1, y, Texel<F16>::unpack(t1.albedo), Texel<F16>::unpack(t1.emissive), Texel<F16>::unpack(t1.normal), t1.depth));
      storePair(a.g.albedo, x0, y, t0.albedo, t1.albedo, has1);
      storePair(a.g.emissive, x0, y, t0.emissive, t1.emissive, has1);
      storePair(a.g.normal, x0, y, t0.normal, t1.normal, has1);
      storePair(a.g.depthMoments, x0, y, asBits(t0.moments), asBits(t1.moments), has1);
      storePair(a.blurMoments0, x0, y, asBits(t0.moments), asBits(t1.moments), has1); // blurLayerBuilder radius 0
      storePair(a.l.directLight, x0, y, lit0, lit1, has1);
      storePair(a.blurLight0, x0, y, lit0, lit1, has1);
      float *dz = reinterpret_cast<float *>(a.g.depthStencil.ptr + (size_t)y * a.g.depthStencil.pitch) + x0;
      if (has1)
        *reinterpret_cast<float2 *>(dz) = make_float2(t0.depth, t1.depth);
      else
        *dz = t0.depth;
      px[r][0].light = lit0, px[r][0].moments = t0.moments;
      px[r][1].light = lit1, px[r][1].moments = t1.moments;
    }
    if (a.mipLevels < 1) continue;
    // ---- level 1: the thread's own quad -------------------------------------------------------------------------------------
    const int x1 = x0 >> 1, yl1 = y0 >> 1;
    MipTexel m1 = average4(px[0][0], px[0][1], px[1][0], px[1][1]); // garbage where the quad is cut; never stored, never used
    if (x1 < a.lightMip[0].w && yl1 < a.lightMip[0].h && y0 + 1 < y1) storeMip(a, 1, x1, yl1, m1);
    if (a.mipLevels < 2) continue;
    // ---- level 2: 2x2 quads of one warp ---------------------------------------------------------------------------------------
    const MipTexel m1r = shuffleXor(m1, 1), m1d = shuffleXor(m1, 16), m1rd = shuffleXor(m1, 17);
    const MipTexel m2 = average4(m1, m1r, m1d, m1rd);
    const bool owner2 = ((lane & 1) == 0) && ((lane & 16) == 0);
    const int x2 = x0 >> 2, yl2 = y0 >> 2;
    if (owner2) {
      if (x2 < a.lightMip[1].w && yl2 < a.lightMip[1].h && y0 + 3 < y1) storeMip(a, 2, x2, yl2, m2);
      s2[QY >> 1][QX >> 1] = m2;
    }
    if (a.mipLevels < 3) continue; // uniform
    __syncthreads();
    // ---- level 3 (8 x 2 texels per tile) and level 4 (4 x 1) through shared memory ------------------------------------------------
    if (threadIdx.x < 16) {
      const int cx = threadIdx.x & 7, cy = threadIdx.x >> 3;
      const MipTexel m3 = average4(s2[2 * cy][2 * cx], s2[2 * cy][2 * cx + 1], s2[2 * cy + 1][2 * cx], s2[2 * cy + 1][2 * cx + 1]);
      const int x3 = (tileX >> 3) + cx, y3 = (tileY >> 3) + cy;
      if (x3 < a.lightMip[2].w && y3 < a.lightMip[2].h && tileY + 8 * cy + 7 < y1) storeMip(a, 3, x3, y3, m3);
      s3[cy][cx] = m3;
    }
    if (a.mipLevels >= 4) {
      __syncthreads();
      if (threadIdx.x < 4) {
        const int cx = threadIdx.x;
        const MipTexel m4 = average4(s3[0][2 * cx], s3[0][2 * cx + 1], s3[1][2 * cx], s3[1][2 * cx + 1]);
        const int x4 = (tileX >> 4) + cx, y4 = tileY >> 4;
        if (x4 < a.lightMip[3].w && y4 < a.lightMip[3].h && tileY + 15 < y1) storeMip(a, 4, x4, y4, m4);
      }
    } else {
      __syncthreads(); // s2 is rewritten by the next tile
    }
  }
}

} // namespace

cudaError_t launchFrameFront(const FrontArgs &args, int smCount, cudaStream_t s) {
  FrontArgs a = args;
  if (a.g.rows.y1 <= a.g.rows.y0) return cudaSuccess;
  a.tilesX = (a.g.albedo.w + kTileW - 1) / kTileW;
  a.tilesY = (a.g.rows.y1 - a.g.rows.y0 + kTileH - 1) / kTileH;
  a.tileCount = a.tilesX * a.tilesY;
  const size_t smem = a.g.nObjects <= (uint32_t)kMaxSharedObjects ? (size_t)a.g.nObjects * sizeof(ObjectColors) : 0;
  static const int blocksPerSm = getenv("LGCU_FRONT_BLOCKS") ? atoi(getenv("LGCU_FRONT_BLOCKS")) : kFrontBlocksPerSmDefault; // development switch: 2, 3, 4
  int grid = smCount * blocksPerSm * 2; // the resident CTAs per SM, two rounds: evens out the tail without re-staging the table often
  if (grid > a.tileCount) grid = a.tileCount;
  if (blocksPerSm == 3)
    frameFrontKernel<3><<<grid, kThreads, smem, s>>>(a);
  else if (blocksPerSm == 4)
    frameFrontKernel<4><<<grid, kThreads, smem, s>>>(a);
  else
    frameFrontKernel<2><<<grid, kThreads, smem, s>>>(a);
  return cudaGetLastError();
}

} // namespace lgcu
