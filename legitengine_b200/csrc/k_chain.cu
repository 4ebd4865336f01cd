// k_chain.cu — mip build + blur of one whole chain (K3+K4) behind a single entry point.
//
// MipBuilder::BuildMips (src/Render/Common/MipBuilder.h:142-181) followed by the ten BlurBuilder::ApplyBlur passes
// of SSVGIRenderer::RenderFrame (:209-221) for one MippedProxy. This first version issues the per-level kernels of
// k_streaming.cu back to back on the stream (19 launches instead of 19 render passes with barriers).
#include "lgcu_kernels.h"

namespace lgcu {

cudaError_t launchMipBlurChain(const ChainArgs &a, cudaStream_t s) {
  auto levelRows = [&](int level, int h) {
    const int y0 = a.rows.y0 >> level, y1 = (a.rows.y1 + ((1 << level) - 1)) >> level;
    return RowRange{y0 < h ? y0 : h, y1 < h ? y1 : h};
  };
  for (int l = 1; l < a.levels; l++) {
    MipLevelArgs m;
    m.format = a.format;
    m.depthFilter = 0; // BuildMips(..., FilterTypes::Avg) on the live path (SSVGIRenderer.h:207-208)
    m.src = a.chain.lv[l - 1];
    m.dst = a.chain.lv[l];
    m.rows = levelRows(l, m.dst.h);
    cudaError_t e = launchMipLevel(m, s);
    if (e != cudaSuccess) return e;
  }
  for (int l = 0; l < a.levels; l++) {
    BlurLevelArgs b;
    b.format = a.format;
    b.src = a.chain.lv[l];
    b.dst = a.blurred.lv[l];
    b.sizeX = b.src.w;
    b.sizeY = b.src.h;
    b.radius = l == 0 ? 0 : a.radius;
    b.rows = levelRows(l, b.dst.h);
    cudaError_t e = launchBlurLevel(b, s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

} // namespace lgcu

// ---------------------------------------------------------------------------------------------------------------------
// frameChainsKernel — what remains of K3 + K4 after the frame-front kernel (k_front.cu), in ONE launch:
//   grid part : blurLayerBuilder (radius 2, SH/Common/blurLayerBuilder.frag:17-35) of levels 1..gridLevels of both chains;
//   tail part : one thread-block CLUSTER of 8 CTAs per chain (chainTailKernel) builds mip levels gridLevels+1..9
//               (mipLevelBuilder.frag:17-28) and blurs them. Level l+1 needs all of level l, so the levels are walked one after
//               the other with a hardware cluster barrier (release / acquire at cluster scope) in between. These levels hold
//               < 2 % of the texels; the cluster replaces 20 launch-latency-bound launches, and its 2048 threads keep this
//               serial chain short on a multi-GPU strip, where it costs as much as on a whole frame.
// Blur: a thread produces four vertically adjacent output texels of one column from a 4 x 7 register window (28 loads
// instead of 64), each output summed in the shader's order (x outer, y inner, sequential fp32) so the result is bit-exact.
#include <cooperative_groups.h>

#include <cstdlib>

namespace lgcu {
namespace {

constexpr uint32_t kF16 = LGCU_FORMAT_R16G16B16A16_SFLOAT, kRG32 = LGCU_FORMAT_R32G32_SFLOAT;
constexpr int kBlurTileW = 32, kBlurTileH = 32, kChainThreads = 256, kBlurRowsPerThread = 4;
constexpr int kTailCluster = 8; // CTAs per chain in the tail cluster (portable cluster size limit)

struct ChainsLaunch {
  ChainsArgs a;
  int blockBegin[2][kFrontMipLevels + 1]; // first CTA of (chain, level - 1); [..][gridLevels] = end
  int rowBegin[kFrontMipLevels], rowEnd[kFrontMipLevels];
};

template <bool kCoherent> __device__ __forceinline__ uint2 loadTexel(const LevelView &l, int x, int y) {
  const uint2 *p = reinterpret_cast<const uint2 *>(l.ptr + (size_t)y * l.pitch) + x;
  return kCoherent ? __ldcg(p) : __ldg(p);
}

// blurLayerBuilder.frag:20-31 for the outputs (x, y..y+rows-1) of one column
template <uint32_t F, int R, bool kCoherent>
__device__ __forceinline__ void blurColumn(const LevelView &src, const LevelView &dst, int x, int y, int rows) {
  constexpr int kWin = 2 * R, kSpan = kBlurRowsPerThread + kWin - 1;
  uint2 t[kWin][kSpan];
#pragma unroll
  for (int i = 0; i < kWin; i++) {
    const int sx = clampi(x + i - R, 0, src.w - 1);
#pragma unroll
    for (int j = 0; j < kSpan; j++) t[i][j] = loadTexel<kCoherent>(src, sx, clampi(y + j - R, 0, src.h - 1));
  }
#pragma unroll
  for (int o = 0; o < kBlurRowsPerThread; o++) {
    if (o >= rows) break;
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float totalWeight = 0.0f;
#pragma unroll
    for (int i = 0; i < kWin; i++)
#pragma unroll
      for (int j = 0; j < kWin; j++) {
        const float4 v = Texel<F>::unpack(t[i][o + j]);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        totalWeight += 1.0f;
      }
    Texel<F>::store(dst, x, y + o, make_float4(sum.x / totalWeight, sum.y / totalWeight, sum.z / totalWeight, sum.w / totalWeight));
  }
}

template <uint32_t F, bool kCoherent> __device__ __forceinline__ void blurColumnR(int radius, const LevelView &src, const LevelView &dst, int x, int y, int rows) {
  if (radius == 2)
    blurColumn<F, 2, kCoherent>(src, dst, x, y, rows);
  else
    blurColumn<F, 1, kCoherent>(src, dst, x, y, rows);
}

template <uint32_t F> __device__ __forceinline__ void mipTexel(const LevelView &src, const LevelView &dst, int x, int y) {
  const uint4 top = __ldcg(reinterpret_cast<const uint4 *>(src.ptr + (size_t)(2 * y) * src.pitch) + x);
  const uint4 bot = __ldcg(reinterpret_cast<const uint4 *>(src.ptr + (size_t)(2 * y + 1) * src.pitch) + x);
  const float4 s00 = Texel<F>::unpack(make_uint2(top.x, top.y)), s10 = Texel<F>::unpack(make_uint2(top.z, top.w));
  const float4 s01 = Texel<F>::unpack(make_uint2(bot.x, bot.y)), s11 = Texel<F>::unpack(make_uint2(bot.z, bot.w));
  float4 sum;
  sum.x = ((((0.0f + s00.x) + s10.x) + s01.x) + s11.x) / 4.0f;
  sum.y = ((((0.0f + s00.y) + s10.y) + s01.y) + s11.y) / 4.0f;
  sum.z = ((((0.0f + s00.z) + s10.z) + s01.z) + s11.z) / 4.0f;
  sum.w = ((((0.0f + s00.w) + s10.w) + s01.w) + s11.w) / 4.0f;
  Texel<F>::store(dst, x, y, sum);
}

template <uint32_t F> __device__ void chainTail(const ChainsArgs &a, const PyramidView &chain, const PyramidView &blurred, int ctaRank) {
  namespace cg = cooperative_groups;
  const int first = ctaRank * kChainThreads + threadIdx.x, stride = kTailCluster * kChainThreads;
  // mip levels one after the other (each reads the level the cluster has just written: __ldcg, i.e. from L2), then their blurs
  for (int l = a.gridLevels + 1; l < a.levels; l++) {
    const LevelView &src = chain.lv[l - 1], &dst = chain.lv[l];
    for (int i = first; i < dst.w * dst.h; i += stride) mipTexel<F>(src, dst, i % dst.w, i / dst.w);
    cg::this_cluster().sync();
  }
  for (int l = a.gridLevels + 1; l < a.levels; l++) {
    const LevelView &src = chain.lv[l], &dst = blurred.lv[l];
    const int groups = (src.h + kBlurRowsPerThread - 1) / kBlurRowsPerThread;
    for (int i = first; i < src.w * groups; i += stride) {
      const int x = i % src.w, y = (i / src.w) * kBlurRowsPerThread;
      blurColumnR<F, true>(a.radius, src, dst, x, y, min(kBlurRowsPerThread, src.h - y));
    }
  }
}

// grid = 2 clusters of kTailCluster CTAs: cluster 0 walks the directLight chain, cluster 1 the depthMoments chain
__global__ void __cluster_dims__(kTailCluster, 1, 1) __launch_bounds__(kChainThreads) chainTailKernel(const __grid_constant__ ChainsArgs a) {
  const int chain = blockIdx.x / kTailCluster, ctaRank = blockIdx.x % kTailCluster;
  if (chain == 0)
    chainTail<kF16>(a, a.light, a.blurredLight, ctaRank);
  else
    chainTail<kRG32>(a, a.moments, a.blurredMoments, ctaRank);
}

// A/B variant of the blur grid (LGCU_BLUR_SMEM=1): the CTA's 32 x 32 output tile and its halo (window -R..R-1: 35 x 35 texels for
// R = 2) are staged in shared memory with coalesced loads, and the register windows are filled from there. Measured against the
// direct version below (whose overlapping window loads are served by L1): see profiles/README.md, round 2.
template <uint32_t F, int R> __device__ __forceinline__ void blurTileStaged(uint2 (*tile)[kBlurTileW + 2 * R], const LevelView &src, const LevelView &dst, int x0, int y0,
                                                                              int rowEnd) {
  constexpr int kW = kBlurTileW + 2 * R - 1, kH = kBlurTileH + 2 * R - 1; // texels the tile's windows touch
  for (int i = threadIdx.x; i < kW * kH; i += kChainThreads) {
    const int tx = i % kW, ty = i / kW;
    tile[ty][tx] = loadTexel<false>(src, clampi(x0 + tx - R, 0, src.w - 1), clampi(y0 + ty - R, 0, src.h - 1));
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = (threadIdx.x >> 5) * kBlurRowsPerThread;
  const int x = x0 + lx, y = y0 + ly;
  if (x >= src.w || y >= rowEnd) return;
  const int rows = min(kBlurRowsPerThread, rowEnd - y);
  constexpr int kWin = 2 * R;
#pragma unroll
  for (int o = 0; o < kBlurRowsPerThread; o++) {
    if (o >= rows) break;
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float totalWeight = 0.0f;
#pragma unroll
    for (int i = 0; i < kWin; i++)
#pragma unroll
      for (int j = 0; j < kWin; j++) {
        const float4 v = Texel<F>::unpack(tile[ly + o + j][lx + i]);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        totalWeight += 1.0f;
      }
    Texel<F>::store(dst, x, y + o, make_float4(sum.x / totalWeight, sum.y / totalWeight, sum.z / totalWeight, sum.w / totalWeight));
  }
}

__global__ void __launch_bounds__(kChainThreads) frameChainsStagedKernel(const __grid_constant__ ChainsLaunch p) {
  __shared__ uint2 tile[kBlurTileH + 3][kBlurTileW + 4];
  const ChainsArgs &a = p.a;
  int b = blockIdx.x;
  const int chain = b >= p.blockBegin[1][0] ? 1 : 0;
  int l = 1;
  while (l < a.gridLevels && b >= p.blockBegin[chain][l]) l++;
  b -= p.blockBegin[chain][l - 1];
  const LevelView &src = chain ? a.moments.lv[l] : a.light.lv[l], &dst = chain ? a.blurredMoments.lv[l] : a.blurredLight.lv[l];
  const int bx = (src.w + kBlurTileW - 1) / kBlurTileW;
  const int x0 = (b % bx) * kBlurTileW, y0 = p.rowBegin[l - 1] + (b / bx) * kBlurTileH;
  if (chain)
    blurTileStaged<kRG32, 2>(tile, src, dst, x0, y0, p.rowEnd[l - 1]);
  else
    blurTileStaged<kF16, 2>(tile, src, dst, x0, y0, p.rowEnd[l - 1]);
}

__device__ __forceinline__ void blurGridCta(const ChainsLaunch &p, int b) {
  const ChainsArgs &a = p.a;
  const int chain = b >= p.blockBegin[1][0] ? 1 : 0;
  int l = 1;
  while (l < a.gridLevels && b >= p.blockBegin[chain][l]) l++; // level l occupies [blockBegin[l-1], blockBegin[l])
  b -= p.blockBegin[chain][l - 1];
  const LevelView &src = chain ? a.moments.lv[l] : a.light.lv[l], &dst = chain ? a.blurredMoments.lv[l] : a.blurredLight.lv[l];
  const int bx = (src.w + kBlurTileW - 1) / kBlurTileW;
  const int x = (b % bx) * kBlurTileW + (threadIdx.x & 31);
  const int y = p.rowBegin[l - 1] + (b / bx) * kBlurTileH + (threadIdx.x >> 5) * kBlurRowsPerThread;
  if (x >= src.w || y >= p.rowEnd[l - 1]) return;
  const int rows = min(kBlurRowsPerThread, p.rowEnd[l - 1] - y);
  if (chain)
    blurColumnR<kRG32, false>(a.radius, src, dst, x, y, rows);
  else
    blurColumnR<kF16, false>(a.radius, src, dst, x, y, rows);
}

__global__ void __launch_bounds__(kChainThreads) frameChainsKernel(const __grid_constant__ ChainsLaunch p) { blurGridCta(p, blockIdx.x); }

// Tail and blur grid in ONE launch: the whole grid is launched in clusters of kTailCluster CTAs, the first two clusters are the tail
// (they are dispatched first and walk their serial chain of small levels for ~16 us), every other CTA is a blur tile and ignores its
// cluster. The two parts are data-independent (the tail reads chain level gridLevels and writes levels above it, the grid writes the
// blurred levels 1..gridLevels), so the blur tiles fill the rest of the GPU while the tail runs instead of waiting behind it
// (two launches on one stream: 15.5 us + 36 us at 4K, ncu r02k — and the tail costs the same on a multi-GPU strip).
// blurBlocks CTAs of blur work follow the tail; the grid is padded to a multiple of the cluster size.
__global__ void __cluster_dims__(kTailCluster, 1, 1) __launch_bounds__(kChainThreads, 3) frameChainsWithTailKernel(const __grid_constant__ ChainsLaunch p, int blurBlocks) {
  if (blockIdx.x < 2 * kTailCluster) {
    const int chain = blockIdx.x / kTailCluster, ctaRank = blockIdx.x % kTailCluster;
    if (chain == 0)
      chainTail<kF16>(p.a, p.a.light, p.a.blurredLight, ctaRank);
    else
      chainTail<kRG32>(p.a, p.a.moments, p.a.blurredMoments, ctaRank);
    return;
  }
  const int b = blockIdx.x - 2 * kTailCluster;
  if (b < blurBlocks) blurGridCta(p, b);
}

} // namespace

cudaError_t launchFrameChains(const ChainsArgs &a, cudaStream_t s) {
  ChainsLaunch p;
  p.a = a;
  int blocks = 0;
  for (int chain = 0; chain < 2; chain++) {
    for (int l = 1; l <= a.gridLevels; l++) {
      const LevelView &lv = a.light.lv[l];
      const int y0 = a.rows.y0 >> l, y1 = (a.rows.y1 + ((1 << l) - 1)) >> l;
      p.rowBegin[l - 1] = y0 < lv.h ? y0 : lv.h;
      p.rowEnd[l - 1] = y1 < lv.h ? y1 : lv.h;
      p.blockBegin[chain][l - 1] = blocks;
      const int rows = p.rowEnd[l - 1] - p.rowBegin[l - 1];
      if (rows > 0) blocks += ((lv.w + kBlurTileW - 1) / kBlurTileW) * ((rows + kBlurTileH - 1) / kBlurTileH);
    }
    p.blockBegin[chain][a.gridLevels] = blocks;
  }
  if (a.gridLevels == 0) p.blockBegin[1][0] = 0x7fffffff;
  static const bool staged = getenv("LGCU_BLUR_SMEM") && atoi(getenv("LGCU_BLUR_SMEM")) != 0; // A/B switch (radius 2 only)
  static const bool split = getenv("LGCU_CHAINS_SPLIT") && atoi(getenv("LGCU_CHAINS_SPLIT")) != 0; // A/B switch: tail and grid as two launches
  const bool tail = a.levels > a.gridLevels + 1;
  if (tail && blocks > 0 && !staged && !split) {
    const int grid = 2 * kTailCluster + (blocks + kTailCluster - 1) / kTailCluster * kTailCluster;
    frameChainsWithTailKernel<<<grid, kChainThreads, 0, s>>>(p, blocks);
    return cudaGetLastError();
  }
  if (tail) chainTailKernel<<<2 * kTailCluster, kChainThreads, 0, s>>>(a);
  if (blocks > 0 && staged && a.radius == 2)
    frameChainsStagedKernel<<<blocks, kChainThreads, 0, s>>>(p);
  else if (blocks > 0)
    frameChainsKernel<<<blocks, kChainThreads, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace lgcu
