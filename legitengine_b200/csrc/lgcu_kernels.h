// lgcu_kernels.h — host-callable launchers of the SSVGI CUDA kernels (internal to liblgcu.so).
// Argument blocks carry already-validated device views and the per-frame constants that the C ABI layer
// (lgcu_api.cu) hoists out of the reference's per-fragment code.
#pragma once

#include "lgcu_device.cuh"

namespace lgcu {

struct ClearValues {
  float color[4];
  float depth;
};

struct GBufferArgs {
  const lgcu_fragment *fragments;
  uint64_t fragmentPitch;
  const lgcu_draw_call_data *objects;
  uint32_t nObjects;
  float cam[3]; // (inverse(viewMatrix) * (0,0,0,1)).xyz  — gBufferBuilder.frag:30
  ClearValues clear;
  LevelView albedo, emissive, normal, depthMoments, depthStencil;
  RowRange rows;
};

struct DirectLightArgs {
  Mat4 invViewProj;   // inverse(projMatrix * viewMatrix)          directLighting.frag:50-52
  Mat4 lightViewProj; // lightProjMatrix * lightViewMatrix         directLighting.frag:60
  Mat4 lightView;     //                                            directLighting.frag:62
  float lightPos[3];  // (inverse(lightViewMatrix) * (0,0,0,1)).xyz directLighting.frag:54
  LevelView albedo, emissive, normal, depthStencil, shadowMap, directLight;
  RowRange rows;
};

struct MipLevelArgs {
  uint32_t format;
  int depthFilter = 0; // MipLevelBuilderData.filterType >= 0.5 (MipBuilder::FilterTypes::Depth)
  LevelView src, dst;
  RowRange rows; // in dst rows
};

struct BlurLevelArgs {
  uint32_t format;
  int sizeX, sizeY, radius; // BlurLayerBuilderData
  LevelView src, dst;
  RowRange rows; // in dst rows
};

struct ChainArgs { // fused K3+K4 over one chain
  uint32_t format;
  int levels; // levels present (<= kMaxGatherLevels)
  int radius;
  PyramidView chain, blurred;
  RowRange rows; // base rows
};

constexpr int kGatherDirs = 4;     // indirectLighting.frag:166
constexpr int kGatherMaxSteps = 16; // table capacity for the march (8 steps are reached at 8K)

// Per-frame tables of the GI gather. Everything in here is a pure function of viewportSize.x and is evaluated on
// the host with the same libm calls the reference arithmetic makes, so pattern / step / LOD selection is
// bit-identical to the oracle ("int pattern index, iters, mip level select bit-exact").
struct GatherTables {
  float dirX[16][kGatherDirs], dirY[16][kGatherDirs]; // cos/sin(1.57075*ang + 1.57075*d)   :177-178
  float pixelOffset[16][kGatherMaxSteps];             // near*pow(2.57075, k+lin)+1-near      :217
  float lod[16][kGatherMaxSteps];                     // log(max(0,...))/ln2 - 2, clamped to [0, levels-1]  :234-235
  float iterThreshold[kGatherMaxSteps];               // smallest |tmax| for which iterationsCount > n  :212
  int maxSteps;
};

struct GatherArgs {
  Mat4 invViewProj; // :123
  float cam[3];     // :127
  float viewport[2];
  uint32_t outFormat; // RGBA16F or RGBA32F
  PyramidView light;   // blurredDirectLight  (RGBA16F)
  PyramidView moments; // blurredDepthMoments (RG32F)
  LevelView normal, depthStencil, indirect;
  RowRange rows;
};

struct DenoiseArgs {
  uint32_t format; // noisy/denoised format
  int radius;
  float viewport[2];
  LevelView noisy, depthMoments, denoised;
  RowRange rows;
};

struct FinalGatherArgs {
  uint32_t indirectFormat;
  LevelView directLight, albedo, indirect, swapchain;
  RowRange rows;
};

struct DenoiseFinalArgs { // fused K6 + K7
  uint32_t indirectFormat;
  int radius;        // 0: centre tap (copy), 2: 4x4 depth-guided fit
  LevelView depthMoments; // radius 2 only
  float viewport[2]; // DenoiserData.viewportExtent (the centre tap divides gl_FragCoord by it, denoiser.frag:83)
  LevelView noisy, denoised, directLight, albedo, swapchain;
  RowRange rows;
};

struct GBufferLightArgs { // fused K1 + K2
  GBufferArgs g;
  DirectLightArgs l;
};

constexpr int kFrontMipLevels = 4; // mip levels 1..4 are built inside the frame-front kernel's 64x16 tiles

struct FrontArgs { // fused K1 + K2 + level-0 blur copies + mip levels 1..kFrontMipLevels of both chains
  GBufferArgs g;
  DirectLightArgs l;
  LevelView blurLight0, blurMoments0;
  LevelView lightMip[kFrontMipLevels], momentsMip[kFrontMipLevels]; // levels 1..4
  int mipLevels;                                                    // how many of them exist (0..4)
  int tilesX, tilesY, tileCount;                                    // filled by the launcher
};

struct ChainsArgs { // everything of K3 + K4 that the front kernel leaves: blur of levels >= 1, mips above kFrontMipLevels
  PyramidView light, blurredLight, moments, blurredMoments;
  int levels;     // levels present in the chains
  int radius;     // blur radius of levels >= 1 (1 or 2)
  int gridLevels; // levels 1..gridLevels are blurred by the grid (they were built by the front kernel)
  RowRange rows;  // base rows
};

struct RasterArgs { // ShadowPass / raster half of GBufferPass (k_raster.cu)
  lgcu_mesh_scene scene; // device arrays
  Mat4 viewProj;         // projMatrix * viewMatrix  (gBufferBuilder.vert:36, left-associative product)
  int width, height;
  RowRange rows;
  void *scratch;            // rasterScratchBytes(scene.nTriangles, width, height)
  lgcu_fragment *fragments; // G-buffer target: fragment buffer; nullptr selects the depth-only target
  uint64_t fragmentPitch;
  LevelView depth;          // shadow map (D32F)
};
uint64_t rasterScratchBytes(uint32_t nTriangles, uint32_t width, uint32_t height);
cudaError_t launchRaster(const RasterArgs &a, int smCount, cudaStream_t s);

struct InterleaveArgs { // (de)interleave passes: a texel permutation between two images of one format and size
  int texelBytes;           // 8 or 16
  int gridX, gridY;         // gridSize
  int cellsX, cellsY;       // viewportSize / gridSize (integer division): size of one de-interleaved sub-image
  LevelView interleaved, deinterleaved;
  RowRange rows;            // destination rows
};
cudaError_t launchDeinterleave(const InterleaveArgs &a, cudaStream_t s); // deinterleaved <- interleaved
cudaError_t launchInterleave(const InterleaveArgs &a, cudaStream_t s);   // interleaved <- deinterleaved

struct DebugOverlayArgs { // one quad of DebugInfoPass
  uint32_t srcFormat, targetFormat;
  float wx0, wy0, wx1, wy1; // window-space rectangle of the quad ("rule D")
  int x0, x1, y0, y1;       // conservative pixel box of the rectangle, clipped to the target and to the rows
  LevelView src, target;
};
cudaError_t launchDebugOverlay(const DebugOverlayArgs &a, cudaStream_t s);

constexpr int kMaxCopies = 64, kMaxFlags = 32;
struct RowCopyArgs { // contiguous slabs (16-byte multiples), any of them possibly in a peer GPU's memory
  const void *src[kMaxCopies];
  void *dst[kMaxCopies];
  uint64_t unitEnd[kMaxCopies]; // running end of each slab in 16-byte units
  int count;
};
struct FlagArgs {
  uint32_t *flags[kMaxFlags];
  int count;
};
struct ExchangeArgs { // one fused exchange step: signalBefore -> wait -> copies -> (last CTA) signalAfter
  FlagArgs signalBefore, wait, signalAfter;
  RowCopyArgs copies;
  uint32_t *frame, *done;
  int lag, bump;
};
cudaError_t launchExchange(const ExchangeArgs &a, int smCount, cudaStream_t s);
cudaError_t launchRowCopies(const RowCopyArgs &a, int smCount, cudaStream_t s);
cudaError_t launchBumpFrame(uint32_t *frame, cudaStream_t s);
cudaError_t launchSignal(const FlagArgs &a, const uint32_t *frame, cudaStream_t s);
cudaError_t launchWait(const FlagArgs &a, const uint32_t *frame, int lag, cudaStream_t s);

cudaError_t launchFrameFront(const FrontArgs &a, int smCount, cudaStream_t s);
cudaError_t launchFrameChains(const ChainsArgs &a, cudaStream_t s);
cudaError_t launchGBufferResolve(const GBufferArgs &a, cudaStream_t s);
cudaError_t launchDirectLight(const DirectLightArgs &a, cudaStream_t s);
cudaError_t launchGBufferDirectLight(const GBufferLightArgs &a, cudaStream_t s);
cudaError_t launchMipLevel(const MipLevelArgs &a, cudaStream_t s);
cudaError_t launchBlurLevel(const BlurLevelArgs &a, cudaStream_t s);
cudaError_t launchMipBlurChain(const ChainArgs &a, cudaStream_t s);
cudaError_t launchGatherStrict(const GatherArgs &a, const GatherTables &t, cudaStream_t s);
// scratch == nullptr: individual depth taps; else the quad-packed depth pyramid built by launchGatherPack for the same views
cudaError_t launchGatherFast(const GatherArgs &a, const GatherTables &t, const void *scratch, cudaStream_t s);
cudaError_t launchGatherPack(const GatherArgs &a, void *scratch, cudaStream_t s);
uint64_t gatherScratchBytes(uint32_t width, uint32_t height, uint32_t mips);
cudaError_t launchDenoise(const DenoiseArgs &a, cudaStream_t s);
cudaError_t launchFinalGather(const FinalGatherArgs &a, cudaStream_t s);
cudaError_t launchDenoiseFinalGather(const DenoiseFinalArgs &a, cudaStream_t s);

} // namespace lgcu
