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
