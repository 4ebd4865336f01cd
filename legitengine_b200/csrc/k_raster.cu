// k_raster.cu — rasterisation front end for sm_100a: "ShadowPass" (SSVGIRenderer.h:63-104) and the raster half of "GBufferPass"
// (:107-158), which the reference leaves to the fixed-function rasteriser.
//
// Pipeline (all enqueue-only, graph-capturable, no host round trip):
//   setup   one thread per triangle: vertex stage (SH/Common/gBufferBuilder.vert:32-40 / shadowmapBuilder.vert:32-40) of its three
//           vertices, homogeneous edge equations in fp64 (no clipping: triangles crossing the eye plane rasterise correctly), cull,
//           conservative bounding box (near-plane clipped polygon when a vertex is behind the eye), record -> scratch. Triangles whose
//           box exceeds 32x32 pixels reserve a contiguous run of 32x32 work tiles with one 64-bit atomic (record index in the high
//           word, tile count in the low word, so records are sorted by their first tile).
//   small   one warp per triangle with a box of at most 32x32 pixels: lanes cover the box row-major.
//   big     persistent warps stride over the reserved tiles; a warp finds its triangle with a 32-ary search over the record
//           starts, rejects the tile if an edge function is negative at all four corners, else walks its 32 rows.
//   both write the visibility buffer with atomicMin on (depth bits << 32 | triangle id): depth test LESS in draw order without any
//   ordering between warps (equal depth keeps the earlier triangle, like in-order LESS).
//   resolve one thread per pixel: re-evaluates the winner's edge functions, interpolates vertWorldPos / vertWorldNormal
//           perspective-correctly and writes the lgcu_fragment the fragment stage reads; the shadow pass writes D32F depth.
//   tile    (scenes of at most kTileLoopTriangles triangles, instead of big + resolve) one warp per 32x16 SCREEN tile: the warp culls
//           the big triangles against its tile 32 at a time, rasterises the survivors into a per-lane register column of 16 keys
//           (no atomics, no visibility-buffer round trip per pixel), merges the small triangles' keys from the visibility buffer
//           and resolves in place. The big path's global atomics and dependent L2 round trips made it latency-bound (ncu r01j:
//           29 % issue utilisation, 205 us at 4K); the tile path is plain fp64 arithmetic with 16 independent rows in flight.
//
// The arithmetic of coverage, depth and interpolation is "rule R" (restated in oracle/raster_oracle.c): every
// per-pixel quantity is a pure function of (triangle record, x, y) evaluated in a fixed order without FMA contraction
// (this unit is compiled with -fmad=false), so the result does not depend on how the work is distributed and matches the oracle bit for bit.
#include <cstdlib>

#include "lgcu_kernels.h"

namespace lgcu {

namespace {

constexpr unsigned long long kEmpty = ~0ull;
constexpr int kTile = 32; // work tile of the big-triangle path, and the largest box of the small path

struct __align__(16) TriRecord {
  double a[3], b[3], c[3]; // edge equations, interior positive
  double Z[3], W[3];       // clip z, w of the vertices
  float wp[3][3], wn[3][3]; // vertWorldPos, vertWorldNormal of the vertices
  uint32_t objectId;
  int x0, y0, x1, y1; // inclusive pixel box; x1 < x0: culled
  uint32_t big;       // 1: handled by the tile path
  uint32_t pad[2];
};
static_assert(sizeof(TriRecord) == 224, "TriRecord layout");

struct __align__(16) BigRecord {
  uint32_t tri, firstTile, tilesX, tileCount;
  int x0, y0, x1, y1; // copy of the triangle's pixel box: the tile path culls on it without touching the 224-byte record
};
static_assert(sizeof(BigRecord) == 32, "BigRecord layout");

struct RasterKernelArgs {
  const lgcu_vertex *vertices;
  const uint32_t *indices;
  const lgcu_draw *draws;
  const lgcu_draw_call_data *objects;
  uint32_t nVertices, nIndices, nDraws, nObjects, nTriangles;
  Mat4 viewProj; // projMatrix * viewMatrix (glm order, host)
  int width, height;
  RowRange rows;
  TriRecord *tris;
  BigRecord *big;
  unsigned long long *counter; // high word: big records, low word: tiles
  unsigned long long *vis;     // width * height keys
};

__device__ __forceinline__ void mulMat4(const float *m, float x, float y, float z, float w, float r[4]) { // glm mat4 * vec4
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = (m[i] * x + m[4 + i] * y) + (m[8 + i] * z + m[12 + i] * w);
}

__device__ __forceinline__ int clampToInt(double v, int lo, int hi) { return v < (double)lo ? lo : (v > (double)hi ? hi : (int)v); }

// coverage + depth of one pixel under rule R
__device__ __forceinline__ bool shadePixel(const double a[3], const double b[3], const double c[3], const double Z[3], const double W[3], int x, int y,
                                           double e[3], float *depth) {
  const double px = (double)x + 0.5, py = (double)y + 0.5;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    e[i] = (a[i] * px + b[i] * py) + c[i];
    if (!(e[i] > 0.0 || (e[i] == 0.0 && (a[i] > 0.0 || (a[i] == 0.0 && b[i] > 0.0))))) return false;
  }
  const double zn = (e[0] * Z[0] + e[1] * Z[1]) + e[2] * Z[2];
  const double wn = (e[0] * W[0] + e[1] * W[1]) + e[2] * W[2];
  if (!(wn > 0.0) || !(zn >= 0.0) || !(zn <= wn)) return false;
  float d = (float)(zn / wn);
  if (d <= 0.0f) d = 0.0f;
  *depth = d;
  return true;
}

// The same rule as shadePixel + the LESS test against the cleared depth, without control flow: the visibility key of the pixel or
// kEmpty. The tile path evaluates 16 rows per triangle; with no branches between them their fp64 dependency chains (edge
// functions -> zn, wn -> IEEE divide) overlap instead of running one after the other.
__device__ __forceinline__ unsigned long long shadePixelKey(const double a[3], const double b[3], const double c[3], const double Z[3], const double W[3], int x,
                                                            int y, uint32_t tri, bool inBox) {
  const double px = (double)x + 0.5, py = (double)y + 0.5;
  bool ok = inBox;
  double e[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    e[i] = (a[i] * px + b[i] * py) + c[i];
    ok = ok & ((e[i] > 0.0) | ((e[i] == 0.0) & ((a[i] > 0.0) | ((a[i] == 0.0) & (b[i] > 0.0)))));
  }
  const double zn = (e[0] * Z[0] + e[1] * Z[1]) + e[2] * Z[2];
  const double wn = (e[0] * W[0] + e[1] * W[1]) + e[2] * W[2];
  ok = ok & (wn > 0.0) & (zn >= 0.0) & (zn <= wn);
  float d = (float)(zn / wn); // meaningless where !ok, and then discarded
  d = d <= 0.0f ? 0.0f : d;
  ok = ok & (d < 1.0f);
  return ok ? (((unsigned long long)__float_as_uint(d) << 32) | tri) : kEmpty;
}

__device__ __forceinline__ void writeVisibility(unsigned long long *vis, int width, int x, int y, float depth, uint32_t tri) {
  if (!(depth < 1.0f)) return; // LESS against the cleared 1.0
  const unsigned long long key = ((unsigned long long)__float_as_uint(depth) << 32) | tri;
  unsigned long long *p = vis + (size_t)y * width + x;
  if (key < *reinterpret_cast<volatile unsigned long long *>(p)) atomicMin(p, key);
}

__global__ void __launch_bounds__(128) rasterSetupKernel(const __grid_constant__ RasterKernelArgs A) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nTriangles) return;
  TriRecord &R = A.tris[t];
  // draw call of this triangle: the last draw with firstTriangle <= t
  uint32_t lo = 0, hi = A.nDraws;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (A.draws[mid].firstTriangle <= t) lo = mid; else hi = mid;
  }
  const lgcu_draw dr = A.draws[lo];
  const uint32_t local = t - dr.firstTriangle;
  R.x0 = 0; R.x1 = -1; R.y0 = 0; R.y1 = -1; R.big = 0;
  if (dr.objectId >= A.nObjects || 3 * local + 2 >= dr.indexCount || dr.firstIndex + 3 * local + 2 >= A.nIndices) return;
  const float *model = A.objects[dr.objectId].modelMatrix.m;
  double X[3], Y[3], Z[3], W[3];
  const double hw = 0.5 * (double)A.width, hh = 0.5 * (double)A.height;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint32_t vi = A.indices[dr.firstIndex + 3 * local + k] + dr.vertexOffset;
    if (vi >= A.nVertices) return;
    const float4 *vp = reinterpret_cast<const float4 *>(A.vertices + vi); // 32-byte vertices, 16-byte aligned buffer (checked by the ABI)
    const float4 v0 = __ldg(vp), v1 = __ldg(vp + 1);                      // pos.xyz normal.x | normal.yz uv
    float wp[4], wn[4], clip[4];
    mulMat4(model, v0.x, v0.y, v0.z, 1.0f, wp); // gBufferBuilder.vert:34
    mulMat4(model, v0.w, v1.x, v1.y, 0.0f, wn); // :35
    mulMat4(A.viewProj.m, wp[0], wp[1], wp[2], 1.0f, clip);          // :36
#pragma unroll
    for (int c = 0; c < 3; c++) { R.wp[k][c] = wp[c]; R.wn[k][c] = wn[c]; }
    X[k] = ((double)clip[0] + (double)clip[3]) * hw;
    Y[k] = ((double)clip[1] + (double)clip[3]) * hh;
    Z[k] = (double)clip[2];
    W[k] = (double)clip[3];
  }
  double a[3], b[3], c[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    a[i] = Y[j] * W[k] - W[j] * Y[k];
    b[i] = W[j] * X[k] - X[j] * W[k];
    c[i] = X[j] * Y[k] - Y[j] * X[k];
  }
  const double det = (a[0] * X[0] + b[0] * Y[0]) + c[0] * W[0];
  if (!(det != 0.0) || !isfinite(det)) return;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    R.a[i] = det < 0.0 ? -a[i] : a[i];
    R.b[i] = det < 0.0 ? -b[i] : b[i];
    R.c[i] = det < 0.0 ? -c[i] : c[i];
    R.Z[i] = Z[i];
    R.W[i] = W[i];
  }
  R.objectId = dr.objectId;
  // trivially outside the clip volume in z: no pixel can pass 0 <= zn <= wn
  if ((Z[0] < 0.0 && Z[1] < 0.0 && Z[2] < 0.0) || (Z[0] > W[0] && Z[1] > W[1] && Z[2] > W[2])) return;
  // conservative pixel box of the visible part: vertices in front of the near plane project directly; an edge that crosses
  // z = 0 contributes its intersection point (where w = near > 0 for a perspective matrix)
  double minx = 1e300, maxx = -1e300, miny = 1e300, maxy = -1e300;
  bool whole = false;
  int points = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int n = (k + 1) % 3;
    if (Z[k] >= 0.0) {
      if (W[k] > 0.0) {
        const double px = X[k] / W[k], py = Y[k] / W[k];
        minx = fmin(minx, px); maxx = fmax(maxx, px); miny = fmin(miny, py); maxy = fmax(maxy, py);
        points++;
      } else {
        whole = true;
      }
    }
    if ((Z[k] >= 0.0) != (Z[n] >= 0.0)) {
      const double s = Z[k] / (Z[k] - Z[n]);
      const double ix = X[k] + s * (X[n] - X[k]), iy = Y[k] + s * (Y[n] - Y[k]), iw = W[k] + s * (W[n] - W[k]);
      if (iw > 0.0) {
        const double px = ix / iw, py = iy / iw;
        minx = fmin(minx, px); maxx = fmax(maxx, px); miny = fmin(miny, py); maxy = fmax(maxy, py);
        points++;
      } else {
        whole = true;
      }
    }
  }
  int x0 = 0, x1 = A.width - 1, y0 = A.rows.y0, y1 = A.rows.y1 - 1;
  if (!whole) {
    if (points == 0 || !isfinite(minx) || !isfinite(maxx) || !isfinite(miny) || !isfinite(maxy)) return;
    // the clipped polygon is exact only up to rounding: 2 pixels of margin
    x0 = max(x0, clampToInt(floor(minx - 2.0), 0, A.width));
    x1 = min(x1, clampToInt(ceil(maxx + 2.0), -1, A.width - 1));
    y0 = max(y0, clampToInt(floor(miny - 2.0), 0, A.height));
    y1 = min(y1, clampToInt(ceil(maxy + 2.0), -1, A.height - 1));
  }
  if (x1 < x0 || y1 < y0) { R.x1 = -1; R.x0 = 0; return; }
  R.x0 = x0; R.x1 = x1; R.y0 = y0; R.y1 = y1;
  const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
  if (bw > kTile || bh > kTile) {
    const uint32_t tilesX = (bw + kTile - 1) / kTile, tilesY = (bh + kTile - 1) / kTile;
    const unsigned long long old = atomicAdd(A.counter, (1ull << 32) | (unsigned long long)(tilesX * tilesY));
    R.big = 1;
    A.big[old >> 32] = BigRecord{t, (uint32_t)(old & 0xffffffffull), tilesX, tilesX * tilesY, x0, y0, x1, y1};
  }
}

// one warp per triangle whose box is at most 32x32 pixels
__global__ void __launch_bounds__(256) rasterSmallKernel(const __grid_constant__ RasterKernelArgs A) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= A.nTriangles) return;
  const TriRecord &R = A.tris[t];
  const int x0 = R.x0, x1 = R.x1, y0 = R.y0, y1 = R.y1;
  if (x1 < x0 || R.big) return;
  const int lane = threadIdx.x & 31;
  const int bw = x1 - x0 + 1;
  const int cols = bw <= 4 ? 4 : bw <= 8 ? 8 : bw <= 16 ? 16 : 32, rowsPerIter = 32 / cols;
  double a[3], b[3], c[3], Z[3], W[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { a[i] = R.a[i]; b[i] = R.b[i]; c[i] = R.c[i]; Z[i] = R.Z[i]; W[i] = R.W[i]; }
  const int x = x0 + (lane & (cols - 1));
  for (int y = y0 + lane / cols; y <= y1; y += rowsPerIter) {
    double e[3];
    float depth;
    if (x <= x1 && shadePixel(a, b, c, Z, W, x, y, e, &depth)) writeVisibility(A.vis, A.width, x, y, depth, t);
  }
}

// persistent warps over the 32x32 tiles reserved by the big triangles
__global__ void __launch_bounds__(256) rasterBigKernel(const __grid_constant__ RasterKernelArgs A) {
  const unsigned long long ctr = *A.counter;
  const uint32_t nRecords = (uint32_t)(ctr >> 32), nTiles = (uint32_t)(ctr & 0xffffffffull);
  const int lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nWarps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t g = warp; g < nTiles; g += nWarps) {
    // 32-ary search for the last record with firstTile <= g (records are sorted by firstTile)
    uint32_t lo = 0, hi = nRecords;
    while (hi - lo > 1) {
      const uint32_t step = (hi - lo + 31) / 32;
      const uint32_t probe = lo + lane * step;
      const bool le = probe < hi && A.big[probe].firstTile <= g;
      const int cnt = __popc(__ballot_sync(0xffffffffu, le)); // >= 1: lane 0 probes lo
      lo = lo + (cnt - 1) * step;
      hi = min(lo + step, hi);
    }
    const BigRecord br = A.big[lo];
    const TriRecord &R = A.tris[br.tri];
    const uint32_t local = g - br.firstTile;
    const int tx0 = R.x0 + (int)(local % br.tilesX) * kTile, ty0 = R.y0 + (int)(local / br.tilesX) * kTile;
    const int tx1 = min(tx0 + kTile - 1, R.x1), ty1 = min(ty0 + kTile - 1, R.y1);
    double a[3], b[3], c[3], Z[3], W[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { a[i] = R.a[i]; b[i] = R.b[i]; c[i] = R.c[i]; Z[i] = R.Z[i]; W[i] = R.W[i]; }
    // tile rejection: an edge function that is negative at the corner where it is largest is negative on the whole tile
    // (evaluated with a margin of one pixel so that rounding cannot reject a covered pixel)
    bool reject = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double cx = a[i] > 0.0 ? (double)(tx1 + 2) : (double)(tx0 - 1), cy = b[i] > 0.0 ? (double)(ty1 + 2) : (double)(ty0 - 1);
      reject = reject || ((a[i] * cx + b[i] * cy) + c[i] < 0.0);
    }
    if (reject) continue;
    const int x = tx0 + lane;
    if (x > tx1) continue;
    for (int y = ty0; y <= ty1; y++) {
      double e[3];
      float depth;
      if (shadePixel(a, b, c, Z, W, x, y, e, &depth)) writeVisibility(A.vis, A.width, x, y, depth, br.tri);
    }
  }
}


// ---- tile path --------------------------------------------------------------------------------------------------------------
constexpr int kScreenTileW = 32, kScreenTileH = 16, kTileLoopTriangles = 8192;

__device__ __forceinline__ void resolveFragment(const RasterKernelArgs &A, unsigned long long key, int x, int y, lgcu_fragment *fragments, uint64_t pitch) {
  float4 lo = make_float4(0.0f, 0.0f, 0.0f, 0.0f), hi = make_float4(0.0f, 0.0f, __uint_as_float(LGCU_NO_OBJECT), 1.0f);
  if (key != kEmpty) {
    const TriRecord &R = A.tris[(uint32_t)key];
    const double px = (double)x + 0.5, py = (double)y + 0.5;
    double e[3];
#pragma unroll
    for (int i = 0; i < 3; i++) e[i] = (R.a[i] * px + R.b[i] * py) + R.c[i];
    const double s = (e[0] + e[1]) + e[2];
    const float l0 = (float)(e[0] / s), l1 = (float)(e[1] / s), l2 = (float)(e[2] / s);
    float p[3], n[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      p[c] = (l0 * R.wp[0][c] + l1 * R.wp[1][c]) + l2 * R.wp[2][c];
      n[c] = (l0 * R.wn[0][c] + l1 * R.wn[1][c]) + l2 * R.wn[2][c];
    }
    lo = make_float4(p[0], p[1], p[2], n[0]);
    hi = make_float4(n[1], n[2], __uint_as_float(R.objectId), __uint_as_float((uint32_t)(key >> 32)));
  }
  float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(fragments) + (size_t)y * pitch + (size_t)x * sizeof(lgcu_fragment));
  dst[0] = lo;
  dst[1] = hi;
}

// One warp per 32x16 screen tile; lane = column. kFragments: G-buffer target (fragment buffer) / depth-only target (shadow map).
// 8 resident CTAs per SM (64 registers): the kernel is bound by latency, not by the fp64 pipe (ncu r03c at 126 registers / 4 CTAs: fp64
// pipe 27 %, issue 34 %; 4K G-buffer pass 0.261 ms at 4 CTAs, 0.237 at 5, 0.228 at 6, 0.214 at 8 and at 10). Where the latency was
// (r03c, per-instruction stall samples): 27 % of all samples waited for the 16 visibility-buffer loads of the resolve, one after the other
// behind each row's arithmetic; 16 % were instruction-cache misses of the 16x unrolled resolve (2 900 of the kernel's 4 200
// instructions); 16 % the cull loop's dependent scattered loads (big list -> record box -> record edges). Hence:
//   * the column of keys starts from the visibility buffer (what the one-warp-per-triangle path left there) instead of merging it at the
//     end: the 16 loads are in flight during the cull loop;
//   * the cull tests the box from the 32-byte big record (coalesced) and touches the triangle record only for the corner test;
//   * after the last triangle the keys are parked in shared memory and the resolve runs as a rolled loop (kResolveUnroll rows at a time).
template <bool kFragments, int kResolveUnroll>
__global__ void __launch_bounds__(128, 8) rasterTileKernel(const __grid_constant__ RasterKernelArgs A, lgcu_fragment *fragments, uint64_t pitch, LevelView depth) {
  __shared__ unsigned long long sBest[4][kScreenTileH][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tilesX = (A.width + kScreenTileW - 1) / kScreenTileW;
  const int tile = blockIdx.x * (blockDim.x >> 5) + warp;
  const int tilesY = (A.rows.y1 - A.rows.y0 + kScreenTileH - 1) / kScreenTileH;
  if (tile >= tilesX * tilesY) return;
  const int tx0 = (tile % tilesX) * kScreenTileW, ty0 = A.rows.y0 + (tile / tilesX) * kScreenTileH;
  const int tx1 = min(tx0 + kScreenTileW - 1, A.width - 1), ty1 = min(ty0 + kScreenTileH - 1, A.rows.y1 - 1);
  const int x = tx0 + lane;
  unsigned long long best[kScreenTileH];
#pragma unroll
  for (int r = 0; r < kScreenTileH; r++) best[r] = (x <= tx1 && ty0 + r <= ty1) ? A.vis[(size_t)(ty0 + r) * A.width + x] : kEmpty;
  const uint32_t nRecords = (uint32_t)(*A.counter >> 32);
  for (uint32_t base = 0; base < nRecords; base += 32) {
    // cull 32 big triangles against the tile: box overlap, then the corner test of rasterBigKernel
    bool keep = false;
    const uint32_t rec = base + lane;
    uint32_t myTri = 0;
    if (rec < nRecords) {
      const uint4 id = *reinterpret_cast<const uint4 *>(&A.big[rec]);
      const int4 box = *reinterpret_cast<const int4 *>(&A.big[rec].x0);
      myTri = id.x;
      keep = box.x <= tx1 && box.z >= tx0 && box.y <= ty1 && box.w >= ty0;
      if (keep) {
        const TriRecord &R = A.tris[myTri];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double ai = R.a[i], bi = R.b[i];
          const double cx = ai > 0.0 ? (double)(tx1 + 2) : (double)(tx0 - 1), cy = bi > 0.0 ? (double)(ty1 + 2) : (double)(ty0 - 1);
          keep = keep && !((ai * cx + bi * cy) + R.c[i] < 0.0);
        }
      }
    }
    uint32_t mask = __ballot_sync(0xffffffffu, keep);
    while (mask) {
      const int bit = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint32_t tri = __shfl_sync(0xffffffffu, myTri, bit);
      const TriRecord &R = A.tris[tri];
      double a[3], b[3], c[3], Z[3], W[3];
#pragma unroll
      for (int i = 0; i < 3; i++) { a[i] = R.a[i]; b[i] = R.b[i]; c[i] = R.c[i]; Z[i] = R.Z[i]; W[i] = R.W[i]; }
      const bool xin = x >= R.x0 && x <= R.x1 && x <= tx1;
      const int ry0 = R.y0, ry1 = R.y1;
#pragma unroll
      for (int r = 0; r < kScreenTileH; r++) {
        const int y = ty0 + r;
        const unsigned long long key = shadePixelKey(a, b, c, Z, W, x, y, tri, xin && y <= ty1 && y >= ry0 && y <= ry1);
        best[r] = key < best[r] ? key : best[r];
      }
    }
  }
  if (x > tx1) return;
#pragma unroll
  for (int r = 0; r < kScreenTileH; r++) sBest[warp][r][lane] = best[r]; // read back by the same lane only: no barrier
  const int rowsHere = ty1 - ty0 + 1;
#pragma unroll(kResolveUnroll)
  for (int r = 0; r < rowsHere; r++) {
    const int y = ty0 + r;
    const unsigned long long key = sBest[warp][r][lane];
    if (kFragments)
      resolveFragment(A, key, x, y, fragments, pitch);
    else
      reinterpret_cast<float *>(depth.ptr + (size_t)y * depth.pitch)[x] = key == kEmpty ? 1.0f : __uint_as_float((uint32_t)(key >> 32));
  }
}

__global__ void __launch_bounds__(256) rasterResolveFragmentsKernel(const __grid_constant__ RasterKernelArgs A, lgcu_fragment *fragments, uint64_t pitch) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = A.rows.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= A.width || y >= A.rows.y1) return;
  const unsigned long long key = A.vis[(size_t)y * A.width + x];
  float4 lo = make_float4(0.0f, 0.0f, 0.0f, 0.0f), hi = make_float4(0.0f, 0.0f, __uint_as_float(LGCU_NO_OBJECT), 1.0f);
  if (key != kEmpty) {
    const TriRecord &R = A.tris[(uint32_t)key];
    const double px = (double)x + 0.5, py = (double)y + 0.5;
    double e[3];
#pragma unroll
    for (int i = 0; i < 3; i++) e[i] = (R.a[i] * px + R.b[i] * py) + R.c[i];
    const double s = (e[0] + e[1]) + e[2];
    const float l0 = (float)(e[0] / s), l1 = (float)(e[1] / s), l2 = (float)(e[2] / s);
    float p[3], n[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      p[c] = (l0 * R.wp[0][c] + l1 * R.wp[1][c]) + l2 * R.wp[2][c];
      n[c] = (l0 * R.wn[0][c] + l1 * R.wn[1][c]) + l2 * R.wn[2][c];
    }
    lo = make_float4(p[0], p[1], p[2], n[0]);
    hi = make_float4(n[1], n[2], __uint_as_float(R.objectId), __uint_as_float((uint32_t)(key >> 32)));
  }
  float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(fragments) + (size_t)y * pitch + (size_t)x * sizeof(lgcu_fragment));
  dst[0] = lo;
  dst[1] = hi;
}

__global__ void __launch_bounds__(256) rasterResolveDepthKernel(const __grid_constant__ RasterKernelArgs A, LevelView depth) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= A.width || y >= A.height) return;
  const unsigned long long key = A.vis[(size_t)y * A.width + x];
  reinterpret_cast<float *>(depth.ptr + (size_t)y * depth.pitch)[x] = key == kEmpty ? 1.0f : __uint_as_float((uint32_t)(key >> 32));
}

uint64_t align256(uint64_t v) { return (v + 255) & ~uint64_t(255); }

} // namespace

uint64_t rasterScratchBytes(uint32_t nTriangles, uint32_t width, uint32_t height) {
  const uint64_t n = nTriangles ? nTriangles : 1;
  return 256 + align256(n * sizeof(TriRecord)) + align256(n * sizeof(BigRecord)) + align256(uint64_t(width) * height * 8);
}

cudaError_t launchRaster(const RasterArgs &r, int smCount, cudaStream_t s) {
  RasterKernelArgs A;
  A.vertices = r.scene.vertices;
  A.indices = r.scene.indices;
  A.draws = r.scene.draws;
  A.objects = r.scene.objects;
  A.nVertices = r.scene.nVertices;
  A.nIndices = r.scene.nIndices;
  A.nDraws = r.scene.nDraws;
  A.nObjects = r.scene.nObjects;
  A.nTriangles = r.scene.nTriangles;
  A.viewProj = r.viewProj;
  A.width = r.width;
  A.height = r.height;
  A.rows = r.rows;
  unsigned char *base = static_cast<unsigned char *>(r.scratch);
  const uint64_t n = A.nTriangles ? A.nTriangles : 1;
  A.counter = reinterpret_cast<unsigned long long *>(base);
  A.tris = reinterpret_cast<TriRecord *>(base + 256);
  A.big = reinterpret_cast<BigRecord *>(base + 256 + align256(n * sizeof(TriRecord)));
  A.vis = reinterpret_cast<unsigned long long *>(base + 256 + align256(n * sizeof(TriRecord)) + align256(n * sizeof(BigRecord)));
  cudaError_t e;
  if ((e = cudaMemsetAsync(A.counter, 0, 256, s)) != cudaSuccess) return e;
  const int rowsN = r.rows.y1 - r.rows.y0;
  if (rowsN <= 0) return cudaSuccess;
  if ((e = cudaMemsetAsync(A.vis + (size_t)r.rows.y0 * r.width, 0xFF, (size_t)rowsN * r.width * 8, s)) != cudaSuccess) return e;
  static const int forcePath = getenv("LGCU_RASTER_PATH") ? atoi(getenv("LGCU_RASTER_PATH")) : 0; // development switch: 1 = big + resolve, 2 = tile
  const bool tilePath = forcePath == 2 || (forcePath == 0 && A.nTriangles <= (uint32_t)kTileLoopTriangles);
  if (A.nTriangles && A.nDraws) {
    rasterSetupKernel<<<(A.nTriangles + 127) / 128, 128, 0, s>>>(A);
    rasterSmallKernel<<<(A.nTriangles + 7) / 8, 256, 0, s>>>(A);
    if (!tilePath) rasterBigKernel<<<smCount * 8, 256, 0, s>>>(A);
  }
  if (tilePath) { // every big triangle is tested against every screen tile: only for scenes where that loop is short
    const int tiles = ((r.width + kScreenTileW - 1) / kScreenTileW) * ((rowsN + kScreenTileH - 1) / kScreenTileH);
    static const int unroll = getenv("LGCU_RASTER_RESOLVE_UNROLL") ? atoi(getenv("LGCU_RASTER_RESOLVE_UNROLL")) : 2; // A/B switch
    if (!r.fragments)
      rasterTileKernel<false, 4><<<(tiles + 3) / 4, 128, 0, s>>>(A, nullptr, 0, r.depth);
    else if (unroll == 1)
      rasterTileKernel<true, 1><<<(tiles + 3) / 4, 128, 0, s>>>(A, r.fragments, r.fragmentPitch, r.depth);
    else if (unroll == 4)
      rasterTileKernel<true, 4><<<(tiles + 3) / 4, 128, 0, s>>>(A, r.fragments, r.fragmentPitch, r.depth);
    else if (unroll == 16)
      rasterTileKernel<true, 16><<<(tiles + 3) / 4, 128, 0, s>>>(A, r.fragments, r.fragmentPitch, r.depth);
    else
      rasterTileKernel<true, 2><<<(tiles + 3) / 4, 128, 0, s>>>(A, r.fragments, r.fragmentPitch, r.depth);
    return cudaGetLastError();
  }
  const dim3 grid((r.width + 31) / 32, (rowsN + 7) / 8);
  if (r.fragments)
    rasterResolveFragmentsKernel<<<grid, 256, 0, s>>>(A, r.fragments, r.fragmentPitch);
  else
    rasterResolveDepthKernel<<<grid, 256, 0, s>>>(A, r.depth);
  return cudaGetLastError();
}

} // namespace lgcu
