// k_gather_fast.cu — screen-space GI gather (K5), throughput variant for sm_100a.
//
// Same function as SH/SSVGI/indirectLighting.frag:114-272 (k_gather_strict.cu is the line-by-line form). The pass is bound by
// instruction issue, not by HBM (≈25 trilinear pyramid samples and ≈8 trilinear light fetches per pixel; ncu r01j: issue slots 75 %
// busy, DRAM 12 %), so the design removes instructions, not bytes:
//
//  * Pattern-coherent CTAs. The shader's 4x4 interleaved pattern gives every pixel with the same (x&3, y&3) the same march
//    directions, step offsets and LODs. A CTA owns a 64x64 (or 64x32) pixel tile and walks pattern classes in a CTA-uniform
//    loop; in each pass thread t shades one pixel of the class. Direction, step offset, LOD and mip geometry are therefore
//    uniform-datapath operands, every level branch is uniform, and neighbouring lanes read neighbouring texels from LOD 2 upwards.
//    The 16 classes of a tile are split over 4 CTAs that sit next to each other in the grid (finer work units for the tail of
//    the grid, shared L2 footprint: measured DRAM traffic of the launch = 1.18x its algorithmic bytes).
//  * No transcendental per sample: pow/log (step offset, LOD, iteration count) come from host tables shared with the strict
//    kernel (bit-identical level selection). The per-sample unprojection collapses to FMAs: the ray through pixel s is
//    R(s) = Ra*sx + Rb*sy + Rc (affine in pixel coordinates), so along a march direction every dot product of the horizon test
//    is affine in the step offset with per-direction constants.
//  * The step rows of a pass's pattern class are staged in shared memory (112 bytes per step): five 16-byte broadcast loads per march
//    step replace ~20 uniform-datapath loads, address computations and UR->R moves (ncu r02e: 31 % of the executed instructions were
//    such overhead). Marching two or four directions of a pixel in lock-step was measured and dropped: it needs 80 / 128 registers,
//    and this kernel lives on occupancy (4 -> 3 -> 2 resident CTAs per SM: 1.49 -> 1.69 -> 1.93 ms; lock-step 2 / 4: 1.81 / 2.25 ms).
//  * Private side pyramid (lgcu_gi_gather_pack, built once per frame from the two blurred pyramids):
//      - depth, levels >= 1: per level (w+1)x(h+1) float4 = the four clamp-to-edge .r taps of the bilinear footprint whose top-left
//        texel is (qx-1, qy-1) -> ONE 16-byte load per footprint, no clamps, one address;
//      - light, levels >= 2: the same footprint as 12 floats (4 taps x RGB, already converted to fp32) -> three 16-byte loads, no
//        clamps, no fp16->fp32 converts (24 per trilinear fetch in the plain layout, 15 % of the stall samples in r01b). Levels >= 2
//        hold 1/12 of the pyramid's texels, so this costs 33 MB at 4K; levels 0 and 1 (touched by the first two march steps only)
//        are read from the images themselves (level 0 of the depth pyramid as well: packing it was 3/4 of the pack pass).
//    Without a scratch buffer (lgcu_gi_gather) the same kernel fetches every tap from the images.
//  * Hits are processed inline (they are spatially coherent). A sample is rejected with a half-plane + cross-product test instead of
//    atan; on a hit atan2 is an 8-term minimax polynomial (1.2e-7 rad) and sin(2h)/cos(2h) of ComputeHorizonContribution come
//    algebraically from the horizon vector: cos2h = (x²-y²)/(x²+y²), sin2h = 2xy/(x²+y²). The light fetch reuses the footprint
//    (integer taps + weights) of the depth fetch.
//  * What is numerically delicate is kept in the shader's order: the centre position is reconstructed from the D32 depth exactly
//    as the shader does (its fp32 cancellation noise is part of the reference result and is amplified by 1/sample distance), so
//    the two variants see the same centre. Everything after that is evaluated in cancellation-free forms: against the binary64
//    evaluation of the shader's formula (oracle/gather_noise_probe.c) this kernel has < 1e-4 outliers at 4K and 8K, where the fp32
//    shader itself has 2e-3 / 2e-2 (tests/helpers.py: OUTLIER_BAR).
#include <cmath>
#include <cstdlib>

#include "lgcu_kernels.h"

#ifndef LGCU_DEPTH_QUAD_LEVEL0
#define LGCU_DEPTH_QUAD_LEVEL0 0
#endif

namespace lgcu {

namespace {

constexpr int kTile = 64;      // pixels per tile edge
constexpr int kThreads = 256;  // 16x16 pixels of one pattern class per pass
constexpr int kMaxSteps = 12;  // march steps the tables hold (8 are reached for landscape viewports at any resolution)
constexpr int kDepthQuadLevel0 = LGCU_DEPTH_QUAD_LEVEL0, kLightQuadLevel0 = 2; // first levels of the side pyramid's depth / light quads
constexpr uint32_t F16 = LGCU_FORMAT_R16G16B16A16_SFLOAT, D32 = LGCU_FORMAT_D32_SFLOAT;
constexpr float kFloorMagic = 12582912.0f; // 1.5 * 2^23: x + magic (rounded down) has floor(x) in its low mantissa bits
constexpr int kFloorMagicBits = 0x4B400000;

// Per pyramid level, three 16-byte groups (the kernel reads them as float4 / int4 from shared memory)
struct __align__(16) LevelGeom {
  float scaleX, scaleY; // w_l / viewport.x, h_l / viewport.y
  float maxX, maxY;     // w_l - 1, h_l - 1
  int quadOfs, quadPitch; // depth quads: level origin and row pitch in float4
  int lightOfs;           // light quads: level origin in float4, three per entry, same pitch (levels >= kLightQuadLevel0)
  int mode;               // bit 0: depth quads present, bit 1: light quads present
  int texOfs, texPitch;   // level origin and row pitch in 8-byte texels (both pyramids share the layout)
  int wm1, hm1;
};
struct __align__(16) StepRow { // per (pattern, step): everything the march needs for one sample, in one 112-byte row = 7 x 16 bytes
  float off, frac;
  int l0, l1;
  LevelGeom g0, g1;
};
constexpr int kRowVecs = sizeof(StepRow) / 16;
static_assert(sizeof(StepRow) == 112 && kRowVecs == 7, "StepRow layout");
struct __align__(16) DirEntry { // per (pattern, direction)
  float dirX, dirY, invDirX, invDirY;
  float rdX, rdY, rdZ, q2; // Rd = Ra*dirX + Rb*dirY, q2 = |Rd|²
};
struct FastTables {
  StepRow row[16 * kMaxSteps];
  DirEntry dir[16][kGatherDirs];
  float iterThreshold[kMaxSteps];
  float ra[3], rb[3], rc[3]; // R(p) = ra*px + rb*py + rc  ∝ (far-plane point through pixel p) - cam
  float raySign;             // sign of the homogeneous w of that point: rayDir = raySign * R / |R|
  int maxSteps;
};

__device__ __forceinline__ float fastRcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fastRsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lerpf(float p, float q, float t) { return fmaf(q - p, t, p); }
__device__ __forceinline__ float dotf(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

// atan2 for finite, not-both-zero arguments: 8-term odd minimax polynomial on [0,1] (max error 1.2e-7 rad in fp32)
__device__ __forceinline__ float atan2Poly(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mn = fminf(ax, ay), mxv = fmaxf(ax, ay);
  const float t = mn * fastRcp(fmaxf(mxv, 1e-37f));
  const float s = t * t;
  float p = -0.004054398275911808f;
  p = fmaf(p, s, 0.021862303838133812f);
  p = fmaf(p, s, -0.055911313742399216f);
  p = fmaf(p, s, 0.09642116725444794f);
  p = fmaf(p, s, -0.13908594846725464f);
  p = fmaf(p, s, 0.1994655728340149f);
  p = fmaf(p, s, -0.33329859375953674f);
  p = fmaf(p, s, 0.9999993443489075f);
  float r = p * t;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.0f) r = 3.14159265358979324f - r;
  return copysignf(r, y);
}

// Bilinear footprint at one level: integer top-left tap (>= -1) and the two weights. The coordinate is clamped as a float
// first ([-1, size-1]); that selects the same taps and the same value as clamp-to-edge of the integer coordinates.
struct Footprint {
  int ix, iy;
  float a, b;
};
__device__ __forceinline__ Footprint footprint(const float4 g /* scaleX, scaleY, maxX, maxY */, float sx, float sy) {
  Footprint f;
  const float u = fminf(fmaxf(fmaf(sx, g.x, -0.5f), -1.0f), g.z), v = fminf(fmaxf(fmaf(sy, g.y, -0.5f), -1.0f), g.w);
  const float tu = __fadd_rd(u, kFloorMagic), tv = __fadd_rd(v, kFloorMagic);
  f.a = u - (tu - kFloorMagic);
  f.b = v - (tv - kFloorMagic);
  f.ix = __float_as_int(tu) - kFloorMagicBits;
  f.iy = __float_as_int(tv) - kFloorMagicBits;
  return f;
}

struct Pyramids {
  const float2 *__restrict__ moments; // blurredDepthMoments, level 0 base
  const uint2 *__restrict__ light;    // blurredDirectLight, level 0 base
  const float4 *__restrict__ side;    // side pyramid (depth quads, then light quads) or nullptr
};

// q = {quadOfs, quadPitch, lightOfs, mode}, tex = {texOfs, texPitch, wm1, hm1} of the level (the latter only read without quads)
template <bool kSide> __device__ __forceinline__ float fetchDepth(const int4 q, const uint4 *texRow, const Footprint &f, const Pyramids &p) {
  float t00, t10, t01, t11;
  if (kSide && (kDepthQuadLevel0 == 0 || (q.w & 1))) { // the same for every thread of the CTA
    const float4 v = __ldg(p.side + (unsigned)(q.x + (f.iy + 1) * q.y + (f.ix + 1)));
    t00 = v.x, t10 = v.y, t01 = v.z, t11 = v.w;
  } else {
    const uint4 tex = *texRow;
    const int x0 = max(f.ix, 0), x1 = min(f.ix + 1, (int)tex.z), y0 = max(f.iy, 0), y1 = min(f.iy + 1, (int)tex.w);
    const unsigned r0 = tex.x + y0 * tex.y, r1 = tex.x + y1 * tex.y;
    t00 = __ldg(&p.moments[r0 + x0].x), t10 = __ldg(&p.moments[r0 + x1].x), t01 = __ldg(&p.moments[r1 + x0].x), t11 = __ldg(&p.moments[r1 + x1].x);
  }
  return lerpf(lerpf(t00, t10, f.a), lerpf(t01, t11, f.a), f.b);
}

__device__ __forceinline__ float3 fetchLight(const int4 q, const uint4 *texRow, const Footprint &f, const Pyramids &p) {
  float3 r;
  if (q.w & 2) { // same for every thread of the CTA: 12 fp32 values {t00.rgb, t10.rgb, t01.rgb, t11.rgb} in three 16-byte loads
    const float4 *e = p.side + (unsigned)(q.z + 3 * ((f.iy + 1) * q.y + (f.ix + 1)));
    const float4 A = __ldg(e), B = __ldg(e + 1), C = __ldg(e + 2);
    r.x = lerpf(lerpf(A.x, A.w, f.a), lerpf(B.z, C.y, f.a), f.b);
    r.y = lerpf(lerpf(A.y, B.x, f.a), lerpf(B.w, C.z, f.a), f.b);
    r.z = lerpf(lerpf(A.z, B.y, f.a), lerpf(C.x, C.w, f.a), f.b);
    return r;
  }
  const uint4 tex = *texRow;
  const int x0 = max(f.ix, 0), x1 = min(f.ix + 1, (int)tex.z), y0 = max(f.iy, 0), y1 = min(f.iy + 1, (int)tex.w);
  const unsigned o0 = tex.x + y0 * tex.y, o1 = tex.x + y1 * tex.y;
  const uint2 r00 = __ldg(&p.light[o0 + x0]), r10 = __ldg(&p.light[o0 + x1]), r01 = __ldg(&p.light[o1 + x0]), r11 = __ldg(&p.light[o1 + x1]);
  const float2 a00 = __half22float2(*reinterpret_cast<const __half2 *>(&r00.x)), a10 = __half22float2(*reinterpret_cast<const __half2 *>(&r10.x));
  const float2 a01 = __half22float2(*reinterpret_cast<const __half2 *>(&r01.x)), a11 = __half22float2(*reinterpret_cast<const __half2 *>(&r11.x));
  const float b00 = __low2float(*reinterpret_cast<const __half2 *>(&r00.y)), b10 = __low2float(*reinterpret_cast<const __half2 *>(&r10.y));
  const float b01 = __low2float(*reinterpret_cast<const __half2 *>(&r01.y)), b11 = __low2float(*reinterpret_cast<const __half2 *>(&r11.y));
  r.x = lerpf(lerpf(a00.x, a10.x, f.a), lerpf(a01.x, a11.x, f.a), f.b);
  r.y = lerpf(lerpf(a00.y, a10.y, f.a), lerpf(a01.y, a11.y, f.a), f.b);
  r.z = lerpf(lerpf(b00, b10, f.a), lerpf(b01, b11, f.a), f.b);
  return r;
}

// exact-order pieces shared with the strict kernel (no FMA contraction, IEEE divide / sqrt)
__device__ __forceinline__ float4 mulMat4Exact(const Mat4 &M, float x, float y, float z, float w) {
  float4 r;
  r.x = __fadd_rn(__fadd_rn(__fmul_rn(M.m[0], x), __fmul_rn(M.m[4], y)), __fadd_rn(__fmul_rn(M.m[8], z), __fmul_rn(M.m[12], w)));
  r.y = __fadd_rn(__fadd_rn(__fmul_rn(M.m[1], x), __fmul_rn(M.m[5], y)), __fadd_rn(__fmul_rn(M.m[9], z), __fmul_rn(M.m[13], w)));
  r.z = __fadd_rn(__fadd_rn(__fmul_rn(M.m[2], x), __fmul_rn(M.m[6], y)), __fadd_rn(__fmul_rn(M.m[10], z), __fmul_rn(M.m[14], w)));
  r.w = __fadd_rn(__fadd_rn(__fmul_rn(M.m[3], x), __fmul_rn(M.m[7], y)), __fadd_rn(__fmul_rn(M.m[11], z), __fmul_rn(M.m[15], w)));
  return r;
}
__device__ __forceinline__ float dot3Exact(V3 a, V3 b) { return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)); }

// kT = threads per CTA: a CTA shades a 64 x (kT / 4) pixel tile (256 -> 64x64, 128 -> 64x32). The smaller tile doubles the CTA count
// for row strips and small frames, where the grid would otherwise be a couple of waves with a long tail (multi-GPU strips).
// kSide: the side pyramid exists (depth quads on every level, light quads from level kLightQuadLevel0 up).
// The step rows of the pass's pattern class are staged in shared memory once per pass (7 x 16 bytes per step): the march loop reads
// them with five 16-byte broadcast loads instead of ~20 uniform-datapath loads, address computations and UR->R moves.
template <bool kSide, int kMinBlocks, int kT, bool kCompactWarp>
__global__ void __launch_bounds__(kT, kMinBlocks) gatherFastKernel(const __grid_constant__ GatherArgs a, const __grid_constant__ FastTables tb,
                                                                          const float4 *__restrict__ side, int xSlices) {
  __shared__ uint4 sRow[kMaxSteps * kRowVecs];
  __shared__ float4 sDir[kGatherDirs * 2];
  const int t = threadIdx.x;
  // tiles start on a multiple of 4 rows so that the pass number IS the pattern index (x&3) + 4*(y&3)   (:155, :161)
  // The 16 pattern classes of a tile are split over xSlices CTAs (0 or 1 = one CTA does all 16) that are neighbours in blockIdx.x
  const int nSlices = xSlices > 0 ? xSlices : 1, slice = (int)blockIdx.x % nSlices;
  const int tileX = ((int)blockIdx.x / nSlices) * kTile, tileY = (a.rows.y0 & ~3) + blockIdx.y * (kT / 4);
  const float vpx = a.viewport[0], vpy = a.viewport[1];
  const float invVpx10 = 10.0f / vpx, invVpy10 = 10.0f / vpy;
  const V3 cam = v3(a.cam[0], a.cam[1], a.cam[2]);
  const V3 Ra = v3(tb.ra[0], tb.ra[1], tb.ra[2]), Rb = v3(tb.rb[0], tb.rb[1], tb.rb[2]), Rc = v3(tb.rc[0], tb.rc[1], tb.rc[2]);
  Pyramids pyr;
  pyr.moments = reinterpret_cast<const float2 *>(a.moments.lv[0].ptr);
  pyr.light = reinterpret_cast<const uint2 *>(a.light.lv[0].ptr);
  pyr.side = side;
  // lanes of a warp: 16 x 2 pixels of the class (64 x 8 px) or, compact, 8 x 4 (32 x 16 px)
  const int tx = kCompactWarp ? 4 * ((((t >> 5) & 1) << 3) + (t & 7)) : 4 * (t & 15);
  const int ty = kCompactWarp ? 4 * (((t >> 6) << 2) + ((t >> 3) & 3)) : 4 * (t >> 4);
  const int maxSteps = tb.maxSteps;

  const int idxPerCta = 16 / nSlices, idx0 = slice * idxPerCta;
#pragma unroll 1
  for (int idx = idx0; idx < idx0 + idxPerCta; idx++) { // one pattern class per pass (CTA-uniform)
    __syncthreads(); // the previous pass is done with the staged rows
    for (int v = t; v < maxSteps * kRowVecs; v += kT) sRow[v] = reinterpret_cast<const uint4 *>(&tb.row[idx * kMaxSteps])[v];
    if (t < kGatherDirs * 2) sDir[t] = reinterpret_cast<const float4 *>(&tb.dir[idx][0])[t];
    __syncthreads();
    const int x = tileX + tx + (idx & 3), y = tileY + ty + (idx >> 2);
    const bool active = x < a.indirect.w && y >= a.rows.y0 && y < a.rows.y1;
    if (!__any_sync(0xffffffffu, active)) continue;
    const int cx = active ? x : 0, cy = active ? y : a.rows.y0; // inactive lanes shade a valid pixel and discard it
    const float px = (float)cx + 0.5f, py = (float)cy + 0.5f;
    // --- centre reconstruction in the shader's order (:116-132, :182) -------------------------------------------------
    const float cu = __fdiv_rn(px, vpx), cv = __fdiv_rn(py, vpy);
    const float4 ns = Texel<F16>::load(a.normal, cx, cy);
    const float zc = Texel<D32>::load(a.depthStencil, cx, cy).x;
    const float4 vc = mulMat4Exact(a.invViewProj, __fadd_rn(__fmul_rn(cu, 2.0f), -1.0f), __fadd_rn(__fmul_rn(cv, 2.0f), -1.0f), zc, 1.0f);
    const V3 C = v3(__fdiv_rn(vc.x, vc.w), __fdiv_rn(vc.y, vc.w), __fdiv_rn(vc.z, vc.w));
    const V3 N = v3(ns.x, ns.y, ns.z);
    const V3 E = v3(__fadd_rn(cam.x, -C.x), __fadd_rn(cam.y, -C.y), __fadd_rn(cam.z, -C.z)); // cam - C
    const float invLenE = __fdiv_rn(1.0f, __fsqrt_rn(dot3Exact(E, E)));
    const V3 eye = v3(__fmul_rn(E.x, invLenE), __fmul_rn(E.y, invLenE), __fmul_rn(E.z, invLenE));
    // --- ray through the pixel: R0 ∝ far-plane point - cam --------------------------------------------------------------
    const V3 R0 = v3(fmaf(Ra.x, px, fmaf(Rb.x, py, Rc.x)), fmaf(Ra.y, px, fmaf(Rb.y, py, Rc.y)), fmaf(Ra.z, px, fmaf(Rb.z, py, Rc.z)));
    const float q0 = dotf(R0, R0);
    const float n0 = q0 * fastRsqrt(q0);
    const float eN = dot3Exact(eye, N), eN4 = 0.25f * eN;
    const float xE = dotf(eye, E), xR0 = tb.raySign * dotf(eye, R0);
    // out = sum over directions of (2 L_d) / 4 (:268) with L_d = sum(Ls*c) - 0.01*sum(c) + 0.01*hc0 (:209, :261-262): the light terms go
    // straight into the pixel's sums (weight 0.5 c), the sum(c) and hc0 terms into one scalar for all four directions
    float sumX = 0.0f, sumY = 0.0f, sumZ = 0.0f, cSum = 0.0f;

#pragma unroll 1
    for (int d = 0; d < kGatherDirs; d++) {
      const float4 deA = sDir[2 * d], deB = sDir[2 * d + 1]; // {dirX, dirY, invDirX, invDirY}, {rdX, rdY, rdZ, q2}
      const V3 Rd = v3(deB.x, deB.y, deB.z);
      // tangent = normalize(rayDir(p + dir) - rayDir(p)) in a cancellation-free form (:181-183)
      const float q2 = deB.w, q1 = 2.0f * dotf(R0, Rd);
      const float qs = q0 + q1 + q2, n1 = qs * fastRsqrt(qs);
      const float g = (q1 + q2) * fastRcp(n0 + n1);
      const V3 tanU = v3(fmaf(Rd.x, n0, -R0.x * g), fmaf(Rd.y, n0, -R0.y * g), fmaf(Rd.z, n0, -R0.z * g));
      const float tInv = tb.raySign * fastRsqrt(dotf(tanU, tanU));
      const V3 tang = v3(tanU.x * tInv, tanU.y * tInv, tanU.z * tInv);
      const float tN = dotf(tang, N), tN4 = 0.25f * tN;
      // initial horizon from the surface normal (:193-198): q = cross(-cross(tangent, eye), N) = tangent (eye.N) - eye (tangent.N)
      // for unit eye and tangent, so (q.eye, q.tangent) = (te eN - tN, eN - te tN) with te = tangent.eye
      const float te = dotf(tang, eye);
      float mx = fmaf(te, eN, -tN), my = fmaf(-te, tN, eN);
      float maxH = atan2Poly(my, mx);
      const float invM = fastRcp(fmaxf(fmaf(mx, mx, my * my), 1e-37f));
      float c2m = (mx * mx - my * my) * invM, s2m = 2.0f * mx * my * invM;
      // BoxRayCast + iteration count (:83-99, :202-212), exact like the strict kernel
      const float t1 = __fmul_rn(0.0f - px, deA.z), t2 = __fmul_rn(vpx - px, deA.z);
      const float t3 = __fmul_rn(0.0f - py, deA.w), t4 = __fmul_rn(vpy - py, deA.w);
      const float path = fabsf(glmMin(glmMax(t1, t2), glmMax(t3, t4)));
      int iterations = 0;
#pragma unroll
      for (int n = 0; n < 8; n++) iterations += (path >= tb.iterThreshold[n]) ? 1 : 0; // thresholds beyond maxSteps are +inf
      if (maxSteps > 8) { // uniform; not reached by landscape viewports
#pragma unroll 1
        for (int n = 8; n < maxSteps; n++) iterations += (path >= tb.iterThreshold[n]) ? 1 : 0;
      }
      if (!active) iterations = 0;
      // ambient term 0.01 * HC(0, maxH) (:209): cos(0) = 1, sin(0) = 0.  L = sum(Ls*c) - 0.01*sum(c) + 0.01*hc0
      cSum -= eN4 * (1.0f - c2m) + tN4 * (2.0f * maxH - s2m);
      // per-direction affine coefficients of the horizon vector
      const float yE = dotf(tang, E), yR0 = tb.raySign * dotf(tang, R0);
      const float xRd = tb.raySign * dotf(eye, Rd), yRd = tb.raySign * dotf(tang, Rd);

      const int warpIters = __reduce_max_sync(0xffffffffu, iterations);
      const uint4 *row = sRow;
#pragma unroll 1
      for (int k = 0; k < warpIters; k++, row += kRowVecs) { // :214
        const float4 hdr = *reinterpret_cast<const float4 *>(row); // off, frac, l0, l1
        const float off = hdr.x, frac = hdr.y;
        const float2 dir = *reinterpret_cast<const float2 *>(&sDir[2 * d]); // re-read per step: two registers less across the loop
        const float sx = fmaf(dir.x, off, px), sy = fmaf(dir.y, off, py); // :218-219 (in pixels)
        const Footprint f0 = footprint(*reinterpret_cast<const float4 *>(row + 1), sx, sy);
        float z = fetchDepth<kSide>(*reinterpret_cast<const int4 *>(row + 2), row + 3, f0, pyr);
        Footprint f1; // only set and only read when tri
        const bool tri = frac > 0.0f; // the same for every thread of the CTA
        if (tri) {
          f1 = footprint(*reinterpret_cast<const float4 *>(row + 4), sx, sy);
          z = fmaf(frac, fetchDepth<kSide>(*reinterpret_cast<const int4 *>(row + 5), row + 6, f1, pyr) - z, z); // :240
        }
        // horizon test + hit body (:241-263)
        const float zs = z * fastRsqrt(fmaf(off, fmaf(off, q2, q1), q0)); // z / |R(s)|
        const float hx = fmaf(zs, fmaf(off, xRd, xR0), xE); // dot(eye, P - C)      :241-252
        const float hy = fmaf(zs, fmaf(off, yRd, yR0), yE); // dot(tangent, P - C)
        // h < maxH for angles in (-pi, pi]: different half planes decide directly, otherwise the cross product does
        const bool lower = hy < 0.0f, lowerM = my < 0.0f;
        const bool hit = (lower != lowerM) ? lower : (hx * my - hy * mx > 0.0f);
        if (hit && k < iterations) { // :254
          const float h = atan2Poly(hy, hx);
          const float inv = fastRcp(fmaxf(fmaf(hx, hx, hy * hy), 1e-37f));
          const float c2 = (hx * hx - hy * hy) * inv, s2 = 2.0f * hx * hy * inv;
          const float hc = eN4 * (c2 - c2m) + tN4 * (((2.0f * maxH - 2.0f * h) - s2m) + s2); // :44-49
          // product of the four saturates of :220-228 (at most one per axis is below 1): sat(min(u, 1 - u) * 10)
          const float ex = fminf(sx, vpx - sx) * invVpx10, ey = fminf(sy, vpy - sy) * invVpy10;
          const float sideMult = saturatef(ex) * saturatef(ey);
          float3 ls = fetchLight(*reinterpret_cast<const int4 *>(row + 2), row + 3, f0, pyr); // :256
          if (tri) {
            const float3 hi = fetchLight(*reinterpret_cast<const int4 *>(row + 5), row + 6, f1, pyr);
            ls.x = fmaf(frac, hi.x - ls.x, ls.x);
            ls.y = fmaf(frac, hi.y - ls.y, ls.y);
            ls.z = fmaf(frac, hi.z - ls.z, ls.z);
          }
          const float c = hc * sideMult; // :258
          const float ch = 0.5f * c;
          sumX = fmaf(ls.x, ch, sumX);   // :261, :268
          sumY = fmaf(ls.y, ch, sumY);
          sumZ = fmaf(ls.z, ch, sumZ);
          cSum += c;                     // :262
          maxH = h;                      // :263
          c2m = c2;
          s2m = s2;
          mx = hx;
          my = hy;
        }
      }
    }
    const float amb = -0.005f * cSum;
    if (active) storeColor(a.outFormat, a.indirect, x, y, make_float4(sumX + amb, sumY + amb, sumZ + amb, 1.0f)); // :270
  }
}

// Side pyramid. Entry (qx, qy) of level l, qx in [0, w], qy in [0, h], describes the bilinear footprint whose taps are
// {(x0,y0), (x1,y0), (x0,y1), (x1,y1)} with x0 = clamp(qx-1), x1 = clamp(qx), y0 = clamp(qy-1), y1 = clamp(qy):
// depth quads hold the four .r taps of blurredDepthMoments, light quads the four RGB taps of blurredDirectLight as fp32.
constexpr int kMaxPackJobs = 2 * kMaxGatherLevels;
struct PackArgs {
  PyramidView moments, light;
  int jobLevel[kMaxPackJobs], jobLight[kMaxPackJobs];       // level, 0 = depth quads / 1 = light quads
  int jobOfs[kMaxPackJobs], jobPitch[kMaxPackJobs];         // destination origin (float4) and entries per row
  int jobRowBegin[kMaxPackJobs], jobRowEnd[kMaxPackJobs];   // entry rows to (re)build
  int jobBlockBegin[kMaxPackJobs + 1];                      // first CTA of each job (32x32 entries per CTA)
  int jobs;
};

// One CTA builds 32 x 32 entries: lane = column, every thread kPackRows vertically adjacent entries. Those share their rows — five
// rows x two columns = 10 loads for 4 entries instead of 16 — and all of a thread's loads are issued before its first store. With one
// entry per thread (r02x: 46 767 CTAs that each live for one DRAM round trip, 3.5 TB/s) the pass was bound by that latency, not by HBM.
constexpr int kPackRows = 4, kPackCtaRows = 8 * kPackRows;
__global__ void __launch_bounds__(256) packSidePyramidKernel(const __grid_constant__ PackArgs p, float4 *__restrict__ side) {
  int j = 0;
  while (j + 1 < p.jobs && (int)blockIdx.x >= p.jobBlockBegin[j + 1]) j++;
  const int l = p.jobLevel[j], pitch = p.jobPitch[j], rowEnd = p.jobRowEnd[j];
  const int bx = (pitch + 31) / 32, b = blockIdx.x - p.jobBlockBegin[j];
  const int qx = (b % bx) * 32 + (threadIdx.x & 31), qy0 = p.jobRowBegin[j] + (b / bx) * kPackCtaRows + (threadIdx.x >> 5) * kPackRows;
  if (qx >= pitch || qy0 >= rowEnd) return;
  // entry (qx, qy) = taps (x0, y0) (x1, y0) (x0, y1) (x1, y1) with y0 = clamp(qy - 1), y1 = clamp(qy): row k of the thread = clamp(qy0 - 1 + k)
  if (!p.jobLight[j]) {
    const LevelView lv = p.moments.lv[l];
    const int x0 = max(qx - 1, 0), x1 = min(qx, lv.w - 1);
    float t0[kPackRows + 1], t1[kPackRows + 1];
#pragma unroll
    for (int k = 0; k <= kPackRows; k++) {
      const float *row = reinterpret_cast<const float *>(lv.ptr + (size_t)min(max(qy0 - 1 + k, 0), lv.h - 1) * lv.pitch);
      t0[k] = __ldg(row + 2 * x0);
      t1[k] = __ldg(row + 2 * x1);
    }
    float4 *dst = side + (size_t)p.jobOfs[j] + (size_t)qy0 * pitch + qx;
#pragma unroll
    for (int e = 0; e < kPackRows; e++)
      if (qy0 + e < rowEnd) dst[(size_t)e * pitch] = make_float4(t0[e], t1[e], t0[e + 1], t1[e + 1]);
  } else {
    const LevelView lv = p.light.lv[l];
    const int x0 = max(qx - 1, 0), x1 = min(qx, lv.w - 1);
    uint2 t0[kPackRows + 1], t1[kPackRows + 1];
#pragma unroll
    for (int k = 0; k <= kPackRows; k++) {
      const uint2 *row = reinterpret_cast<const uint2 *>(lv.ptr + (size_t)min(max(qy0 - 1 + k, 0), lv.h - 1) * lv.pitch);
      t0[k] = __ldg(row + x0);
      t1[k] = __ldg(row + x1);
    }
    float4 *dst = side + (size_t)p.jobOfs[j] + 3 * ((size_t)qy0 * pitch + qx);
#pragma unroll
    for (int e = 0; e < kPackRows; e++) {
      if (qy0 + e >= rowEnd) break;
      const float4 t00 = Texel<F16>::unpack(t0[e]), t10 = Texel<F16>::unpack(t1[e]), t01 = Texel<F16>::unpack(t0[e + 1]), t11 = Texel<F16>::unpack(t1[e + 1]);
      float4 *d = dst + 3 * (size_t)e * pitch;
      d[0] = make_float4(t00.x, t00.y, t00.z, t10.x);
      d[1] = make_float4(t10.y, t10.z, t01.x, t01.y);
      d[2] = make_float4(t01.z, t11.x, t11.y, t11.z);
    }
  }
}

// Host: derive the fast tables. Returns false if the projection is not of the form the affine ray model assumes
// (homogeneous w of the far-plane point must not change sign over the screen) or the tables do not fit -> strict kernel.
bool buildFastTables(const GatherArgs &a, const GatherTables &t, FastTables *f) {
  if (t.maxSteps > kMaxSteps) return false;
  const double vpx = a.viewport[0], vpy = a.viewport[1];
  const float *m = a.invViewProj.m;
  double A4[4], B4[4], D4[4];
  for (int i = 0; i < 4; i++) {
    A4[i] = double(m[0 + i]) * (2.0 / vpx);
    B4[i] = double(m[4 + i]) * (2.0 / vpy);
    D4[i] = double(m[12 + i]) - double(m[0 + i]) - double(m[4 + i]) + double(m[8 + i]); // z = 1
  }
  const double cam[3] = {a.cam[0], a.cam[1], a.cam[2]};
  double ra[3], rb[3], rc[3];
  for (int i = 0; i < 3; i++) {
    ra[i] = A4[i] - cam[i] * A4[3];
    rb[i] = B4[i] - cam[i] * B4[3];
    rc[i] = D4[i] - cam[i] * D4[3];
  }
  // sign of w at the four corners
  const double corners[4][2] = {{0.0, 0.0}, {vpx, 0.0}, {0.0, vpy}, {vpx, vpy}};
  int sign = 0;
  for (int c = 0; c < 4; c++) {
    const double w = A4[3] * corners[c][0] + B4[3] * corners[c][1] + D4[3];
    const int s = w > 0.0 ? 1 : (w < 0.0 ? -1 : 0);
    if (s == 0 || (sign != 0 && s != sign)) return false;
    sign = s;
  }
  // normalise the scale of R so that |R|^2 stays well inside fp32 range
  const double centre[3] = {ra[0] * vpx * 0.5 + rb[0] * vpy * 0.5 + rc[0], ra[1] * vpx * 0.5 + rb[1] * vpy * 0.5 + rc[1],
                            ra[2] * vpx * 0.5 + rb[2] * vpy * 0.5 + rc[2]};
  const double len = std::sqrt(centre[0] * centre[0] + centre[1] * centre[1] + centre[2] * centre[2]);
  if (!(len > 0.0)) return false;
  for (int i = 0; i < 3; i++) {
    f->ra[i] = float(ra[i] / len);
    f->rb[i] = float(rb[i] / len);
    f->rc[i] = float(rc[i] / len);
  }
  f->raySign = float(sign);
  f->maxSteps = t.maxSteps;
  for (int n = 0; n < kMaxSteps; n++) f->iterThreshold[n] = t.iterThreshold[n];
  for (int idx = 0; idx < 16; idx++) {
    for (int d = 0; d < kGatherDirs; d++) {
      DirEntry &de = f->dir[idx][d];
      de.dirX = t.dirX[idx][d];
      de.dirY = t.dirY[idx][d];
      de.invDirX = 1.0f / de.dirX; // :87 (IEEE divide, like the shader's vec2(1) / rayDir)
      de.invDirY = 1.0f / de.dirY;
      double rd[3];
      for (int i = 0; i < 3; i++) rd[i] = (ra[i] * double(de.dirX) + rb[i] * double(de.dirY)) / len;
      de.rdX = float(rd[0]);
      de.rdY = float(rd[1]);
      de.rdZ = float(rd[2]);
      de.q2 = float(double(de.rdX) * de.rdX + double(de.rdY) * de.rdY + double(de.rdZ) * de.rdZ);
    }
    for (int k = 0; k < kMaxSteps; k++) {
      const float lambda = t.lod[idx][k];
      const float fl = std::floor(lambda);
      const int d0 = int(fl);
      StepRow &st = f->row[idx * kMaxSteps + k];
      st.off = t.pixelOffset[idx][k];
      st.frac = lambda - fl;
      st.l0 = d0;
      st.l1 = d0 + 1 < a.moments.count ? d0 + 1 : a.moments.count - 1;
    }
  }
  return true;
}

// Level geometry of the march rows and of the packer. Returns false if the two pyramids do not share one layout.
// Side pyramid layout: depth quads of levels kDepthQuadLevel0.., then light quads (three float4 per entry) of levels kLightQuadLevel0..
bool buildLevelGeometry(const GatherArgs &a, bool withSide, LevelGeom *geom, long long *sideFloat4s) {
  const double vpx = a.viewport[0], vpy = a.viewport[1];
  long long ofs = 0;
  for (int l = 0; l < a.moments.count; l++) {
    LevelGeom &g = geom[l];
    const LevelView &mv = a.moments.lv[l], &lv = a.light.lv[l];
    const long long mOfs = mv.ptr - a.moments.lv[0].ptr, lOfs = lv.ptr - a.light.lv[0].ptr;
    if (mv.w != lv.w || mv.h != lv.h || mv.pitch != lv.pitch || mOfs != lOfs || (mv.pitch % 8) != 0 || (mOfs % 8) != 0 ||
        (mOfs + (long long)mv.pitch * mv.h) / 8 > 0x7fffffffLL)
      return false;
    g.scaleX = float(double(mv.w) / vpx);
    g.scaleY = float(double(mv.h) / vpy);
    g.maxX = float(mv.w - 1);
    g.maxY = float(mv.h - 1);
    g.texOfs = int(mOfs / 8);
    g.texPitch = int(mv.pitch / 8);
    g.wm1 = mv.w - 1;
    g.hm1 = mv.h - 1;
    g.quadPitch = mv.w + 1;
    g.quadOfs = 0;
    g.lightOfs = 0;
    g.mode = 0;
    if (withSide && l >= kDepthQuadLevel0) {
      g.quadOfs = int(ofs);
      g.mode |= 1;
      ofs += (long long)(mv.w + 1) * (mv.h + 1);
    }
  }
  for (int l = kLightQuadLevel0; withSide && l < a.moments.count; l++) {
    LevelGeom &g = geom[l];
    g.lightOfs = int(ofs);
    g.mode |= 2;
    ofs += 3LL * (a.moments.lv[l].w + 1) * (a.moments.lv[l].h + 1);
  }
  if (ofs > 0x7fffffffLL) return false;
  if (sideFloat4s) *sideFloat4s = ofs;
  return true;
}

} // namespace

uint64_t gatherScratchBytes(uint32_t width, uint32_t height, uint32_t mips) {
  uint64_t entries = 0;
  for (uint32_t l = 0; l < mips && l < (uint32_t)kMaxGatherLevels; l++) {
    const uint64_t w = (width >> l) ? (width >> l) : 1, h = (height >> l) ? (height >> l) : 1;
    if (l >= (uint32_t)kDepthQuadLevel0) entries += (w + 1) * (h + 1);
    if (l >= (uint32_t)kLightQuadLevel0) entries += 3 * (w + 1) * (h + 1);
  }
  return (entries ? entries : 1) * sizeof(float4);
}

cudaError_t launchGatherPack(const GatherArgs &a, void *scratch, cudaStream_t s) {
  LevelGeom geom[kMaxGatherLevels];
  if (!buildLevelGeometry(a, true, geom, nullptr)) return cudaErrorInvalidValue;
  PackArgs p;
  p.moments = a.moments;
  p.light = a.light;
  p.jobs = 0;
  int blocks = 0;
  for (int light = 0; light < 2; light++)
    for (int l = light ? kLightQuadLevel0 : kDepthQuadLevel0; l < a.moments.count; l++) {
      // strip: the march reaches ~13 level-l texels beyond the strip; the coarse levels are rebuilt whole
      const int reach = 15; // entry row q reads rows q-1 and q: rows [strip-16, strip+16) of each level must be present (sharding.GATHER_REACH)
      const int entryRows = a.moments.lv[l].h + 1;
      int r0 = (a.rows.y0 >> l) - reach, r1 = ((a.rows.y1 + (1 << l) - 1) >> l) + reach + 1;
      const bool whole = l >= 6 || l == a.moments.count - 1; // the top level serves every clamped LOD
      if (whole || r0 < 0) r0 = 0;
      if (whole || r1 > entryRows) r1 = entryRows;
      const int j = p.jobs++;
      p.jobLevel[j] = l;
      p.jobLight[j] = light;
      p.jobOfs[j] = light ? geom[l].lightOfs : geom[l].quadOfs;
      p.jobPitch[j] = geom[l].quadPitch;
      p.jobRowBegin[j] = r0;
      p.jobRowEnd[j] = r1;
      p.jobBlockBegin[j] = blocks;
      blocks += ((geom[l].quadPitch + 31) / 32) * ((r1 - r0 + kPackCtaRows - 1) / kPackCtaRows);
    }
  p.jobBlockBegin[p.jobs] = blocks;
  if (blocks == 0) return cudaSuccess;
  packSidePyramidKernel<<<blocks, 256, 0, s>>>(p, static_cast<float4 *>(scratch));
  return cudaGetLastError();
}

cudaError_t launchGatherFast(const GatherArgs &a, const GatherTables &t, const void *scratch, cudaStream_t s) {
  if (a.rows.y1 <= a.rows.y0) return cudaSuccess;
  FastTables f;
  LevelGeom geom[kMaxGatherLevels];
  if (!buildFastTables(a, t, &f) || !buildLevelGeometry(a, scratch != nullptr, geom, nullptr)) return launchGatherStrict(a, t, s);
  for (int r = 0; r < 16 * kMaxSteps; r++) {
    f.row[r].g0 = geom[f.row[r].l0];
    f.row[r].g1 = geom[f.row[r].l1];
  }
  const int rowsSpan = a.rows.y1 - (a.rows.y0 & ~3);
  const dim3 tiles((a.indirect.w + kTile - 1) / kTile, (rowsSpan + kTile - 1) / kTile);
  int dev = 0, smCount = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, dev);
  const float4 *side = static_cast<const float4 *>(scratch);
  // Work-unit granularity (measured, profiles/README.md): the 16 pattern classes of a tile are split over `slices` CTAs that are
  // neighbours in blockIdx.x, so they run at the same time and share the tile's pyramid neighbourhood through L2 (DRAM traffic stays at
  // 1.2x the algorithmic bytes). Whole 4K / 8K frames are indifferent to 4, 8 or 16 slices (1.348 / 1.346 / 1.354 ms at 4K; 5.05 / 5.08 /
  // 5.12 ms at 8K); grids of a few waves — a multi-GPU row strip, a 1080p frame — gain 4-5 % from 16 (one class per CTA: the shortest
  // tail), r02q.
  static const int envSlices = getenv("LGCU_GATHER_SLICES") ? atoi(getenv("LGCU_GATHER_SLICES")) : 0; // A/B switch: 1, 2, 4, 8 or 16
  const long long tileCount = (long long)tiles.x * tiles.y;
  const int kSlices = (envSlices == 1 || envSlices == 2 || envSlices == 4 || envSlices == 8 || envSlices == 16) ? envSlices : (tileCount < 10LL * smCount ? 16 : 4);
  static const int envTiles = getenv("LGCU_GATHER_SMALL_TILES") ? atoi(getenv("LGCU_GATHER_SMALL_TILES")) : -1; // A/B switch
  const bool smallTiles = envTiles >= 0 ? envTiles != 0 : true; // 64x32 tiles of 128 threads everywhere: 8K whole frame 4.88 vs 5.04 ms with 64x64 / 256 threads (r02v)
  const dim3 gridBig(tiles.x * kSlices, tiles.y), gridSmall(tiles.x * kSlices, (rowsSpan + 31) / 32);
  // LGCU_GATHER_MINB: resident CTAs per SM the kernel is compiled for (A/B switch; 4 x 256 or 8 x 128 threads = 64 registers by default)
  static const int minb = getenv("LGCU_GATHER_MINB") ? atoi(getenv("LGCU_GATHER_MINB")) : 4;
  static const bool compact = getenv("LGCU_GATHER_WARP") ? atoi(getenv("LGCU_GATHER_WARP")) != 0 : false;
#define LGCU_LAUNCH_GATHER(SIDE, MINB_BIG, MINB_SMALL, COMPACT)                                              \
  do {                                                                                                        \
    if (smallTiles)                                                                                           \
      gatherFastKernel<SIDE, MINB_SMALL, 128, COMPACT><<<gridSmall, 128, 0, s>>>(a, f, side, kSlices);        \
    else                                                                                                      \
      gatherFastKernel<SIDE, MINB_BIG, kThreads, COMPACT><<<gridBig, kThreads, 0, s>>>(a, f, side, kSlices);  \
  } while (0)
  if (!side)
    LGCU_LAUNCH_GATHER(false, 4, 8, false);
  else if (minb == 5)
    LGCU_LAUNCH_GATHER(true, 5, 10, false);
  else if (compact)
    LGCU_LAUNCH_GATHER(true, 4, 8, true);
  else if (minb == 8)
    LGCU_LAUNCH_GATHER(true, 4, 8, false);
  else // default: 9 resident CTAs of 128 threads (56 registers, 36 warps / SM: -1.7 % against 8 x 64 registers at 4K, r02v) or 4 of 256
    LGCU_LAUNCH_GATHER(true, 4, 9, false);
#undef LGCU_LAUNCH_GATHER
  return cudaGetLastError();
}

} // namespace lgcu
