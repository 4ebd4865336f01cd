// k_gather_fast.cu — screen-space GI gather (K5), throughput variant for sm_100a.
//
// Same function as SH/SSVGI/indirectLighting.frag:114-272 (see k_gather_strict.cu for the line-by-line form), reorganised
// around what the hardware is short of. The pass is FP32-issue / L1 bound (≈25 pyramid samples per pixel),
// so the design removes instructions, not bytes:
//
//  * Pattern-coherent warps. The shader's 4x4 interleaved pattern gives every pixel with the same (x&3, y&3) the same
//    march directions, step offsets and LODs. A CTA owns a 32x32 pixel tile and warp w processes the 64 pixels of
//    pattern index w (2 per lane), so direction / offset / level / mip geometry are warp-uniform constant-bank operands,
//    the level branch is uniform, and at LOD >= 2 the lanes of a warp read adjacent texels.
//  * No transcendental per sample. pow/log (step offset, LOD, iteration count) come from the host tables shared with the
//    strict kernel (bit-identical level selection). Miss samples are rejected with a half-plane + cross-product test
//    instead of atan; sin(2h)/cos(2h) of ComputeHorizonContribution are evaluated algebraically from the horizon vector
//    (x,y): cos2h = (x²-y²)/(x²+y²), sin2h = 2xy/(x²+y²). atan2f runs on hits only (the 2·maxH - 2·h term needs the angle).
//  * The per-sample unprojection collapses to a few FMAs: the ray through pixel s is R(s) = Ra·sx + Rb·sy + Rc (affine
//    in pixel coordinates), along a march direction R(s) = R0 + off·Rd, and every dot product the horizon test needs is
//    affine in `off` with per-direction constants.
//  * What is numerically delicate is kept in the shader's order: the centre position is reconstructed from the D32 depth
//    exactly as the shader does (its fp32 cancellation noise is part of the reference result and is amplified by
//    1/sample distance), so the two variants see the same centre.
//  * Results are staged in shared memory and written with 16-byte coalesced stores.
#include <cmath>

#include "lgcu_kernels.h"

namespace lgcu {

namespace {

constexpr int kTile = 32;           // pixels per tile edge
constexpr int kThreads = 512;       // 16 warps = 16 pattern indices
constexpr uint32_t F16 = LGCU_FORMAT_R16G16B16A16_SFLOAT, D32 = LGCU_FORMAT_D32_SFLOAT;

// per-(pattern, step) and per-level constants derived on the host from GatherTables (see buildFastTables)
struct FastTables {
  float dirX[16][kGatherDirs], dirY[16][kGatherDirs];
  float rd[16][kGatherDirs][3]; // Rd = Ra*dirX + Rb*dirY
  float pixelOffset[16][kGatherMaxSteps];
  float lodFrac[16][kGatherMaxSteps];
  signed char lod0[16][kGatherMaxSteps], lod1[16][kGatherMaxSteps];
  float iterThreshold[kGatherMaxSteps];
  float levelScaleX[kMaxGatherLevels], levelScaleY[kMaxGatherLevels]; // w_l / viewport.x, h_l / viewport.y
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

struct Footprint { // bilinear footprint at one level, shared by the depth and the light fetch
  int o00, o10, o01, o11; // byte offsets / 8 (texel index) from the level base
  float a, b;
};

__device__ __forceinline__ Footprint footprint(const LevelView &l, float scaleX, float scaleY, float sx, float sy) {
  Footprint f;
  const float u = fmaf(sx, scaleX, -0.5f), v = fmaf(sy, scaleY, -0.5f);
  const float fu = floorf(u), fv = floorf(v);
  f.a = u - fu;
  f.b = v - fv;
  const int ix = (int)fu, iy = (int)fv;
  const int x0 = min(max(ix, 0), l.w - 1), x1 = min(max(ix + 1, 0), l.w - 1);
  const int y0 = min(max(iy, 0), l.h - 1), y1 = min(max(iy + 1, 0), l.h - 1);
  const int r0 = y0 * (int)(l.pitch >> 3), r1 = y1 * (int)(l.pitch >> 3); // both pyramid formats are 8 bytes per texel
  f.o00 = r0 + x0;
  f.o10 = r0 + x1;
  f.o01 = r1 + x0;
  f.o11 = r1 + x1;
  return f;
}

__device__ __forceinline__ float lerpf(float p, float q, float t) { return fmaf(q - p, t, p); }

__device__ __forceinline__ float fetchDepth(const LevelView &l, const Footprint &f) {
  const float2 *base = reinterpret_cast<const float2 *>(l.ptr);
  const float t00 = __ldg(&base[f.o00].x), t10 = __ldg(&base[f.o10].x), t01 = __ldg(&base[f.o01].x), t11 = __ldg(&base[f.o11].x);
  return lerpf(lerpf(t00, t10, f.a), lerpf(t01, t11, f.a), f.b);
}

__device__ __forceinline__ float3 fetchLight(const LevelView &l, const Footprint &f) {
  const uint2 *base = reinterpret_cast<const uint2 *>(l.ptr);
  const uint2 r00 = __ldg(&base[f.o00]), r10 = __ldg(&base[f.o10]), r01 = __ldg(&base[f.o01]), r11 = __ldg(&base[f.o11]);
  const float2 a00 = __half22float2(*reinterpret_cast<const __half2 *>(&r00.x)), a10 = __half22float2(*reinterpret_cast<const __half2 *>(&r10.x));
  const float2 a01 = __half22float2(*reinterpret_cast<const __half2 *>(&r01.x)), a11 = __half22float2(*reinterpret_cast<const __half2 *>(&r11.x));
  const float b00 = __low2float(*reinterpret_cast<const __half2 *>(&r00.y)), b10 = __low2float(*reinterpret_cast<const __half2 *>(&r10.y));
  const float b01 = __low2float(*reinterpret_cast<const __half2 *>(&r01.y)), b11 = __low2float(*reinterpret_cast<const __half2 *>(&r11.y));
  float3 r;
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

__device__ __forceinline__ float dotf(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

__global__ void __launch_bounds__(kThreads, 2) gatherFastKernel(const __grid_constant__ GatherArgs a, const __grid_constant__ FastTables tb) {
  __shared__ __align__(16) float4 stage[kTile * kTile]; // fp32 RGBA staging of the tile (16 KiB)

  const int lane = threadIdx.x & 31, idx = threadIdx.x >> 5; // warp index == pattern index (x&3) + 4*(y&3)
  const int tileX = blockIdx.x * kTile, tileY = a.rows.y0 + blockIdx.y * kTile;
  const float vpx = a.viewport[0], vpy = a.viewport[1];
  const float invVpx = 1.0f / vpx, invVpy = 1.0f / vpy;
  const V3 cam = v3(a.cam[0], a.cam[1], a.cam[2]);
  const V3 Ra = v3(tb.ra[0], tb.ra[1], tb.ra[2]), Rb = v3(tb.rb[0], tb.rb[1], tb.rb[2]), Rc = v3(tb.rc[0], tb.rc[1], tb.rc[2]);

#pragma unroll 1
  for (int half = 0; half < 2; half++) {
    const int lx = 4 * (lane & 7) + (idx & 3), ly = 4 * ((lane >> 3) + 4 * half) + (idx >> 2);
    const int x = tileX + lx, y = tileY + ly;
    float4 result = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    if (x < a.indirect.w && y < a.rows.y1) {
      const float px = (float)x + 0.5f, py = (float)y + 0.5f;
      // --- centre reconstruction in the shader's order (:116-132, :182) -----------------------------------------------
      const float cu = __fdiv_rn(px, vpx), cv = __fdiv_rn(py, vpy);
      const float4 ns = Texel<F16>::load(a.normal, x, y);
      const float zc = Texel<D32>::load(a.depthStencil, x, y).x;
      const float4 vc = mulMat4Exact(a.invViewProj, __fadd_rn(__fmul_rn(cu, 2.0f), -1.0f), __fadd_rn(__fmul_rn(cv, 2.0f), -1.0f), zc, 1.0f);
      const V3 C = v3(__fdiv_rn(vc.x, vc.w), __fdiv_rn(vc.y, vc.w), __fdiv_rn(vc.z, vc.w));
      const V3 N = v3(ns.x, ns.y, ns.z);
      const V3 E = v3(__fadd_rn(cam.x, -C.x), __fadd_rn(cam.y, -C.y), __fadd_rn(cam.z, -C.z)); // cam - C
      const float invLenE = __fdiv_rn(1.0f, __fsqrt_rn(dot3Exact(E, E)));
      const V3 eye = v3(__fmul_rn(E.x, invLenE), __fmul_rn(E.y, invLenE), __fmul_rn(E.z, invLenE));
      const float eN = dot3Exact(eye, N);
      // --- ray through the pixel: R0 ∝ far-plane point - cam ----------------------------------------------------------
      const V3 R0 = v3(fmaf(Ra.x, px, fmaf(Rb.x, py, Rc.x)), fmaf(Ra.y, px, fmaf(Rb.y, py, Rc.y)), fmaf(Ra.z, px, fmaf(Rb.z, py, Rc.z)));
      const float q0 = dotf(R0, R0);
      const float n0 = sqrtf(q0);
      const float xE = dotf(eye, E), xR0 = dotf(eye, R0);
      float sumX = 0.0f, sumY = 0.0f, sumZ = 0.0f;

#pragma unroll 1
      for (int d = 0; d < kGatherDirs; d++) {
        const float dirx = tb.dirX[idx][d], diry = tb.dirY[idx][d];
        const V3 Rd = v3(tb.rd[idx][d][0], tb.rd[idx][d][1], tb.rd[idx][d][2]);
        // tangent = normalize(rayDir(p + dir) - rayDir(p)) in a cancellation-free form (:181-183)
        const float r0rd = dotf(R0, Rd), q2 = dotf(Rd, Rd);
        const float q1 = 2.0f * r0rd;
        const float n1 = sqrtf(q0 + q1 + q2);
        const float g = (q1 + q2) * fastRcp(n0 + n1);
        V3 tanU = v3(fmaf(Rd.x, n0, -R0.x * g), fmaf(Rd.y, n0, -R0.y * g), fmaf(Rd.z, n0, -R0.z * g));
        const float tInv = tb.raySign * fastRsqrt(dotf(tanU, tanU));
        const V3 tang = v3(tanU.x * tInv, tanU.y * tInv, tanU.z * tInv);
        const float tN = dotf(tang, N);
        // initial horizon from the surface normal (:193-198)
        const V3 bn = cross3(eye, tang); // -cross(tangent, eye)
        const V3 q = cross3(bn, N);
        float mx = dotf(q, eye), my = dotf(q, tang);
        float maxH = atan2f(my, mx);
        float invM = fastRcp(fmaxf(fmaf(mx, mx, my * my), 1e-37f));
        float c2m = (mx * mx - my * my) * invM, s2m = 2.0f * mx * my * invM;
        // BoxRayCast + iteration count (:83-99, :202-212), exact like the strict kernel
        const float ivx = __fdiv_rn(1.0f, dirx), ivy = __fdiv_rn(1.0f, diry);
        const float t1 = __fmul_rn(0.0f - px, ivx), t2 = __fmul_rn(vpx - px, ivx), t3 = __fmul_rn(0.0f - py, ivy), t4 = __fmul_rn(vpy - py, ivy);
        const float path = fabsf(glmMin(glmMax(t1, t2), glmMax(t3, t4)));
        int iterations = 0;
        for (int n = 0; n < tb.maxSteps; n++) iterations += (path >= tb.iterThreshold[n]) ? 1 : 0;
        // ambient term 0.01 * HC(0, maxH) (:209): cos(0) = 1, sin(0) = 0
        const float hc0 = 0.25f * eN * (1.0f - c2m) + 0.25f * tN * (2.0f * maxH - s2m);
        float Lx = 0.01f * hc0, Ly = Lx, Lz = Lx;
        // per-direction affine coefficients of the horizon vector
        const float yE = dotf(tang, E), yR0 = dotf(tang, R0), xRd = dotf(eye, Rd), yRd = dotf(tang, Rd);

#pragma unroll 1
        for (int k = 0; k < iterations; k++) {
          const float off = tb.pixelOffset[idx][k];
          const int d0 = tb.lod0[idx][k], d1 = tb.lod1[idx][k];
          const float frac = tb.lodFrac[idx][k];
          const float sx = fmaf(dirx, off, px), sy = fmaf(diry, off, py);
          const Footprint f0 = footprint(a.moments.lv[d0], tb.levelScaleX[d0], tb.levelScaleY[d0], sx, sy);
          float z = fetchDepth(a.moments.lv[d0], f0);
          Footprint f1 = f0;
          if (frac > 0.0f) { // warp-uniform
            f1 = footprint(a.moments.lv[d1], tb.levelScaleX[d1], tb.levelScaleY[d1], sx, sy);
            z = fmaf(frac, fetchDepth(a.moments.lv[d1], f1) - z, z); // (1-frac)*lo + frac*hi
          }
          const float rs = fastRsqrt(fmaf(off, fmaf(off, q2, q1), q0)); // 1 / |R(s)|
          const float zs = tb.raySign * z * rs;
          const float hx = fmaf(zs, fmaf(off, xRd, xR0), xE); // dot(eye, P - C)
          const float hy = fmaf(zs, fmaf(off, yRd, yR0), yE); // dot(tangent, P - C)
          // h < maxH for angles in (-pi, pi]: different half planes decide directly, otherwise the cross product does
          const bool lower = hy < 0.0f, lowerM = my < 0.0f;
          const bool maybe = (lower != lowerM) ? lower : (hx * my - hy * mx > 0.0f);
          if (maybe) {
            const float h = atan2f(hy, hx);
            if (h < maxH) { // :254
              const float su = sx * invVpx, sv = sy * invVpy;
              float side = saturatef((1.0f - su) * 10.0f);
              side *= saturatef(su * 10.0f);
              side *= saturatef((1.0f - sv) * 10.0f);
              side *= saturatef(sv * 10.0f);
              float3 ls = fetchLight(a.light.lv[d0], f0);
              if (frac > 0.0f) {
                const float3 hi = fetchLight(a.light.lv[d1], f1);
                ls.x = fmaf(frac, hi.x - ls.x, ls.x);
                ls.y = fmaf(frac, hi.y - ls.y, ls.y);
                ls.z = fmaf(frac, hi.z - ls.z, ls.z);
              }
              const float inv = fastRcp(fmaxf(fmaf(hx, hx, hy * hy), 1e-37f));
              const float c2 = (hx * hx - hy * hy) * inv, s2 = 2.0f * hx * hy * inv;
              const float hc = 0.25f * eN * (c2 - c2m) + 0.25f * tN * (((2.0f * maxH - 2.0f * h) - s2m) + s2);
              const float c = hc * side;
              Lx = fmaf(ls.x - 0.01f, c, Lx);
              Ly = fmaf(ls.y - 0.01f, c, Ly);
              Lz = fmaf(ls.z - 0.01f, c, Lz);
              maxH = h;
              c2m = c2;
              s2m = s2;
              mx = hx;
              my = hy;
            }
          }
        }
        sumX = fmaf(0.5f, Lx, sumX);
        sumY = fmaf(0.5f, Ly, sumY);
        sumZ = fmaf(0.5f, Lz, sumZ);
      }
      result = make_float4(sumX, sumY, sumZ, 1.0f);
    }
    stage[ly * kTile + lx] = result;
  }
  __syncthreads();

  // coalesced write-out of the tile
  if (a.outFormat == F16) {
    // 32 rows x 32 px x 8 B: 16 threads per row, 16 bytes (2 px) each
    const int row = threadIdx.x >> 4, col = (threadIdx.x & 15) * 2;
    const int x = tileX + col, y = tileY + row;
    if (y < a.rows.y1 && x < a.indirect.w) {
      const uint2 p0 = Texel<F16>::pack(stage[row * kTile + col]);
      unsigned char *dst = a.indirect.ptr + (size_t)y * a.indirect.pitch + (size_t)x * 8;
      if (x + 1 < a.indirect.w) {
        const uint2 p1 = Texel<F16>::pack(stage[row * kTile + col + 1]);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(p0.x, p0.y, p1.x, p1.y);
      } else {
        *reinterpret_cast<uint2 *>(dst) = p0;
      }
    }
  } else {
    for (int i = threadIdx.x; i < kTile * kTile; i += kThreads) {
      const int row = i >> 5, col = i & 31;
      const int x = tileX + col, y = tileY + row;
      if (y < a.rows.y1 && x < a.indirect.w) reinterpret_cast<float4 *>(a.indirect.ptr + (size_t)y * a.indirect.pitch)[x] = stage[i];
    }
  }
}

// Host: derive the fast tables. Returns false if the projection is not of the form the affine ray model assumes
// (homogeneous w of the far-plane point must not change sign over the screen) -> caller uses the strict kernel.
bool buildFastTables(const GatherArgs &a, const GatherTables &t, FastTables *f) {
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
  for (int n = 0; n < kGatherMaxSteps; n++) f->iterThreshold[n] = t.iterThreshold[n];
  for (int idx = 0; idx < 16; idx++) {
    for (int d = 0; d < kGatherDirs; d++) {
      f->dirX[idx][d] = t.dirX[idx][d];
      f->dirY[idx][d] = t.dirY[idx][d];
      for (int i = 0; i < 3; i++) f->rd[idx][d][i] = float((ra[i] * double(t.dirX[idx][d]) + rb[i] * double(t.dirY[idx][d])) / len);
    }
    for (int k = 0; k < kGatherMaxSteps; k++) {
      const float lambda = t.lod[idx][k];
      const float fl = std::floor(lambda);
      const int d0 = int(fl);
      f->pixelOffset[idx][k] = t.pixelOffset[idx][k];
      f->lodFrac[idx][k] = lambda - fl;
      f->lod0[idx][k] = (signed char)d0;
      f->lod1[idx][k] = (signed char)(d0 + 1 < a.moments.count ? d0 + 1 : a.moments.count - 1);
    }
  }
  for (int l = 0; l < kMaxGatherLevels; l++) {
    const int w = l < a.moments.count ? a.moments.lv[l].w : 1, h = l < a.moments.count ? a.moments.lv[l].h : 1;
    f->levelScaleX[l] = float(double(w) / vpx);
    f->levelScaleY[l] = float(double(h) / vpy);
  }
  return true;
}

} // namespace

cudaError_t launchGatherFast(const GatherArgs &a, const GatherTables &t, cudaStream_t s) {
  if (a.rows.y1 <= a.rows.y0) return cudaSuccess;
  FastTables f;
  // the tile / pattern mapping needs the strip to start on a pattern row; both pyramids must share their geometry
  bool ok = (a.rows.y0 % 4) == 0 && buildFastTables(a, t, &f);
  for (int l = 0; ok && l < a.moments.count; l++)
    ok = a.moments.lv[l].w == a.light.lv[l].w && a.moments.lv[l].h == a.light.lv[l].h && a.moments.lv[l].pitch == a.light.lv[l].pitch &&
         (a.moments.lv[l].pitch % 8) == 0;
  if (!ok) return launchGatherStrict(a, t, s);
  const dim3 grid((a.indirect.w + kTile - 1) / kTile, (a.rows.y1 - a.rows.y0 + kTile - 1) / kTile);
  gatherFastKernel<<<grid, kThreads, 0, s>>>(a, f);
  return cudaGetLastError();
}

} // namespace lgcu
