/*
 * gather_noise_probe.c — TEST INFRASTRUCTURE. The GI gather (SH/SSVGI/indirectLighting.frag:114-272) evaluated in binary64.
 *
 * Purpose: measure the reference's OWN fp32 rounding noise. orc_gi_gather (ssvgi_oracle.c) is the shader in binary32, bit for
 * bit; this file is the same formula on the same stored inputs (fp16 / fp32 texels, fp32 UBO) with every continuous quantity in
 * binary64. |orc_gi_gather - orc_gi_gather_f64| is therefore how far the reference sits from the value its own formula defines —
 * the floor below which no independent evaluation order can agree with it. The discrete contracts stay those of the fp32
 * shader: pattern index, march direction table, step offset, LOD (level pair + fraction) and iteration count are computed in
 * binary32 exactly as orc_gi_gather computes them (they are table inputs of both CUDA kernels as well).
 * flags bit 0: take the centre position C from the fp32 shader-order reconstruction (what both CUDA kernels do), so that the
 * difference isolates the noise of the march itself (the tangent still comes from the binary64 pair of unprojections). flags bit 1: take inverse(proj*view) and the camera position from the fp32
 * shader-order computation (frame constants, which the CUDA kernels receive in fp32 as well).
 * branchCut (optional, one byte per pixel, row 0 = rows->y0): set to 1 where a horizon angle of the pixel — the initial one from the
 * surface normal (:194-198) or a sample's (:250-252) — lies within ORC_BRANCH_CUT_RAD of +-pi. atan() jumps by 2 pi there, `h < maxH`
 * (:254) flips for ALL following samples, and the shader's result changes by O(radiance) under a 1-ulp perturbation: the formula is
 * discontinuous at that pixel, so it has no value to agree with (the fp32 shader and the binary64 evaluation themselves land on
 * different sides). Parity tests report these pixels and leave them out, like the flat windows of the radius-2 denoiser.
 * Output: unrounded RGB as float32, `outPitchFloats` floats per row, row 0 = rows->y0.
 */
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <string.h>

#include "ssvgi_oracle.h"
#include "texel_codec.h"

void orc_inv_view_proj_f32(const lgcu_indirect_lighting_data *params, float *invViewProj, float *cam); /* ssvgi_oracle.c: :122-127 in glm order */

#define ORC_BRANCH_CUT_RAD 1e-3 /* ~10x the fp32 noise of the reference's tangent at 8K (2^-23 / pixel angle) */
int orc_debug_px = -1, orc_debug_py = -1; /* set from a debugger / ctypes to trace one pixel's march on stderr */

typedef struct { double x, y, z; } d3;
static inline d3 d3_sub(d3 a, d3 b) { d3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline d3 d3_add(d3 a, d3 b) { d3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline d3 d3_scale(d3 a, double s) { d3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline double d3_dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline d3 d3_cross(d3 x, d3 y) { d3 r = {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; return r; }
static inline d3 d3_normalize(d3 a) { return d3_scale(a, 1.0 / sqrt(d3_dot(a, a))); }
static inline double d_sat(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }

/* general 4x4 inverse in binary64 (Gauss-Jordan with partial pivoting), column-major in and out */
static void m4d_inverse(const double *m, double *inv) {
  double a[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) { a[r][c] = m[c * 4 + r]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
  for (int i = 0; i < 4; i++) {
    int p = i;
    for (int r = i + 1; r < 4; r++) if (fabs(a[r][i]) > fabs(a[p][i])) p = r;
    if (p != i) for (int c = 0; c < 8; c++) { double t = a[i][c]; a[i][c] = a[p][c]; a[p][c] = t; }
    const double s = 1.0 / a[i][i];
    for (int c = 0; c < 8; c++) a[i][c] *= s;
    for (int r = 0; r < 4; r++) if (r != i) { const double f = a[r][i]; for (int c = 0; c < 8; c++) a[r][c] -= f * a[i][c]; }
  }
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) inv[c * 4 + r] = a[r][4 + c];
}
static inline d3 d_unproject(double sx, double sy, double sz, const double *M) {
  const double x = sx * 2.0 - 1.0, y = sy * 2.0 - 1.0;
  const double vx = M[0] * x + M[4] * y + M[8] * sz + M[12], vy = M[1] * x + M[5] * y + M[9] * sz + M[13];
  const double vz = M[2] * x + M[6] * y + M[10] * sz + M[14], vw = M[3] * x + M[7] * y + M[11] * sz + M[15];
  d3 r = {vx / vw, vy / vw, vz / vw};
  return r;
}
static void d_bilinear(const lgcu_image *img, uint32_t lod, double u_, double v_, double out[4]) {
  int w, h;
  orc_level_size(img, lod, &w, &h);
  const double u = u_ * (double)w - 0.5, v = v_ * (double)h - 0.5, fu = floor(u), fv = floor(v), a = u - fu, b = v - fv;
  const int x0 = orc_clampi((int)fu, 0, w - 1), x1 = orc_clampi((int)fu + 1, 0, w - 1);
  const int y0 = orc_clampi((int)fv, 0, h - 1), y1 = orc_clampi((int)fv + 1, 0, h - 1);
  float t00[4], t10[4], t01[4], t11[4];
  orc_load_texel(img, lod, x0, y0, t00);
  orc_load_texel(img, lod, x1, y0, t10);
  orc_load_texel(img, lod, x0, y1, t01);
  orc_load_texel(img, lod, x1, y1, t11);
  for (int c = 0; c < 4; c++) {
    const double top = t00[c] + ((double)t10[c] - t00[c]) * a, bot = t01[c] + ((double)t11[c] - t01[c]) * a;
    out[c] = top + (bot - top) * b;
  }
}
/* lod is the fp32 shader value: the level pair and the mip weight are part of the discrete contract */
static void d_texture_lod(const lgcu_image *img, double u, double v, float lod, double out[4]) {
  const float last = (float)(img->mipCount - 1);
  float lambda = lod;
  if (!(lambda > 0.0f)) lambda = 0.0f;
  if (lambda > last) lambda = last;
  const float fd = floorf(lambda), delta = lambda - fd;
  const uint32_t d = (uint32_t)fd, d1 = d + 1 < img->mipCount ? d + 1 : img->mipCount - 1;
  double lo[4], hi[4];
  d_bilinear(img, d, u, v, lo);
  d_bilinear(img, d1, u, v, hi);
  for (int c = 0; c < 4; c++) out[c] = (1.0 - (double)delta) * lo[c] + (double)delta * hi[c];
}
static inline double d_hc(d3 eye, d3 tang, d3 n, double lo, double hi) { /* :44-49 */
  return 0.25 * d3_dot(eye, n) * (-cos(2.0 * hi) + cos(2.0 * lo)) + 0.25 * d3_dot(tang, n) * (2.0 * hi - 2.0 * lo - sin(2.0 * hi) + sin(2.0 * lo));
}

int orc_gi_gather_f64(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight, const lgcu_image *blurredDepthMoments,
                      const lgcu_image *normal, const lgcu_image *depthStencil, float *outRgb, uint64_t outPitchFloats, const lgcu_rows *rows,
                      uint32_t flags, uint8_t *branchCut, uint64_t branchCutPitch) {
  int w, h;
  orc_level_size(depthStencil, 0, &w, &h);
  int y0 = rows ? (int)rows->y0 : 0, y1 = rows ? (int)rows->y1 : h;
  if (y1 > h) y1 = h;
  const float vpxf = params->viewportExtent[0];
  const double vpx = vpxf, vpy = params->viewportExtent[1];
  double V[16], VP[16], iVP[16], iV[16];
  for (int i = 0; i < 16; i++) V[i] = params->viewMatrix.m[i];
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s += (double)params->projMatrix.m[k * 4 + i] * (double)params->viewMatrix.m[j * 4 + k];
      VP[j * 4 + i] = s;
    }
  m4d_inverse(VP, iVP);
  m4d_inverse(V, iV);
  d3 cam = {iV[12], iV[13], iV[14]};
  float iVPf[16], camf[3];
  orc_inv_view_proj_f32(params, iVPf, camf);
  if (flags & 2u) { /* frame constants (inverse(viewProj), camera position) as the fp32 shader computes them */
    for (int i = 0; i < 16; i++) iVP[i] = iVPf[i];
    cam.x = camf[0]; cam.y = camf[1]; cam.z = camf[2];
  }
  const float nearStepSize = vpxf / 1000.0f;
#pragma omp parallel for schedule(dynamic, 2)
  for (int y = y0; y < y1; y++)
    for (int x = 0; x < w; x++) {
      const float pxf = (float)x + 0.5f, pyf = (float)y + 0.5f;
      const double px = pxf, py = pyf;
      float ns[4], ds[4];
      orc_load_texel(normal, 0, x, y, ns);
      orc_load_texel(depthStencil, 0, x, y, ds);
      const d3 Cray = d_unproject(px / vpx, py / vpy, ds[0], iVP); /* the centre in binary64: defines the pixel's ray for the tangent */
      d3 C = Cray;
      if (flags & 1u) { /* centre position exactly as the fp32 shader reconstructs it (:118, :129; glm order) */
        const float cu = pxf / vpxf, cv = pyf / (float)vpy, X = cu * 2.0f - 1.0f, Y = cv * 2.0f - 1.0f, Z = ds[0];
        const float *M = iVPf;
        const float vx = (M[0] * X + M[4] * Y) + (M[8] * Z + M[12] * 1.0f), vy = (M[1] * X + M[5] * Y) + (M[9] * Z + M[13] * 1.0f);
        const float vz = (M[2] * X + M[6] * Y) + (M[10] * Z + M[14] * 1.0f), vw = (M[3] * X + M[7] * Y) + (M[11] * Z + M[15] * 1.0f);
        C.x = vx / vw; C.y = vy / vw; C.z = vz / vw;
      }
      const d3 N = {ns[0], ns[1], ns[2]};
      const int index = ((int)pxf % 4) + ((int)pyf % 4) * 4;
      uint32_t b = ((uint32_t)index << 16) | ((uint32_t)index >> 16);
      b = ((b & 0x55555555u) << 1) | ((b & 0xAAAAAAAAu) >> 1);
      b = ((b & 0x33333333u) << 2) | ((b & 0xCCCCCCCCu) >> 2);
      b = ((b & 0x0F0F0F0Fu) << 4) | ((b & 0xF0F0F0F0u) >> 4);
      b = ((b & 0x00FF00FFu) << 8) | ((b & 0xFF00FF00u) >> 8);
      const float angOffset = (float)index / 16.0f, linOffset = (float)b / 4294967296.0f;
      const float pixelAngOffset = 1.57075f * angOffset;
      double sum[3] = {0.0, 0.0, 0.0};
      int nearCut = 0;
      for (int dirIndex = 0; dirIndex < 4; dirIndex++) {
        const float screenAng = pixelAngOffset + (1.57075f * (float)dirIndex);
        const float dirxf = cosf(screenAng), diryf = sinf(screenAng); /* direction table: fp32 contract */
        const double dirx = dirxf, diry = diryf;
        const d3 O = d_unproject((px + dirx) / vpx, (py + diry) / vpy, ds[0], iVP);
        const d3 eye = d3_normalize(d3_sub(cam, C));
        /* tangent from a CONSISTENT pair of unprojections: in the fp32 shader C and O carry the same cancellation error of
         * M*(x, y, zc, 1) (same zc), which drops out of the difference of the two ray directions; mixing the fp32 C with a binary64 O
         * would put that error (1e-4 relative under a rotated camera) into the tangent instead */
        const d3 tang = d3_normalize(d3_sub(d3_normalize(d3_sub(O, cam)), d3_normalize(d3_sub(Cray, cam))));
        const d3 bn = d3_cross(tang, eye);
        const d3 nbn = {-bn.x, -bn.y, -bn.z};
        const d3 q = d3_cross(nbn, N);
        double maxH = atan2(d3_dot(q, tang), d3_dot(q, eye));
        if (fabs(maxH) > 3.14159265358979323846 - ORC_BRANCH_CUT_RAD) nearCut = 1;
        if (x == orc_debug_px && y == orc_debug_py)
          fprintf(stderr, "f64 d%d tang %.9f %.9f %.9f eye %.9f %.9f %.9f N %.6f %.6f %.6f q %.3e %.3e %.3e C %.7f %.7f %.7f\n", dirIndex, tang.x, tang.y, tang.z, eye.x, eye.y,
                  eye.z, N.x, N.y, N.z, q.x, q.y, q.z, C.x, C.y, C.z);
        const float ivx = 1.0f / dirxf, ivy = 1.0f / diryf; /* iteration count: fp32 contract */
        const float t1 = (0.0f - pxf) * ivx, t2 = (vpxf - pxf) * ivx, t3 = (0.0f - pyf) * ivy, t4 = ((float)vpy - pyf) * ivy;
        const float m12 = (t1 < t2) ? t2 : t1, m34 = (t3 < t4) ? t4 : t3;
        const float totalPixelPath = fabsf((m34 < m12) ? m34 : m12);
        const int iterations = (int)(logf(totalPixelPath / nearStepSize) / 0.944197714328765869140625f) + 1;
        double L = 0.01 * d_hc(eye, tang, N, 0.0, maxH), Lc[3] = {L, L, L};
        for (int k = 0; k < iterations; k++) {
          const float pixelOffset = ((nearStepSize * powf(2.57075f, (float)k + linOffset)) + 1.0f) - nearStepSize; /* step table: fp32 contract */
          const float lod = (logf(fmaxf(0.0f, (1.57075f * (pixelOffset - 1.0f)) * 0.5f)) / 0.693147182464599609375f) + -2.0f;
          const double su = (px + dirx * (double)pixelOffset) / vpx, sv = (py + diry * (double)pixelOffset) / vpy;
          const double side = d_sat((1.0 - su) * 10.0) * d_sat(su * 10.0) * d_sat((1.0 - sv) * 10.0) * d_sat(sv * 10.0);
          double zs[4];
          d_texture_lod(blurredDepthMoments, su, sv, lod, zs);
          const d3 ray = d3_normalize(d3_sub(d_unproject(su, sv, 1.0, iVP), cam));
          const d3 delta = d3_sub(d3_add(cam, d3_scale(ray, zs[0])), C);
          const double sampleH = atan2(d3_dot(tang, delta), d3_dot(eye, delta));
          if (fabs(sampleH) > 3.14159265358979323846 - ORC_BRANCH_CUT_RAD) nearCut = 1;
          if (x == orc_debug_px && y == orc_debug_py)
            fprintf(stderr, "f64 d%d k%d off %.6f lod %.4f z %.7f h %.7f maxH %.7f hit %d dx %.3e dy %.3e\n", dirIndex, k, pixelOffset, lod, zs[0], sampleH, maxH,
                    sampleH < maxH, d3_dot(eye, delta), d3_dot(tang, delta));
          if (sampleH < maxH) {
            double ls[4];
            d_texture_lod(blurredDirectLight, su, sv, lod, ls);
            const double c = d_hc(eye, tang, N, sampleH, maxH) * side;
            for (int ch = 0; ch < 3; ch++) Lc[ch] += ls[ch] * c - 0.01 * c;
            if (x == orc_debug_px && y == orc_debug_py) fprintf(stderr, "    c %.6e ls %.4f %.4f %.4f side %.4f\n", c, ls[0], ls[1], ls[2], side);
            maxH = sampleH;
          }
        }
        for (int ch = 0; ch < 3; ch++) sum[ch] += (2.0 * Lc[ch]) / 4.0;
      }
      float *o = outRgb + (uint64_t)(y - y0) * outPitchFloats + (uint64_t)x * 3u;
      o[0] = (float)sum[0]; o[1] = (float)sum[1]; o[2] = (float)sum[2];
      if (branchCut) branchCut[(uint64_t)(y - y0) * branchCutPitch + (uint64_t)x] = (uint8_t)nearCut;
    }
  return LGCU_OK;
}
