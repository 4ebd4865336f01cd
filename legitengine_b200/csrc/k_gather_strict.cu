// k_gather_strict.cu — screen-space GI gather (K5), parity variant.
//
// SH/SSVGI/indirectLighting.frag:114-272 evaluated per pixel in the shader's own operation order (compiled with
// -fmad=false, IEEE divide / sqrt), with three hoists that do not change a single bit:
//   * inverse(projMatrix * viewMatrix) and the camera position are frame constants (:122-127) -> GatherArgs;
//   * the march direction, step offset and LOD depend only on (pattern index, direction, step) and viewportSize.x
//     (:155-178, :217, :234-235) -> GatherTables, built on the host with the libm the CPU oracle uses;
//   * iterationsCount = int(log(|tmax| / near) / log(2.57075)) + 1 (:212) is monotone in |tmax| -> a threshold
//     table, so the per-pixel count is an exact comparison chain instead of a device logf.
// What remains transcendental on the device is atan (horizon angles) and sin/cos (ComputeHorizonContribution); CUDA's
// atan2f/sinf/cosf are 1-2 ulp functions like the CPU's, and the result is continuous in them (the h < maxH branch
// contributes HC(h, maxH) -> 0 at the flip), so the output agrees with the oracle to a few fp32 ulp before the fp16
// render-target rounding. This kernel defines the algorithmic instruction count of the pass; k_gather_fast.cu is
// the throughput variant.
#include "lgcu_kernels.h"

namespace lgcu {

namespace {

constexpr int kBlockX = 32, kBlockY = 8;
constexpr uint32_t F16 = LGCU_FORMAT_R16G16B16A16_SFLOAT, RG32 = LGCU_FORMAT_R32G32_SFLOAT, D32 = LGCU_FORMAT_D32_SFLOAT;

// ComputeHorizonContribution :44-49
__device__ __forceinline__ float horizonContribution(float eyeDotN, float tanDotN, float minAngle, float maxAngle) {
  return ((0.25f * eyeDotN) * ((-cosf(2.0f * maxAngle)) + cosf(2.0f * minAngle))) +
         ((0.25f * tanDotN) * ((((2.0f * maxAngle) - (2.0f * minAngle)) - sinf(2.0f * maxAngle)) + sinf(2.0f * minAngle)));
}

// textureLod(...).r on the RG32F moments pyramid (Appendix B), exact order
__device__ __forceinline__ float bilinearR(const LevelView &l, float u, float v) {
  const BilinearTaps t = bilinearTaps(l, u, v);
  const unsigned char *r0 = l.ptr + (size_t)t.y0 * l.pitch, *r1 = l.ptr + (size_t)t.y1 * l.pitch;
  const float t00 = __ldg(reinterpret_cast<const float *>(r0) + 2 * t.x0), t10 = __ldg(reinterpret_cast<const float *>(r0) + 2 * t.x1);
  const float t01 = __ldg(reinterpret_cast<const float *>(r1) + 2 * t.x0), t11 = __ldg(reinterpret_cast<const float *>(r1) + 2 * t.x1);
  return lerpExact(lerpExact(t00, t10, t.a), lerpExact(t01, t11, t.a), t.b);
}

__global__ void __launch_bounds__(kBlockX *kBlockY) gatherStrictKernel(const __grid_constant__ GatherArgs a, const __grid_constant__ GatherTables tb) {
  const int x = blockIdx.x * kBlockX + threadIdx.x, y = a.rows.y0 + blockIdx.y * kBlockY + threadIdx.y;
  if (x >= a.indirect.w || y >= a.rows.y1) return;
  const float vpx = a.viewport[0], vpy = a.viewport[1];
  const float px = (float)x + 0.5f, py = (float)y + 0.5f; // gl_FragCoord.xy :116
  const float cu = px / vpx, cv = py / vpy;               // :118
  const float4 ns = Texel<F16>::load(a.normal, x, y);     // :119 (centre tap)
  const float zc = Texel<D32>::load(a.depthStencil, x, y).x; // :120
  const V3 cam = v3(a.cam[0], a.cam[1], a.cam[2]);
  const V3 C = unproject(cu, cv, zc, a.invViewProj); // :129
  const V3 N = v3(ns.x, ns.y, ns.z);                  // :132
  const int idx = (x & 3) + (y & 3) * 4;              // :155, :161
  const V3 eye = normalize3(cam - C);                 // :182
  const V3 centreDir = normalize3(C - cam);           // :183 (second operand)
  const float eyeDotN = dot3(eye, N);
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  for (int d = 0; d < kGatherDirs; d++) { // :174
    const float dirx = tb.dirX[idx][d], diry = tb.dirY[idx][d]; // :177-178
    const V3 O = unproject((px + dirx * 1.0f) / vpx, (py + diry * 1.0f) / vpy, zc, a.invViewProj); // :181
    const V3 tang = normalize3(normalize3(O - cam) - centreDir);                                   // :183
    const float tanDotN = dot3(tang, N);
    const V3 q = cross3(-cross3(tang, eye), N);            // :193
    float maxH = atan2f(dot3(q, tang), dot3(q, eye));      // :194-198
    const float ivx = 1.0f / dirx, ivy = 1.0f / diry;      // BoxRayCast :83-99
    const float t1 = (0.0f - px) * ivx, t2 = (vpx - px) * ivx, t3 = (0.0f - py) * ivy, t4 = (vpy - py) * ivy;
    const float path = fabsf(glmMin(glmMax(t1, t2), glmMax(t3, t4))); // :202-204
    int iterations = 0;                                      // :212
    for (int n = 0; n < tb.maxSteps; n++) iterations += (path >= tb.iterThreshold[n]) ? 1 : 0;
    const float hc0 = horizonContribution(eyeDotN, tanDotN, 0.0f, maxH);
    float Lx = 0.01f * hc0, Ly = Lx, Lz = Lx; // :209
    for (int k = 0; k < iterations; k++) { // :214
      const float off = tb.pixelOffset[idx][k];                         // :217
      const float su = (px + dirx * off) / vpx, sv = (py + diry * off) / vpy; // :218-219
      const float lambda = tb.lod[idx][k];                              // :234-235 + sampler / view clamp
      const float fl = floorf(lambda), delta = lambda - fl;
      const int d0 = (int)fl, d1 = min(d0 + 1, a.moments.count - 1);
      const float z = (1.0f - delta) * bilinearR(a.moments.lv[d0], su, sv) + delta * bilinearR(a.moments.lv[d1], su, sv); // :240
      const V3 ray = normalize3(unproject(su, sv, 1.0f, a.invViewProj) - cam);
      const V3 delta3 = (cam + ray * z) - C;                            // :241, :249
      const float h = atan2f(dot3(tang, delta3), dot3(eye, delta3));    // :250-252
      if (h < maxH) {                                                   // :254
        float side = 1.0f;                                              // :220-228
        const float invWidth = 1.0f / 0.1f;
        side *= saturatef((1.0f - su) * invWidth);
        side *= saturatef(su * invWidth);
        side *= saturatef((1.0f - sv) * invWidth);
        side *= saturatef(sv * invWidth);
        const float4 lo = bilinear<F16>(a.light.lv[d0], su, sv), hi = bilinear<F16>(a.light.lv[d1], su, sv); // :256
        const float lx = (1.0f - delta) * lo.x + delta * hi.x, ly = (1.0f - delta) * lo.y + delta * hi.y,
                    lz = (1.0f - delta) * lo.z + delta * hi.z;
        const float c = horizonContribution(eyeDotN, tanDotN, h, maxH) * side; // :258
        Lx += lx * c; Ly += ly * c; Lz += lz * c;                        // :261
        Lx -= 0.01f * c; Ly -= 0.01f * c; Lz -= 0.01f * c;               // :262
        maxH = h;                                                        // :263
      }
    }
    sx += (2.0f * Lx) / 4.0f; sy += (2.0f * Ly) / 4.0f; sz += (2.0f * Lz) / 4.0f; // :268
  }
  storeColor(a.outFormat, a.indirect, x, y, make_float4(sx, sy, sz, 1.0f)); // :270
}

} // namespace

cudaError_t launchGatherStrict(const GatherArgs &a, const GatherTables &t, cudaStream_t s) {
  if (a.rows.y1 <= a.rows.y0) return cudaSuccess;
  const dim3 grid((a.indirect.w + kBlockX - 1) / kBlockX, (a.rows.y1 - a.rows.y0 + kBlockY - 1) / kBlockY);
  gatherStrictKernel<<<grid, dim3(kBlockX, kBlockY), 0, s>>>(a, t);
  return cudaGetLastError();
}

} // namespace lgcu
