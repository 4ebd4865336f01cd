/*
 * lgcu_mat4.h — host-side 4x4 float matrix helpers for the per-frame constants of the SSVGI passes.
 *
 * The reference shaders recompute inverse(projMatrix * viewMatrix), inverse(viewMatrix) ... in EVERY fragment
 * (SH/Common/directLighting.frag:51-55, SH/SSVGI/indirectLighting.frag:123-126, SH/Common/gBufferBuilder.frag:30).
 * They are frame constants, so the C ABI hoists them to the host once per call. To keep results identical to a
 * shader that evaluates them per fragment in fp32, the operation order below is the cofactor / column-combination
 * order GLSL compilers and glm (the reference's host maths, dependencies/glm 0.9.9.2) use:
 *   m * n     : col_j = ((m0*n_j0 + m1*n_j1) + m2*n_j2) + m3*n_j3
 *   m * v     : (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
 *   inverse(m): 2x2 sub-determinant cofactors, sign-alternated, scaled by 1/det with det = (a+b)+(c+d).
 * Compile without FMA contraction (-ffp-contract=off / nvcc host code never contracts) so the order is honoured.
 * Column-major storage: m[c*4 + r], like glm::mat4 and lgcu_mat4.
 */
#ifndef LGCU_MAT4_H
#define LGCU_MAT4_H

#include "../../include/lgcu.h"

#ifdef __cplusplus
extern "C" {
#endif

static inline lgcu_mat4 lgcu_mat4_mul(const lgcu_mat4 *a, const lgcu_mat4 *b) {
  lgcu_mat4 r;
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++)
      r.m[j * 4 + i] = ((a->m[0 + i] * b->m[j * 4 + 0] + a->m[4 + i] * b->m[j * 4 + 1]) + a->m[8 + i] * b->m[j * 4 + 2]) +
                       a->m[12 + i] * b->m[j * 4 + 3];
  return r;
}

static inline void lgcu_mat4_mul_vec4(const lgcu_mat4 *a, const float v[4], float out[4]) {
  for (int i = 0; i < 4; i++)
    out[i] = (a->m[0 + i] * v[0] + a->m[4 + i] * v[1]) + (a->m[8 + i] * v[2] + a->m[12 + i] * v[3]);
}

#define LGCU_M(c, r) (m->m[(c) * 4 + (r)])
static inline lgcu_mat4 lgcu_mat4_inverse(const lgcu_mat4 *m) {
  const float c00 = LGCU_M(2, 2) * LGCU_M(3, 3) - LGCU_M(3, 2) * LGCU_M(2, 3);
  const float c02 = LGCU_M(1, 2) * LGCU_M(3, 3) - LGCU_M(3, 2) * LGCU_M(1, 3);
  const float c03 = LGCU_M(1, 2) * LGCU_M(2, 3) - LGCU_M(2, 2) * LGCU_M(1, 3);
  const float c04 = LGCU_M(2, 1) * LGCU_M(3, 3) - LGCU_M(3, 1) * LGCU_M(2, 3);
  const float c06 = LGCU_M(1, 1) * LGCU_M(3, 3) - LGCU_M(3, 1) * LGCU_M(1, 3);
  const float c07 = LGCU_M(1, 1) * LGCU_M(2, 3) - LGCU_M(2, 1) * LGCU_M(1, 3);
  const float c08 = LGCU_M(2, 1) * LGCU_M(3, 2) - LGCU_M(3, 1) * LGCU_M(2, 2);
  const float c10 = LGCU_M(1, 1) * LGCU_M(3, 2) - LGCU_M(3, 1) * LGCU_M(1, 2);
  const float c11 = LGCU_M(1, 1) * LGCU_M(2, 2) - LGCU_M(2, 1) * LGCU_M(1, 2);
  const float c12 = LGCU_M(2, 0) * LGCU_M(3, 3) - LGCU_M(3, 0) * LGCU_M(2, 3);
  const float c14 = LGCU_M(1, 0) * LGCU_M(3, 3) - LGCU_M(3, 0) * LGCU_M(1, 3);
  const float c15 = LGCU_M(1, 0) * LGCU_M(2, 3) - LGCU_M(2, 0) * LGCU_M(1, 3);
  const float c16 = LGCU_M(2, 0) * LGCU_M(3, 2) - LGCU_M(3, 0) * LGCU_M(2, 2);
  const float c18 = LGCU_M(1, 0) * LGCU_M(3, 2) - LGCU_M(3, 0) * LGCU_M(1, 2);
  const float c19 = LGCU_M(1, 0) * LGCU_M(2, 2) - LGCU_M(2, 0) * LGCU_M(1, 2);
  const float c20 = LGCU_M(2, 0) * LGCU_M(3, 1) - LGCU_M(3, 0) * LGCU_M(2, 1);
  const float c22 = LGCU_M(1, 0) * LGCU_M(3, 1) - LGCU_M(3, 0) * LGCU_M(1, 1);
  const float c23 = LGCU_M(1, 0) * LGCU_M(2, 1) - LGCU_M(2, 0) * LGCU_M(1, 1);

  const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {LGCU_M(1, 0), LGCU_M(0, 0), LGCU_M(0, 0), LGCU_M(0, 0)};
  const float v1[4] = {LGCU_M(1, 1), LGCU_M(0, 1), LGCU_M(0, 1), LGCU_M(0, 1)};
  const float v2[4] = {LGCU_M(1, 2), LGCU_M(0, 2), LGCU_M(0, 2), LGCU_M(0, 2)};
  const float v3[4] = {LGCU_M(1, 3), LGCU_M(0, 3), LGCU_M(0, 3), LGCU_M(0, 3)};
  static const float signA[4] = {+1.0f, -1.0f, +1.0f, -1.0f}, signB[4] = {-1.0f, +1.0f, -1.0f, +1.0f};

  lgcu_mat4 inv;
  for (int i = 0; i < 4; i++) {
    inv.m[0 * 4 + i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * signA[i];
    inv.m[1 * 4 + i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * signB[i];
    inv.m[2 * 4 + i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * signA[i];
    inv.m[3 * 4 + i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * signB[i];
  }
  const float d0 = LGCU_M(0, 0) * inv.m[0], d1 = LGCU_M(0, 1) * inv.m[4], d2 = LGCU_M(0, 2) * inv.m[8],
              d3 = LGCU_M(0, 3) * inv.m[12];
  const float det = (d0 + d1) + (d2 + d3);
  const float oneOverDet = 1.0f / det;
  for (int i = 0; i < 16; i++) inv.m[i] = inv.m[i] * oneOverDet;
  return inv;
}
#undef LGCU_M

#ifdef __cplusplus
}
#endif
#endif
