/*
 * raster_oracle.c — TEST INFRASTRUCTURE. CPU restatement of the rasterisation front end of LegitEngine's SSVGI frame:
 * "ShadowPass" (src/Render/Renderers/SSVGIRenderer.h:63-104) and the raster half of "GBufferPass" (:107-158).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product never does.
 *
 * What is restated, and from where:
 *  - vertex stage: SH/Common/gBufferBuilder.vert:32-40 and SH/Common/shadowmapBuilder.vert:32-40 (identical maths):
 *      vertWorldPos = (modelMatrix * vec4(attribPosition, 1)).xyz, vertWorldNormal = (modelMatrix * vec4(attribNormal, 0)).xyz,
 *      gl_Position = projMatrix * viewMatrix * vec4(vertWorldPos, 1)   [left-associative: (proj * view) * v]
 *    in fp32 with glm's evaluation order (mat*mat columns ((a+b)+c)+d, mat*vec (a+b)+(c+d)). PINNED bit-for-bit against the
 *    reference's shipped gBufferBuilder.vert.spv / shadowmapBuilder.vert.spv run through oracle/_ref (tests/test_raster_cpu.py).
 *  - draw loop: Scene::IterateObjects order, drawIndexed(indicesCount, 1, 0, 0, 0) per object, triangle list (:84-102, :138-156).
 *  - fixed-function state: fill, cull none, depth test LESS with write, depth clamp off, 1 sample, viewport = render area with
 *    depth range [0,1] (LV/Pipeline.h:6-12, 148-178; LV/RenderPassCache.h:94-101); attachments cleared first (depth 1.0).
 *  - the rasteriser itself is fixed-function hardware in the reference — there is no reference source to follow, so "parity
 *    unpinned by the reference" holds for coverage at triangle edges. This file states Vulkan's rules (pixel-centre sampling,
 *    top-left rule, perspective-correct attribute interpolation, z interpolated as z_clip / w_clip, clip volume 0 <= z <= w)
 *    in ONE fixed arithmetic — "rule R" of DESIGN.md §8 — which the CUDA kernels (csrc/k_raster.cu) follow operation for
 *    operation, so that the two agree bit for bit. Rule R is validated independently against the analytic ray caster of the
 *    synthetic scene (legitengine_b200/host/synth_scene.cpp) in tests/test_raster_cpu.py.
 *
 * Rule R (all in IEEE double unless stated; no FMA contraction: build with -ffp-contract=off):
 *   per vertex k:   X_k = (x_clip + w_clip) * (0.5 * width),  Y_k = (y_clip + w_clip) * (0.5 * height),  Z_k = z_clip,  W_k = w_clip
 *   per edge i (opposite vertex i; j = i+1, k = i+2 mod 3):   a_i = Y_j*W_k - W_j*Y_k,  b_i = W_j*X_k - X_j*W_k,  c_i = X_j*Y_k - Y_j*X_k
 *   det = (a_0*X_0 + b_0*Y_0) + c_0*W_0;  det == 0 or non-finite -> triangle culled;  det < 0 -> negate every a, b, c
 *   per pixel (x, y): px = x + 0.5, py = y + 0.5,  e_i = (a_i*px + b_i*py) + c_i
 *     covered  <=>  for all i: e_i > 0, or e_i == 0 and (a_i > 0 or (a_i == 0 and b_i > 0))          [top-left rule]
 *     zn = (e_0*Z_0 + e_1*Z_1) + e_2*Z_2,  wn = (e_0*W_0 + e_1*W_1) + e_2*W_2;  clipped unless 0 <= zn <= wn and wn > 0
 *     depth = (float)(zn / wn), -0 -> +0;  depth test: depth < current (LESS), equal depth keeps the earlier triangle
 *     barycentrics  s = (e_0 + e_1) + e_2,  l_i = (float)(e_i / s);  attribute = (l_0*A_0 + l_1*A_1) + l_2*A_2 in fp32
 */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lgcu.h"

typedef struct {
  double a[3], b[3], c[3], Z[3], W[3];
  float wp[3][3], wn[3][3];
  uint32_t objectId;
  int culled, x0, y0, x1, y1; /* inclusive pixel bounding box */
} tri_setup;

static void m4_mul(const float *a, const float *b, float *r) { /* glm mat4 * mat4 */
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++)
      r[j * 4 + i] = ((a[i] * b[j * 4] + a[4 + i] * b[j * 4 + 1]) + a[8 + i] * b[j * 4 + 2]) + a[12 + i] * b[j * 4 + 3];
}
static void m4_mul_v4(const float *m, float x, float y, float z, float w, float *r) { /* glm mat4 * vec4 */
  for (int i = 0; i < 4; i++) r[i] = (m[i] * x + m[4 + i] * y) + (m[8 + i] * z + m[12 + i] * w);
}

/* vertex stage of one vertex: world position / normal and clip position (gBufferBuilder.vert:34-36) */
static void vertex_stage(const float *model, const float *viewProj, const lgcu_vertex *v, float wp[3], float wn[3], float clip[4]) {
  float t[4];
  m4_mul_v4(model, v->pos[0], v->pos[1], v->pos[2], 1.0f, t);
  wp[0] = t[0]; wp[1] = t[1]; wp[2] = t[2];
  m4_mul_v4(model, v->normal[0], v->normal[1], v->normal[2], 0.0f, t);
  wn[0] = t[0]; wn[1] = t[1]; wn[2] = t[2];
  m4_mul_v4(viewProj, wp[0], wp[1], wp[2], 1.0f, clip);
}

static int clampi(double v, int lo, int hi) { return v < (double)lo ? lo : (v > (double)hi ? hi : (int)v); }

static void setup_triangle(tri_setup *t, const float clip[3][4], int width, int height, int rowBegin, int rowEnd) {
  double X[3], Y[3];
  const double hw = 0.5 * (double)width, hh = 0.5 * (double)height;
  for (int k = 0; k < 3; k++) {
    X[k] = ((double)clip[k][0] + (double)clip[k][3]) * hw;
    Y[k] = ((double)clip[k][1] + (double)clip[k][3]) * hh;
    t->Z[k] = (double)clip[k][2];
    t->W[k] = (double)clip[k][3];
  }
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    t->a[i] = Y[j] * t->W[k] - t->W[j] * Y[k];
    t->b[i] = t->W[j] * X[k] - X[j] * t->W[k];
    t->c[i] = X[j] * Y[k] - Y[j] * X[k];
  }
  const double det = (t->a[0] * X[0] + t->b[0] * Y[0]) + t->c[0] * t->W[0];
  t->culled = !(det != 0.0) || !isfinite(det);
  if (t->culled) return;
  if (det < 0.0)
    for (int i = 0; i < 3; i++) { t->a[i] = -t->a[i]; t->b[i] = -t->b[i]; t->c[i] = -t->c[i]; }
  /* conservative bounding box: the projected vertices when all are in front of the eye plane, else the whole target */
  t->x0 = 0; t->y0 = rowBegin; t->x1 = width - 1; t->y1 = rowEnd - 1;
  if (t->W[0] > 0.0 && t->W[1] > 0.0 && t->W[2] > 0.0) {
    double minx = 1e300, maxx = -1e300, miny = 1e300, maxy = -1e300;
    for (int k = 0; k < 3; k++) {
      const double px = X[k] / t->W[k], py = Y[k] / t->W[k];
      if (px < minx) minx = px;
      if (px > maxx) maxx = px;
      if (py < miny) miny = py;
      if (py > maxy) maxy = py;
    }
    const int bx0 = clampi(floor(minx - 1.0), 0, width), bx1 = clampi(ceil(maxx + 1.0), -1, width - 1);
    const int by0 = clampi(floor(miny - 1.0), 0, height), by1 = clampi(ceil(maxy + 1.0), -1, height - 1);
    if (bx0 > t->x0) t->x0 = bx0;
    if (bx1 < t->x1) t->x1 = bx1;
    if (by0 > t->y0) t->y0 = by0;
    if (by1 < t->y1) t->y1 = by1;
  }
}

/* coverage + depth of one pixel; returns 0 if not covered / clipped */
static inline int shade_pixel(const tri_setup *t, int x, int y, double e[3], float *depth) {
  const double px = (double)x + 0.5, py = (double)y + 0.5;
  for (int i = 0; i < 3; i++) {
    e[i] = (t->a[i] * px + t->b[i] * py) + t->c[i];
    if (!(e[i] > 0.0 || (e[i] == 0.0 && (t->a[i] > 0.0 || (t->a[i] == 0.0 && t->b[i] > 0.0))))) return 0;
  }
  const double zn = (e[0] * t->Z[0] + e[1] * t->Z[1]) + e[2] * t->Z[2];
  const double wn = (e[0] * t->W[0] + e[1] * t->W[1]) + e[2] * t->W[2];
  if (!(wn > 0.0) || !(zn >= 0.0) || !(zn <= wn)) return 0;
  float d = (float)(zn / wn);
  if (d <= 0.0f) d = 0.0f;
  *depth = d;
  return 1;
}

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* Rasterises the scene into vis[y*width + x] = (depth bits << 32) | triangle id, ~0 = uncovered; setups[] receives the triangle
 * records. Scene arrays are HOST memory here. */
static int rasterise(const lgcu_mesh_scene *sc, const float *view, const float *proj, int width, int height, int rowBegin, int rowEnd,
                     uint64_t *vis, tri_setup *setups) {
  float viewProj[16];
  m4_mul(proj, view, viewProj);
  for (size_t i = 0; i < (size_t)width * height; i++) vis[i] = ~0ull;
  uint32_t prim = 0;
  for (uint32_t d = 0; d < sc->nDraws; d++) {
    const lgcu_draw *dr = &sc->draws[d];
    if (dr->objectId >= sc->nObjects || dr->firstIndex + dr->indexCount > sc->nIndices) return LGCU_ERR_INVALID_ARGUMENT;
    const float *model = sc->objects[dr->objectId].modelMatrix.m;
    for (uint32_t tI = 0; tI < dr->indexCount / 3; tI++, prim++) {
      tri_setup *t = &setups[prim];
      float clip[3][4];
      for (int k = 0; k < 3; k++) {
        const uint32_t vi = sc->indices[dr->firstIndex + 3 * tI + k] + dr->vertexOffset;
        if (vi >= sc->nVertices) return LGCU_ERR_INVALID_ARGUMENT;
        vertex_stage(model, viewProj, &sc->vertices[vi], t->wp[k], t->wn[k], clip[k]);
      }
      t->objectId = dr->objectId;
      setup_triangle(t, clip, width, height, rowBegin, rowEnd);
    }
  }
  const uint32_t nTri = prim;
  /* depth test LESS against the cleared 1.0, in draw order; rows are independent */
#pragma omp parallel for schedule(dynamic, 8)
  for (int y = rowBegin; y < rowEnd; y++) {
    for (uint32_t p = 0; p < nTri; p++) {
      const tri_setup *t = &setups[p];
      if (t->culled || y < t->y0 || y > t->y1) continue;
      for (int x = t->x0; x <= t->x1; x++) {
        double e[3];
        float depth;
        if (!shade_pixel(t, x, y, e, &depth)) continue;
        if (!(depth < 1.0f)) continue;
        const uint64_t key = ((uint64_t)f2u(depth) << 32) | p;
        if (key < vis[(size_t)y * width + x]) vis[(size_t)y * width + x] = key;
      }
    }
  }
  return LGCU_OK;
}

/* K0: shadow map (D32F host image, `pitchBytes` per row), cleared to 1.0 */
int orc_raster_shadow_map(const lgcu_shadowmap_builder_data *params, const lgcu_mesh_scene *scene, uint32_t size, float *depth,
                          uint64_t pitchBytes) {
  if (!params || !scene || !depth) return LGCU_ERR_INVALID_ARGUMENT;
  uint64_t *vis = (uint64_t *)malloc((size_t)size * size * 8);
  tri_setup *setups = (tri_setup *)malloc(sizeof(tri_setup) * (scene->nTriangles ? scene->nTriangles : 1));
  int st = rasterise(scene, params->lightViewMatrix.m, params->lightProjMatrix.m, (int)size, (int)size, 0, (int)size, vis, setups);
  if (st == LGCU_OK)
    for (uint32_t y = 0; y < size; y++) {
      float *row = (float *)((unsigned char *)depth + (uint64_t)y * pitchBytes);
      for (uint32_t x = 0; x < size; x++) {
        const uint64_t key = vis[(size_t)y * size + x];
        row[x] = key == ~0ull ? 1.0f : u2f((uint32_t)(key >> 32));
      }
    }
  free(vis);
  free(setups);
  return st;
}

/* raster half of K1: the lgcu_fragment buffer (host), rows [rowBegin,rowEnd) */
int orc_raster_gbuffer(const lgcu_gbuffer_builder_data *params, const lgcu_mesh_scene *scene, uint32_t width, uint32_t height,
                       lgcu_fragment *fragments, uint64_t fragmentPitchBytes, const lgcu_rows *rows) {
  if (!params || !scene || !fragments) return LGCU_ERR_INVALID_ARGUMENT;
  const int rowBegin = rows ? (int)rows->y0 : 0, rowEnd = rows ? (int)(rows->y1 < height ? rows->y1 : height) : (int)height;
  uint64_t *vis = (uint64_t *)malloc((size_t)width * height * 8);
  tri_setup *setups = (tri_setup *)malloc(sizeof(tri_setup) * (scene->nTriangles ? scene->nTriangles : 1));
  int st = rasterise(scene, params->viewMatrix.m, params->projMatrix.m, (int)width, (int)height, rowBegin, rowEnd, vis, setups);
  if (st == LGCU_OK) {
#pragma omp parallel for schedule(static)
    for (int y = rowBegin; y < rowEnd; y++) {
      lgcu_fragment *row = (lgcu_fragment *)((unsigned char *)fragments + (uint64_t)y * fragmentPitchBytes);
      for (uint32_t x = 0; x < width; x++) {
        const uint64_t key = vis[(size_t)y * width + x];
        lgcu_fragment f;
        memset(&f, 0, sizeof(f));
        f.objectId = LGCU_NO_OBJECT;
        f.ndcDepth = 1.0f;
        if (key != ~0ull) {
          const tri_setup *t = &setups[(uint32_t)key];
          double e[3];
          float depth;
          shade_pixel(t, (int)x, y, e, &depth);
          const double s = (e[0] + e[1]) + e[2];
          const float l0 = (float)(e[0] / s), l1 = (float)(e[1] / s), l2 = (float)(e[2] / s);
          for (int c = 0; c < 3; c++) {
            f.worldPos[c] = (l0 * t->wp[0][c] + l1 * t->wp[1][c]) + l2 * t->wp[2][c];
            f.worldNormal[c] = (l0 * t->wn[0][c] + l1 * t->wn[1][c]) + l2 * t->wn[2][c];
          }
          f.objectId = t->objectId;
          f.ndcDepth = u2f((uint32_t)(key >> 32));
        }
        row[x] = f;
      }
    }
  }
  free(vis);
  free(setups);
  return st;
}

/* vertex stage alone, for the pin against the reference's vertex SPIR-V: out = n x {worldPos[3], worldNormal[3], clip[4]} */
int orc_vertex_stage(const float *model, const float *view, const float *proj, const lgcu_vertex *vertices, uint32_t n, float *out) {
  float viewProj[16];
  m4_mul(proj, view, viewProj);
  for (uint32_t i = 0; i < n; i++) vertex_stage(model, viewProj, &vertices[i], out + 10 * i, out + 10 * i + 3, out + 10 * i + 6);
  return LGCU_OK;
}
