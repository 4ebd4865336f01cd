// synth_scene.cpp — deterministic synthetic scenes for the SSVGI lighting/GI path (host only, no CUDA).
//
// The reference feeds its G-buffer pass from a rasteriser (src/Render/Renderers/SSVGIRenderer.h:107-158 drawing
// src/Scene/Scene.h objects). This module stands in for that rasteriser on a procedural scene so that the CUDA
// path, the CPU oracles and the benchmark all start from the same per-pixel fragments (lgcu_fragment, include/lgcu.h):
// a "city" of axis-aligned boxes on a ground plane with three enclosing walls, ray-cast from the
// camera for the fragments and from the spot light for the 1024² shadow map (the ShadowPass output, :63-104).
//
// All geometry is evaluated in double precision with +,-,*,/ and sqrt only and rounded once to float, so the output
// is bit-reproducible across machines (build with -ffp-contract=off).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/lgcu.h"
#include "legit_cuda/Camera.h"

namespace {

struct Rng { // splitmix64
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return double(next() >> 11) * (1.0 / 9007199254740992.0); } // [0,1)
  double uniform(double a, double b) { return a + (b - a) * uniform(); }
};

struct Box {
  double lo[3], hi[3];
  double normalScale;
};

struct Scene {
  std::vector<Box> boxes; // object id = index + kFirstBoxId
  std::vector<lgcu_draw_call_data> objects;
};

constexpr uint32_t kGroundId = 0, kWallBackId = 1, kWallLeftId = 2, kWallRightId = 3, kFirstBoxId = 4;
constexpr double kRoomHalf = 9.0, kWallHeight = 4.0;

void setIdentityScaled(lgcu_mat4 &m, float s) {
  std::memset(&m, 0, sizeof(m));
  m.m[0] = m.m[5] = m.m[10] = s;
  m.m[15] = 1.0f;
}

lgcu_draw_call_data makeObject(float scale, double r, double g, double b, double er, double eg, double eb) {
  lgcu_draw_call_data o;
  setIdentityScaled(o.modelMatrix, scale);
  o.albedoColor[0] = float(r); o.albedoColor[1] = float(g); o.albedoColor[2] = float(b); o.albedoColor[3] = 1.0f;
  o.emissiveColor[0] = float(er); o.emissiveColor[1] = float(eg); o.emissiveColor[2] = float(eb); o.emissiveColor[3] = 1.0f;
  return o;
}

Scene buildScene(uint64_t seed, uint32_t nBoxes) {
  Scene sc;
  Rng rng(seed * 0x2545F4914F6CDD1Dull + 0xC0FFEEull);
  sc.objects.push_back(makeObject(1.0f, 0.55, 0.55, 0.50, 0, 0, 0)); // ground
  sc.objects.push_back(makeObject(1.0f, 0.70, 0.35, 0.30, 0, 0, 0)); // back wall  z = +9
  sc.objects.push_back(makeObject(1.0f, 0.30, 0.65, 0.35, 0, 0, 0)); // left wall  x = -9
  sc.objects.push_back(makeObject(1.0f, 0.30, 0.40, 0.75, 0, 0, 0)); // right wall x = +9
  for (uint32_t i = 0; i < nBoxes; i++) {
    double cx, cz, sx, sz, hgt;
    // keep the camera (0, .5, -2) and the light axis clear so neither sits inside a box
    for (;;) {
      cx = rng.uniform(-8.0, 8.0);
      cz = rng.uniform(-8.0, 8.0);
      sx = rng.uniform(0.2, 1.5);
      sz = rng.uniform(0.2, 1.5);
      hgt = rng.uniform(0.2, 3.0);
      double dxc = std::fabs(cx - 0.0) - 0.5 * sx, dzc = std::fabs(cz + 2.0) - 0.5 * sz;
      if (dxc < 0.6 && dzc < 0.6) continue;
      break;
    }
    Box b;
    b.lo[0] = cx - 0.5 * sx; b.hi[0] = cx + 0.5 * sx;
    b.lo[1] = 0.0;           b.hi[1] = hgt;
    b.lo[2] = cz - 0.5 * sz; b.hi[2] = cz + 0.5 * sz;
    b.normalScale = rng.uniform(0.5, 2.0); // models a uniformly scaled modelMatrix: normals stay unnormalised
    double ar = rng.uniform(0.05, 1.0), ag = rng.uniform(0.05, 1.0), ab = rng.uniform(0.05, 1.0);
    double er = 0, eg = 0, eb = 0;
    if (rng.uniform() < 0.1) {
      er = rng.uniform(0.0, 8.0); eg = rng.uniform(0.0, 8.0); eb = rng.uniform(0.0, 8.0);
    }
    sc.boxes.push_back(b);
    sc.objects.push_back(makeObject(float(b.normalScale), ar, ag, ab, er, eg, eb));
  }
  return sc;
}

struct Hit {
  double t;
  uint32_t id;
  double n[3];
};

// nearest intersection of origin + t*dir, t > tMin
bool trace(const Scene &sc, const double o[3], const double d[3], double tMin, Hit &hit) {
  hit.t = 1e300;
  hit.id = LGCU_NO_OBJECT;
  auto consider = [&](double t, uint32_t id, double nx, double ny, double nz) {
    if (t > tMin && t < hit.t) {
      hit.t = t; hit.id = id; hit.n[0] = nx; hit.n[1] = ny; hit.n[2] = nz;
    }
  };
  if (d[1] != 0.0) { // ground y = 0, inside the room
    double t = -o[1] / d[1];
    double x = o[0] + t * d[0], z = o[2] + t * d[2];
    if (std::fabs(x) <= kRoomHalf && z <= kRoomHalf && z >= -kRoomHalf) consider(t, kGroundId, 0, 1, 0);
  }
  if (d[2] != 0.0) { // back wall z = +9 facing -z
    double t = (kRoomHalf - o[2]) / d[2];
    double x = o[0] + t * d[0], y = o[1] + t * d[1];
    if (std::fabs(x) <= kRoomHalf && y >= 0.0 && y <= kWallHeight) consider(t, kWallBackId, 0, 0, -1);
  }
  if (d[0] != 0.0) {
    double t = (-kRoomHalf - o[0]) / d[0]; // left wall facing +x
    double y = o[1] + t * d[1], z = o[2] + t * d[2];
    if (std::fabs(z) <= kRoomHalf && y >= 0.0 && y <= kWallHeight) consider(t, kWallLeftId, 1, 0, 0);
    t = (kRoomHalf - o[0]) / d[0]; // right wall facing -x
    y = o[1] + t * d[1]; z = o[2] + t * d[2];
    if (std::fabs(z) <= kRoomHalf && y >= 0.0 && y <= kWallHeight) consider(t, kWallRightId, -1, 0, 0);
  }
  for (size_t i = 0; i < sc.boxes.size(); i++) {
    const Box &b = sc.boxes[i];
    double t0 = -1e300, t1 = 1e300;
    int axis0 = -1;
    double sign0 = 0;
    bool miss = false;
    for (int a = 0; a < 3; a++) {
      if (d[a] == 0.0) {
        if (o[a] < b.lo[a] || o[a] > b.hi[a]) { miss = true; break; }
        continue;
      }
      double ta = (b.lo[a] - o[a]) / d[a], tb = (b.hi[a] - o[a]) / d[a];
      double sgn = -1.0; // entering through the lo face -> outward normal -axis
      if (ta > tb) { double tmp = ta; ta = tb; tb = tmp; sgn = 1.0; }
      if (ta > t0) { t0 = ta; axis0 = a; sign0 = sgn; }
      if (tb < t1) t1 = tb;
      if (t0 > t1) { miss = true; break; }
    }
    if (miss || axis0 < 0) continue;
    double n[3] = {0, 0, 0};
    n[axis0] = sign0 * b.normalScale;
    consider(t0, kFirstBoxId + uint32_t(i), n[0], n[1], n[2]);
  }
  return hit.id != LGCU_NO_OBJECT;
}

void mat4MulD(const double a[16], const double b[16], double out[16]) {
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += a[k * 4 + r] * b[c * 4 + k];
      out[c * 4 + r] = s;
    }
}

bool mat4InvD(const double m[16], double out[16]) { // Gauss-Jordan on column-major data
  double a[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      a[r][c] = m[c * 4 + r];
      a[r][c + 4] = r == c ? 1.0 : 0.0;
    }
  for (int i = 0; i < 4; i++) {
    int p = i;
    for (int r = i + 1; r < 4; r++)
      if (std::fabs(a[r][i]) > std::fabs(a[p][i])) p = r;
    if (a[p][i] == 0.0) return false;
    if (p != i)
      for (int c = 0; c < 8; c++) { double t = a[i][c]; a[i][c] = a[p][c]; a[p][c] = t; }
    double inv = 1.0 / a[i][i];
    for (int c = 0; c < 8; c++) a[i][c] *= inv;
    for (int r = 0; r < 4; r++)
      if (r != i) {
        double f = a[r][i];
        if (f != 0.0)
          for (int c = 0; c < 8; c++) a[r][c] -= f * a[i][c];
      }
  }
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[c * 4 + r] = a[r][c + 4];
  return true;
}

void mulPoint(const double m[16], double x, double y, double z, double w, double out[4]) {
  for (int r = 0; r < 4; r++) out[r] = m[0 * 4 + r] * x + m[1 * 4 + r] * y + m[2 * 4 + r] * z + m[3 * 4 + r] * w;
}

struct Projector {
  double viewProj[16], invViewProj[16], eye[3];
  bool init(const float view[16], const float proj[16]) {
    double v[16], p[16], invV[16];
    for (int i = 0; i < 16; i++) { v[i] = view[i]; p[i] = proj[i]; }
    mat4MulD(p, v, viewProj);
    if (!mat4InvD(viewProj, invViewProj) || !mat4InvD(v, invV)) return false;
    eye[0] = invV[12]; eye[1] = invV[13]; eye[2] = invV[14];
    return true;
  }
  // ray through the pixel centre, pointing at the far plane
  void ray(double u, double v, double dir[3]) const {
    double f[4];
    mulPoint(invViewProj, u * 2.0 - 1.0, v * 2.0 - 1.0, 1.0, 1.0, f);
    double d[3] = {f[0] / f[3] - eye[0], f[1] / f[3] - eye[1], f[2] / f[3] - eye[2]};
    double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    dir[0] = d[0] / len; dir[1] = d[1] / len; dir[2] = d[2] / len;
  }
  // depth-buffer value of a world point (NDC z; the Vulkan viewport maps it 1:1 to [0,1])
  double ndcZ(const double p[3]) const {
    double c[4];
    mulPoint(viewProj, p[0], p[1], p[2], 1.0, c);
    return c[2] / c[3];
  }
};

} // namespace

extern "C" {

// Number of draw calls (objects) a scene with nBoxes boxes has.
uint32_t lgs_object_count(uint32_t nBoxes) { return kFirstBoxId + nBoxes; }

// Per-draw-call constants of the scene (objects[i] belongs to lgcu_fragment.objectId == i).
int lgs_scene_objects(uint64_t seed, uint32_t nBoxes, lgcu_draw_call_data *objects, uint32_t maxObjects) {
  Scene sc = buildScene(seed, nBoxes);
  if (sc.objects.size() > maxObjects) return LGCU_ERR_INVALID_ARGUMENT;
  std::memcpy(objects, sc.objects.data(), sc.objects.size() * sizeof(lgcu_draw_call_data));
  return LGCU_OK;
}

// Rows [rowBegin, rowEnd) of the width x height fragment buffer as seen through (view, proj).
// `fragments` points at row 0 of the FULL buffer; fragmentPitchBytes >= width * 32.
int lgs_scene_fragments(uint64_t seed, uint32_t nBoxes, uint32_t width, uint32_t height, const float view[16],
                        const float proj[16], lgcu_fragment *fragments, uint64_t fragmentPitchBytes, uint32_t rowBegin,
                        uint32_t rowEnd) {
  Scene sc = buildScene(seed, nBoxes);
  Projector pr;
  if (!pr.init(view, proj) || rowEnd > height || rowBegin > rowEnd) return LGCU_ERR_INVALID_ARGUMENT;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t y = rowBegin; y < int64_t(rowEnd); y++) {
    lgcu_fragment *row = reinterpret_cast<lgcu_fragment *>(reinterpret_cast<uint8_t *>(fragments) + uint64_t(y) * fragmentPitchBytes);
    for (uint32_t x = 0; x < width; x++) {
      double dir[3];
      pr.ray((double(x) + 0.5) / double(width), (double(y) + 0.5) / double(height), dir);
      Hit hit;
      lgcu_fragment f;
      std::memset(&f, 0, sizeof(f));
      f.objectId = LGCU_NO_OBJECT;
      f.ndcDepth = 1.0f;
      if (trace(sc, pr.eye, dir, 1e-6, hit)) {
        double p[3] = {pr.eye[0] + hit.t * dir[0], pr.eye[1] + hit.t * dir[1], pr.eye[2] + hit.t * dir[2]};
        double z = pr.ndcZ(p);
        if (z >= 0.0 && z < 1.0) { // Vulkan clip volume 0 <= z <= w; z == 1 would fail the LESS depth test vs clear
          f.worldPos[0] = float(p[0]); f.worldPos[1] = float(p[1]); f.worldPos[2] = float(p[2]);
          f.worldNormal[0] = float(hit.n[0]); f.worldNormal[1] = float(hit.n[1]); f.worldNormal[2] = float(hit.n[2]);
          f.objectId = hit.id;
          f.ndcDepth = float(z);
          if (!(f.ndcDepth < 1.0f)) f.objectId = LGCU_NO_OBJECT, f.ndcDepth = 1.0f;
        }
      }
      row[x] = f;
    }
  }
  return LGCU_OK;
}

// size x size depth map of the scene seen from the light (ShadowPass stand-in), cleared to 1.0.
int lgs_scene_shadow_map(uint64_t seed, uint32_t nBoxes, uint32_t size, const float lightView[16],
                         const float lightProj[16], float *depth, uint64_t pitchBytes) {
  Scene sc = buildScene(seed, nBoxes);
  Projector pr;
  if (!pr.init(lightView, lightProj)) return LGCU_ERR_INVALID_ARGUMENT;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t y = 0; y < int64_t(size); y++) {
    float *row = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(depth) + uint64_t(y) * pitchBytes);
    for (uint32_t x = 0; x < size; x++) {
      double dir[3];
      pr.ray((double(x) + 0.5) / double(size), (double(y) + 0.5) / double(size), dir);
      Hit hit;
      float z = 1.0f;
      if (trace(sc, pr.eye, dir, 1e-6, hit)) {
        double p[3] = {pr.eye[0] + hit.t * dir[0], pr.eye[1] + hit.t * dir[1], pr.eye[2] + hit.t * dir[2]};
        double zz = pr.ndcZ(p);
        if (zz >= 0.0 && zz < 1.0) z = float(zz);
      }
      row[x] = z;
    }
  }
  return LGCU_OK;
}

// Frame matrices of SSVGIRenderer::RenderFrame (SSVGIRenderer.h:54-59) for a camera / light given as
// (position, vertAngle, horAngle) — src/Scene/Scene.h:20-34. Outputs are column-major float[16].
void lgs_frame_matrices(const float camPos[3], float camVertAngle, float camHorAngle, const float lightPos[3],
                        float lightVertAngle, float lightHorAngle, uint32_t width, uint32_t height, float *view,
                        float *proj, float *lightView, float *lightProj) {
  legit_cuda::Camera cam, light;
  cam.pos = legit_cuda::vec3{camPos[0], camPos[1], camPos[2]};
  cam.vertAngle = camVertAngle;
  cam.horAngle = camHorAngle;
  light.pos = legit_cuda::vec3{lightPos[0], lightPos[1], lightPos[2]};
  light.vertAngle = lightVertAngle;
  light.horAngle = lightHorAngle;
  legit_cuda::FrameMatrices f = legit_cuda::MakeFrameMatrices(cam, light, width, height);
  std::memcpy(view, f.viewMatrix.m, 64);
  std::memcpy(proj, f.projMatrix.m, 64);
  std::memcpy(lightView, f.lightViewMatrix.m, 64);
  std::memcpy(lightProj, f.lightProjMatrix.m, 64);
}

void lgs_mat4_inverse(const float *m, float *out) {
  lgcu_mat4 a;
  std::memcpy(a.m, m, 64);
  lgcu_mat4 r = lgcu_mat4_inverse(&a);
  std::memcpy(out, r.m, 64);
}

void lgs_mat4_mul(const float *a_, const float *b_, float *out) {
  lgcu_mat4 a, b;
  std::memcpy(a.m, a_, 64);
  std::memcpy(b.m, b_, 64);
  lgcu_mat4 r = lgcu_mat4_mul(&a, &b);
  std::memcpy(out, r.m, 64);
}

} // extern "C"

// ---- the same scene as an indexed triangle mesh (what the reference's Scene holds: src/Scene/Mesh.h:209-216 vertices, uint32
// indices, one draw per object) for the rasterisation front end (lgcu_raster_*). Draw i uses objects[i]; boxes are stored in
// object space (world / normalScale, unit normals) so that modelMatrix = scale(normalScale) reproduces the unnormalised world
// normals of the ray-cast fragments. Box bottoms (coplanar with the ground) are not emitted.
namespace {

struct MeshBuilder {
  std::vector<lgcu_vertex> vertices;
  std::vector<uint32_t> indices;
  std::vector<lgcu_draw> draws;
  void beginDraw(uint32_t objectId) {
    lgcu_draw d;
    std::memset(&d, 0, sizeof(d));
    d.firstIndex = uint32_t(indices.size());
    d.objectId = objectId;
    draws.push_back(d);
  }
  void endDraw() { draws.back().indexCount = uint32_t(indices.size()) - draws.back().firstIndex; }
  // quad p0,p1,p2,p3 (in order around the face) with normal n, positions divided by `scale`
  void quad(const double p[4][3], const double n[3], double scale) {
    const uint32_t base = uint32_t(vertices.size());
    for (int k = 0; k < 4; k++) {
      lgcu_vertex v;
      for (int c = 0; c < 3; c++) {
        v.pos[c] = float(p[k][c] / scale);
        v.normal[c] = float(n[c]);
      }
      v.uv[0] = (k == 1 || k == 2) ? 1.0f : 0.0f;
      v.uv[1] = (k >= 2) ? 1.0f : 0.0f;
      vertices.push_back(v);
    }
    const uint32_t idx[6] = {base, base + 1, base + 2, base, base + 2, base + 3};
    indices.insert(indices.end(), idx, idx + 6);
  }
};

MeshBuilder buildMesh(const Scene &sc) {
  MeshBuilder mb;
  const double R = kRoomHalf, Hh = kWallHeight;
  {
    const double p[4][3] = {{-R, 0, -R}, {R, 0, -R}, {R, 0, R}, {-R, 0, R}}, n[3] = {0, 1, 0};
    mb.beginDraw(kGroundId); mb.quad(p, n, 1.0); mb.endDraw();
  }
  {
    const double p[4][3] = {{-R, 0, R}, {R, 0, R}, {R, Hh, R}, {-R, Hh, R}}, n[3] = {0, 0, -1};
    mb.beginDraw(kWallBackId); mb.quad(p, n, 1.0); mb.endDraw();
  }
  {
    const double p[4][3] = {{-R, 0, -R}, {-R, 0, R}, {-R, Hh, R}, {-R, Hh, -R}}, n[3] = {1, 0, 0};
    mb.beginDraw(kWallLeftId); mb.quad(p, n, 1.0); mb.endDraw();
  }
  {
    const double p[4][3] = {{R, 0, -R}, {R, 0, R}, {R, Hh, R}, {R, Hh, -R}}, n[3] = {-1, 0, 0};
    mb.beginDraw(kWallRightId); mb.quad(p, n, 1.0); mb.endDraw();
  }
  for (size_t i = 0; i < sc.boxes.size(); i++) {
    const Box &b = sc.boxes[i];
    const double s = double(float(b.normalScale));
    const double *lo = b.lo, *hi = b.hi;
    mb.beginDraw(kFirstBoxId + uint32_t(i));
    {
      const double p[4][3] = {{lo[0], hi[1], lo[2]}, {hi[0], hi[1], lo[2]}, {hi[0], hi[1], hi[2]}, {lo[0], hi[1], hi[2]}}, n[3] = {0, 1, 0};
      mb.quad(p, n, s); // top
    }
    {
      const double p[4][3] = {{lo[0], lo[1], lo[2]}, {hi[0], lo[1], lo[2]}, {hi[0], hi[1], lo[2]}, {lo[0], hi[1], lo[2]}}, n[3] = {0, 0, -1};
      mb.quad(p, n, s); // z = lo
    }
    {
      const double p[4][3] = {{lo[0], lo[1], hi[2]}, {hi[0], lo[1], hi[2]}, {hi[0], hi[1], hi[2]}, {lo[0], hi[1], hi[2]}}, n[3] = {0, 0, 1};
      mb.quad(p, n, s); // z = hi
    }
    {
      const double p[4][3] = {{lo[0], lo[1], lo[2]}, {lo[0], lo[1], hi[2]}, {lo[0], hi[1], hi[2]}, {lo[0], hi[1], lo[2]}}, n[3] = {-1, 0, 0};
      mb.quad(p, n, s); // x = lo
    }
    {
      const double p[4][3] = {{hi[0], lo[1], lo[2]}, {hi[0], lo[1], hi[2]}, {hi[0], hi[1], hi[2]}, {hi[0], hi[1], lo[2]}}, n[3] = {1, 0, 0};
      mb.quad(p, n, s); // x = hi
    }
    mb.endDraw();
  }
  uint32_t tri = 0;
  for (lgcu_draw &d : mb.draws) {
    d.firstTriangle = tri;
    tri += d.indexCount / 3;
  }
  return mb;
}

} // namespace

extern "C" {

// sizes of the mesh form of a scene with nBoxes boxes
void lgs_scene_mesh_counts(uint32_t nBoxes, uint32_t *nVertices, uint32_t *nIndices, uint32_t *nDraws) {
  const uint32_t quads = 4 + 5 * nBoxes;
  *nVertices = 4 * quads;
  *nIndices = 6 * quads;
  *nDraws = kFirstBoxId + nBoxes;
}

int lgs_scene_mesh(uint64_t seed, uint32_t nBoxes, lgcu_vertex *vertices, uint32_t *indices, lgcu_draw *draws) {
  Scene sc = buildScene(seed, nBoxes);
  MeshBuilder mb = buildMesh(sc);
  std::memcpy(vertices, mb.vertices.data(), mb.vertices.size() * sizeof(lgcu_vertex));
  std::memcpy(indices, mb.indices.data(), mb.indices.size() * sizeof(uint32_t));
  std::memcpy(draws, mb.draws.data(), mb.draws.size() * sizeof(lgcu_draw));
  return LGCU_OK;
}

} // extern "C"
