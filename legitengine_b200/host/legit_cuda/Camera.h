// Camera.h — camera / light transforms and the frame's projection matrices, dependency-free.
//
// Mirrors src/Scene/Scene.h:20-34 (struct Camera) and the matrix set-up at the top of
// SSVGIRenderer::RenderFrame (src/Render/Renderers/SSVGIRenderer.h:54-59). The reference uses glm 0.9.9.2 with its
// default configuration (right-handed, NDC z in [-1,1]: main.cpp:5 defines GLM_DEPTH_ZERO_TO_ONE too late and
// without the FORCE_ prefix); the functions below evaluate the same expressions in the same order
// so the matrices are bit-identical (pinned by tests/test_frame_math.py against the vendored glm).
#pragma once

#include <cmath>
#include <cstring>

#include "../../csrc/lgcu_mat4.h"

namespace legit_cuda {

struct vec3 {
  float x, y, z;
};

inline lgcu_mat4 Identity() {
  lgcu_mat4 m;
  std::memset(&m, 0, sizeof(m));
  m.m[0] = m.m[5] = m.m[10] = m.m[15] = 1.0f;
  return m;
}

// glm::translate(v) == translate(mat4(1), v): column 3 = m0*v.x + m1*v.y + m2*v.z + m3
inline lgcu_mat4 Translate(vec3 v) {
  lgcu_mat4 m = Identity(), r = m;
  for (int i = 0; i < 4; i++) r.m[12 + i] = ((m.m[0 + i] * v.x + m.m[4 + i] * v.y) + m.m[8 + i] * v.z) + m.m[12 + i];
  return r;
}

// glm::rotate(angle, axis) == rotate(mat4(1), angle, axis)
inline lgcu_mat4 Rotate(float angle, vec3 v) {
  const float c = std::cos(angle), s = std::sin(angle);
  const float invLen = 1.0f / std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);
  const float axis[3] = {v.x * invLen, v.y * invLen, v.z * invLen};
  const float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
  float rot[3][3];
  rot[0][0] = c + temp[0] * axis[0];
  rot[0][1] = temp[0] * axis[1] + s * axis[2];
  rot[0][2] = temp[0] * axis[2] - s * axis[1];
  rot[1][0] = temp[1] * axis[0] - s * axis[2];
  rot[1][1] = c + temp[1] * axis[1];
  rot[1][2] = temp[1] * axis[2] + s * axis[0];
  rot[2][0] = temp[2] * axis[0] + s * axis[1];
  rot[2][1] = temp[2] * axis[1] - s * axis[0];
  rot[2][2] = c + temp[2] * axis[2];
  lgcu_mat4 m = Identity(), r;
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 4; i++) r.m[j * 4 + i] = (m.m[0 + i] * rot[j][0] + m.m[4 + i] * rot[j][1]) + m.m[8 + i] * rot[j][2];
  for (int i = 0; i < 4; i++) r.m[12 + i] = m.m[12 + i];
  return r;
}

inline lgcu_mat4 Scale(vec3 v) {
  // whole columns are scaled (so a negative factor yields -0 off the diagonal, as in glm)
  lgcu_mat4 r = Identity();
  const float s[3] = {v.x, v.y, v.z};
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 4; i++) r.m[j * 4 + i] *= s[j];
  return r;
}

// glm::perspective with the default clip control (RH, depth -1..1)
inline lgcu_mat4 Perspective(float fovy, float aspect, float zNear, float zFar) {
  const float tanHalfFovy = std::tan(fovy / 2.0f);
  lgcu_mat4 r;
  std::memset(&r, 0, sizeof(r));
  r.m[0 * 4 + 0] = 1.0f / (aspect * tanHalfFovy);
  r.m[1 * 4 + 1] = 1.0f / (tanHalfFovy);
  r.m[2 * 4 + 2] = -(zFar + zNear) / (zFar - zNear);
  r.m[2 * 4 + 3] = -1.0f;
  r.m[3 * 4 + 2] = -(2.0f * zFar * zNear) / (zFar - zNear);
  return r;
}

// src/Scene/Scene.h:20-34
struct Camera {
  vec3 pos = {0.0f, 0.0f, 0.0f};
  float vertAngle = 0.0f, horAngle = 0.0f;
  lgcu_mat4 GetTransformMatrix() const {
    lgcu_mat4 t = Translate(pos), ry = Rotate(horAngle, vec3{0.0f, 1.0f, 0.0f}), rx = Rotate(vertAngle, vec3{1.0f, 0.0f, 0.0f});
    lgcu_mat4 tr = lgcu_mat4_mul(&t, &ry);
    return lgcu_mat4_mul(&tr, &rx);
  }
};

// The defaults of the reference application (src/main.cpp:166-172).
inline Camera DefaultCamera() {
  Camera c;
  c.pos = vec3{0.0f, 0.5f, -2.0f};
  return c;
}
inline Camera DefaultLight() {
  Camera l;
  l.pos = vec3{0.0f, 5.0f, 0.0f};
  l.vertAngle = 3.1415f / 2.0f;
  return l;
}

struct FrameMatrices {
  lgcu_mat4 viewMatrix, projMatrix, lightViewMatrix, lightProjMatrix;
};

// SSVGIRenderer.h:54-59
inline FrameMatrices MakeFrameMatrices(const Camera &camera, const Camera &light, unsigned width, unsigned height) {
  FrameMatrices f;
  lgcu_mat4 ct = camera.GetTransformMatrix(), lt = light.GetTransformMatrix();
  f.viewMatrix = lgcu_mat4_inverse(&ct);
  f.lightViewMatrix = lgcu_mat4_inverse(&lt);
  const float aspect = float(width) / float(height);
  lgcu_mat4 flip = Scale(vec3{1.0f, -1.0f, -1.0f});
  lgcu_mat4 p = Perspective(1.0f, aspect, 0.01f, 1000.0f), lp = Perspective(0.8f, 1.0f, 0.1f, 100.0f);
  f.projMatrix = lgcu_mat4_mul(&p, &flip);
  f.lightProjMatrix = lgcu_mat4_mul(&lp, &flip);
  return f;
}

} // namespace legit_cuda
