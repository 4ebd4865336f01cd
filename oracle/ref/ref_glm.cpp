// ref_glm.cpp — TEST INFRASTRUCTURE (oracle/_ref reference arm; never linked into the product path).
//
// Frame matrices computed with the reference's vendored GLM (dependencies/glm, 0.9.9.2, default configuration:
// main.cpp defines GLM_DEPTH_ZERO_TO_ONE only after glm was already included, so it has no effect — SURVEY.md §8a),
// following the reference's own expressions. Used to pin legit_cuda's dependency-free matrix code bit for bit.
#include <glm/glm.hpp>
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/gtx/transform.hpp>

#include <cstring>

namespace {
// src/Scene/Scene.h:30-33 Camera::GetTransformMatrix
glm::mat4 cameraTransform(const float pos[3], float vertAngle, float horAngle) {
  return glm::translate(glm::vec3(pos[0], pos[1], pos[2])) * glm::rotate(horAngle, glm::vec3(0.0f, 1.0f, 0.0f)) *
         glm::rotate(vertAngle, glm::vec3(1.0f, 0.0f, 0.0f));
}
} // namespace

// src/Render/Renderers/SSVGIRenderer.h:54-59. Outputs are column-major float[16].
extern "C" void ref_frame_matrices(const float camPos[3], float camVertAngle, float camHorAngle, const float lightPos[3],
                                   float lightVertAngle, float lightHorAngle, unsigned width, unsigned height,
                                   float *view, float *proj, float *lightView, float *lightProj) {
  glm::mat4 v = glm::inverse(cameraTransform(camPos, camVertAngle, camHorAngle));
  glm::mat4 lv = glm::inverse(cameraTransform(lightPos, lightVertAngle, lightHorAngle));
  float aspect = float(width) / float(height);
  glm::mat4 p = glm::perspective(1.0f, aspect, 0.01f, 1000.0f) * glm::scale(glm::vec3(1.0f, -1.0f, -1.0f));
  glm::mat4 lp = glm::perspective(0.8f, 1.0f, 0.1f, 100.0f) * glm::scale(glm::vec3(1.0f, -1.0f, -1.0f));
  std::memcpy(view, &v[0][0], 64);
  std::memcpy(proj, &p[0][0], 64);
  std::memcpy(lightView, &lv[0][0], 64);
  std::memcpy(lightProj, &lp[0][0], 64);
}

// glm::inverse / operator* on raw column-major matrices, for known-answer tests of the hoisted per-frame constants.
extern "C" void ref_mat4_inverse(const float *m, float *out) {
  glm::mat4 a;
  std::memcpy(&a[0][0], m, 64);
  glm::mat4 r = glm::inverse(a);
  std::memcpy(out, &r[0][0], 64);
}
extern "C" void ref_mat4_mul(const float *a_, const float *b_, float *out) {
  glm::mat4 a, b;
  std::memcpy(&a[0][0], a_, 64);
  std::memcpy(&b[0][0], b_, 64);
  glm::mat4 r = a * b;
  std::memcpy(out, &r[0][0], 64);
}
