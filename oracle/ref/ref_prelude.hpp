// ref_prelude.hpp — TEST INFRASTRUCTURE (oracle/_ref reference arm; never linked into the product path).
//
// Glue that lets the C++ emitted by the reference's vendored SPIRV-Cross (dependencies/spirv-cross) from the
// reference's shipped SPIR-V run on the CPU. Everything arithmetic comes from the generated code + the vendored
// GLM; the only thing restated here is the TEXTURE UNIT, which the vendored CPU runtime stubs out
// (dependencies/spirv-cross/include/spirv_cross/sampler.hpp:46-67 returns constants for linear filtering, has
// no texelFetch / textureLod / shadow compare, and does not even instantiate: its sampleLod calls a
// three-argument sample() that does not exist). That one header is therefore masked through its include guard and
// the sampler types are declared here. Their behaviour follows SURVEY.md Appendix B and is implemented
// once, in oracle/texel_codec.h, shared with the plain-C restatement so both oracles use one texture unit.
//
// Each pass is its own translation unit; the vendored runtime defines non-inline extern "C" functions in a
// header, so they are renamed per pass with the macros below to keep one shared object.
#pragma once

#ifndef REF_PASS_NAME
#error "define REF_PASS_NAME"
#endif
#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define spirv_cross_set_stage_input REF_CAT(REF_PASS_NAME, _scx_set_stage_input)
#define spirv_cross_set_stage_output REF_CAT(REF_PASS_NAME, _scx_set_stage_output)
#define spirv_cross_set_push_constant REF_CAT(REF_PASS_NAME, _scx_set_push_constant)
#define spirv_cross_set_uniform_constant REF_CAT(REF_PASS_NAME, _scx_set_uniform_constant)
#define spirv_cross_set_resource REF_CAT(REF_PASS_NAME, _scx_set_resource)
#define spirv_cross_set_builtin REF_CAT(REF_PASS_NAME, _scx_set_builtin)
#define spirv_cross_get_interface REF_CAT(REF_PASS_NAME, _scx_get_interface)
#define spirv_cross_construct REF_CAT(REF_PASS_NAME, _scx_construct)
#define spirv_cross_destruct REF_CAT(REF_PASS_NAME, _scx_destruct)
#define spirv_cross_invoke REF_CAT(REF_PASS_NAME, _scx_invoke)
#define spirv_cross_create_sampler_2d REF_CAT(REF_PASS_NAME, _scx_create_sampler_2d)
#define spirv_cross_destroy_sampler_2d REF_CAT(REF_PASS_NAME, _scx_destroy_sampler_2d)
// Every generated module declares `namespace Impl { struct Shader ... }`; give each pass its own namespace so the
// seven translation units do not violate the one-definition rule when linked into one shared object.
#define Impl REF_CAT(REF_PASS_NAME, _Impl)

#ifndef GLM_FORCE_SWIZZLE
#define GLM_FORCE_SWIZZLE
#endif
#ifndef GLM_FORCE_RADIANS
#define GLM_FORCE_RADIANS
#endif
#include <glm/glm.hpp> // vendored glm 0.9.9.2, same force-macros as internal_interface.hpp sets

#include "../texel_codec.h"

#define SPIRV_CROSS_SAMPLER_HPP // mask the vendored (non-instantiable) sampler stub
namespace spirv_cross {
struct sampler2D {
  virtual ~sampler2D() {}
};
} // namespace spirv_cross

#include "spirv_cross/internal_interface.hpp" // vendored, unmodified

namespace spirv_cross {

// A bound (image view, sampler) pair. Filtering state is implied by the call the shader makes:
// the hot path uses clamp-to-edge everywhere, linear min/mag + linear mip for textureLod/texture, and
// texelFetch for the nearest-sampler passes (SSVGIRenderer.h:17-18, MipBuilder.h:133, BlurBuilder.h:10).
struct RefSampler2D : sampler2D {
  explicit RefSampler2D(const lgcu_image *image) : img(image) {}
  const lgcu_image *img;
};

struct sampler2DShadow {
  const lgcu_image *img;
};

inline glm::vec4 texelFetch(sampler2D &s, const glm::ivec2 &p, int lod) {
  float t[4];
  orc_load_texel(static_cast<RefSampler2D &>(s).img, (uint32_t)lod, p.x, p.y, t);
  return glm::vec4(t[0], t[1], t[2], t[3]);
}

inline glm::vec4 textureLod(sampler2D &s, const glm::vec2 &uv, float lod) {
  float t[4];
  orc_texture_lod(static_cast<RefSampler2D &>(s).img, uv.x, uv.y, lod, t);
  return glm::vec4(t[0], t[1], t[2], t[3]);
}

// Implicit-LOD sample on a full-screen 1:1 pass over a single-level image == lod 0.
inline glm::vec4 texture(sampler2D &s, const glm::vec2 &uv) { return textureLod(s, uv, 0.0f); }

inline float texture(sampler2DShadow &s, const glm::vec3 &c) { return orc_texture_shadow(s.img, c.x, c.y, c.z); }

} // namespace spirv_cross

namespace { // per-TU: the bodies below call the per-pass renamed runtime functions

// One shader instance (per OpenMP thread) of the generated module.
struct RefShaderInstance {
  const spirv_cross_interface *iface;
  spirv_cross_shader_t *sh;
  glm::vec4 fragCoord;
  glm::vec2 screenCoord;
  RefShaderInstance() : iface(spirv_cross_get_interface()), sh(iface->construct()) {
    spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_FRAG_COORD, &fragCoord, sizeof(fragCoord));
  }
  ~RefShaderInstance() { iface->destruct(sh); }
  void resource(unsigned set, unsigned binding, void *object) {
    void *p = object;
    spirv_cross_set_resource(sh, set, binding, &p, sizeof(p));
  }
  void input(unsigned location, void *data, size_t size) { spirv_cross_set_stage_input(sh, location, data, size); }
  void output(unsigned location, void *data, size_t size) { spirv_cross_set_stage_output(sh, location, data, size); }
  void bindScreenCoord(unsigned location) { input(location, &screenCoord, sizeof(screenCoord)); }
  // gl_FragCoord = (x + .5, y + .5, .5, 1) (SH/Common/screenspaceQuad.vert:13-18 draws z = 0.5, w = 1);
  // fragScreenCoord is the interpolated quad coordinate = pixel centre / render-area size.
  void setPixel(int x, int y, int w, int h) {
    fragCoord = glm::vec4(float(x) + 0.5f, float(y) + 0.5f, 0.5f, 1.0f);
    screenCoord = glm::vec2((float(x) + 0.5f) / float(w), (float(y) + 0.5f) / float(h));
  }
  void invoke() { iface->invoke(sh); }
};

} // namespace

static inline void ref_row_range(const lgcu_rows *rows, uint32_t level, int h, int *y0, int *y1) {
  if (!rows) {
    *y0 = 0;
    *y1 = h;
    return;
  }
  uint32_t a = rows->y0 >> level;
  uint32_t b = (rows->y1 + ((1u << level) - 1u)) >> level;
  *y0 = (int)a < h ? (int)a : h;
  *y1 = (int)b < h ? (int)b : h;
}
