// Vk.h — tiny POD look-alikes of the Vulkan-Hpp and glm types that appear in the public surface of the reference's
// rendergraph (LV/RenderGraph.h:285-584), so that renderer code written against legit::RenderGraph keeps its shape
// when it targets legit_cuda::RenderGraph. No Vulkan or glm headers are needed (neither exists in this image).
// Enumerator names and numeric values follow vulkan.hpp / VkFormat.
#pragma once

#include <array>
#include <cstdint>

namespace legit_cuda {

namespace glm {
struct uvec2 {
  uint32_t x = 0, y = 0;
  uvec2() = default;
  uvec2(uint32_t x_, uint32_t y_) : x(x_), y(y_) {}
  bool operator==(const uvec2 &o) const { return x == o.x && y == o.y; }
};
struct uvec3 {
  uint32_t x = 0, y = 0, z = 0;
  uvec3() = default;
  uvec3(uint32_t x_, uint32_t y_, uint32_t z_) : x(x_), y(y_), z(z_) {}
};
struct vec3 {
  float x = 0, y = 0, z = 0;
  vec3() = default;
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct vec4 {
  float x = 0, y = 0, z = 0, w = 0;
  vec4() = default;
  vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
  vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
};
} // namespace glm

namespace vk {
enum class Format : uint32_t {
  eUndefined = 0,
  eB8G8R8A8Srgb = 50,
  eR16G16B16A16Sfloat = 97,
  eR32G32Sfloat = 103,
  eR32G32B32A32Sfloat = 109,
  eD32Sfloat = 126,
};
enum class AttachmentLoadOp : uint32_t { eLoad = 0, eClear = 1, eDontCare = 2 };

struct Extent2D {
  uint32_t width = 0, height = 0;
  Extent2D() = default;
  Extent2D(uint32_t w, uint32_t h) : width(w), height(h) {}
};

struct ClearColorValue {
  std::array<float, 4> float32{{0, 0, 0, 0}};
  ClearColorValue() = default;
  ClearColorValue(const std::array<float, 4> &v) : float32(v) {}
};
struct ClearDepthStencilValue {
  float depth = 1.0f;
  uint32_t stencil = 0;
  ClearDepthStencilValue() = default;
  ClearDepthStencilValue(float d, uint32_t s) : depth(d), stencil(s) {}
};
// vulkan.hpp's ClearValue is a union; both members are kept so either reading is well defined
struct ClearValue {
  ClearColorValue color;
  ClearDepthStencilValue depthStencil;
  ClearValue() = default;
  ClearValue(const ClearColorValue &c) : color(c) {}
  ClearValue(const ClearDepthStencilValue &d) : depthStencil(d) {}
};

enum class ImageUsageFlagBits : uint32_t {
  eTransferSrc = 0x1,
  eTransferDst = 0x2,
  eSampled = 0x4,
  eStorage = 0x8,
  eColorAttachment = 0x10,
  eDepthStencilAttachment = 0x20,
};
struct ImageUsageFlags {
  uint32_t mask = 0;
  ImageUsageFlags() = default;
  ImageUsageFlags(ImageUsageFlagBits b) : mask(uint32_t(b)) {}
  explicit ImageUsageFlags(uint32_t m) : mask(m) {}
  ImageUsageFlags operator|(ImageUsageFlags o) const { return ImageUsageFlags(mask | o.mask); }
  bool operator<(const ImageUsageFlags &o) const { return mask < o.mask; }
};
inline ImageUsageFlags operator|(ImageUsageFlagBits a, ImageUsageFlagBits b) { return ImageUsageFlags(uint32_t(a) | uint32_t(b)); }
} // namespace vk

// LV/Image.h:21-22
static const vk::ImageUsageFlags colorImageUsage =
    vk::ImageUsageFlagBits::eColorAttachment | vk::ImageUsageFlagBits::eTransferDst | vk::ImageUsageFlags(vk::ImageUsageFlagBits::eSampled);
static const vk::ImageUsageFlags depthImageUsage = vk::ImageUsageFlagBits::eDepthStencilAttachment | vk::ImageUsageFlagBits::eSampled;

// LV/Synchronization.h usage types: kept as labels for API parity; CUDA stream order replaces the barriers they drive.
enum struct ImageUsageTypes { GraphicsShaderRead, GraphicsShaderReadWrite, ComputeShaderRead, ComputeShaderReadWrite, TransferDst, TransferSrc, ColorAttachment, DepthAttachment, Present, None, Unknown };

// LegitProfiler/ProfilerTask.h:9-31 (flat-UI palette, RGBA little endian)
namespace Colors {
constexpr uint32_t rgbaLE(uint32_t c) { return ((c & 0xff000000u) >> 24) | ((c & 0x00ff0000u) >> 8) | ((c & 0x0000ff00u) << 8) | ((c & 0x000000ffu) << 24); }
constexpr uint32_t turqoise = rgbaLE(0x1abc9cffu), greenSea = rgbaLE(0x16a085ffu), emerald = rgbaLE(0x2ecc71ffu), nephritis = rgbaLE(0x27ae60ffu);
constexpr uint32_t peterRiver = rgbaLE(0x3498dbffu), belizeHole = rgbaLE(0x2980b9ffu), amethyst = rgbaLE(0x9b59b6ffu), wisteria = rgbaLE(0x8e44adffu);
constexpr uint32_t sunFlower = rgbaLE(0xf1c40fffu), orange = rgbaLE(0xf39c12ffu), carrot = rgbaLE(0xe67e22ffu), pumpkin = rgbaLE(0xd35400ffu);
constexpr uint32_t alizarin = rgbaLE(0xe74c3cffu), pomegranate = rgbaLE(0xc0392bffu), clouds = rgbaLE(0xecf0f1ffu), silver = rgbaLE(0xbdc3c7ffu);
} // namespace Colors

} // namespace legit_cuda
