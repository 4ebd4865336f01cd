/* TEST STUB — the handful of Vulkan core types include/lgcu_vulkan.h touches, with the enumerant values and member order of the
 * Khronos header (vulkan_core.h, VK_VERSION_1_0), so that the shim can be compiled in an image that has no Vulkan SDK. */
#ifndef VULKAN_CORE_H_
#define VULKAN_CORE_H_ 1
#include <stdint.h>
typedef uint64_t VkDeviceSize;
typedef enum VkFormat {
  VK_FORMAT_UNDEFINED = 0,
  VK_FORMAT_B8G8R8A8_SRGB = 50,
  VK_FORMAT_R16G16B16A16_SFLOAT = 97,
  VK_FORMAT_R32G32_SFLOAT = 103,
  VK_FORMAT_R32G32B32A32_SFLOAT = 109,
  VK_FORMAT_D32_SFLOAT = 126,
  VK_FORMAT_MAX_ENUM = 0x7FFFFFFF
} VkFormat;
typedef struct VkExtent3D {
  uint32_t width, height, depth;
} VkExtent3D;
typedef struct VkSubresourceLayout {
  VkDeviceSize offset, size, rowPitch, arrayPitch, depthPitch;
} VkSubresourceLayout;
#endif
