// ShaderMemoryPool.h — per-frame bump allocator for pass parameter blocks.
// The reference writes UBO data through BeginSet / GetUniformBufferData<T>("Name") / EndSet into a persistently
// mapped 100 MB buffer (LV/ShaderMemoryPool.h:33-96) and binds it with a dynamic offset. The CUDA passes take the same
// tightly packed structs by host pointer, so this pool simply hands out stable host storage for one frame; the name is
// kept as a label (the reference resolves it through SPIR-V reflection, which has no analogue here).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

namespace legit_cuda {

class ShaderMemoryPool {
public:
  void MapBuffer() { blocks_.clear(); } // start of frame (LV/PresentQueue.h:113)
  void UnmapBuffer() {}

  struct SetDynamicUniformBindings {
    uint32_t dynamicOffset = 0;
  };
  SetDynamicUniformBindings BeginSet(const void * /*setInfo*/ = nullptr) {
    SetDynamicUniformBindings b;
    b.dynamicOffset = uint32_t(blocks_.size());
    return b;
  }
  template <typename BufferType> BufferType *GetUniformBufferData(const std::string &bufferName) {
    blocks_.emplace_back(sizeof(BufferType));
    names_.push_back(bufferName);
    std::memset(blocks_.back().data(), 0, sizeof(BufferType));
    return reinterpret_cast<BufferType *>(blocks_.back().data());
  }
  void EndSet() {}

private:
  std::deque<std::vector<uint8_t>> blocks_; // deque: growing never moves handed-out blocks
  std::vector<std::string> names_;
};

} // namespace legit_cuda
