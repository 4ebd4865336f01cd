// Pool.h — slot pool with stable integer ids, the id type behind every rendergraph proxy.
// Same contract as the reference's Utils::Pool (LV/Pool.h): Add() reuses the most recently released slot (LIFO),
// ids are plain indices (`asInt`), a default-constructed id is invalid (size_t(-1)), iteration skips free slots.
#pragma once

#include <cassert>
#include <cstddef>
#include <utility>
#include <vector>

namespace legit_cuda {
namespace Utils {

template <typename T> class Pool {
public:
  struct Id {
    size_t asInt = size_t(-1);
    Id() = default;
    Id(size_t v) : asInt(v) {}
    bool operator==(const Id &o) const { return asInt == o.asInt; }
    bool operator!=(const Id &o) const { return asInt != o.asInt; }
    bool IsValid() const { return asInt != size_t(-1); }
  };

  Id Add(T &&value) {
    if (!recycled_.empty()) {
      const size_t slot = recycled_.back();
      recycled_.pop_back();
      slots_[slot] = std::move(value);
      live_[slot] = true;
      return Id(slot);
    }
    slots_.emplace_back(std::move(value));
    live_.push_back(true);
    return Id(slots_.size() - 1);
  }

  void Release(Id id) {
    assert(IsPresent(id));
    live_[id.asInt] = false;
    recycled_.push_back(id.asInt);
  }

  T &Get(Id id) {
    assert(IsPresent(id));
    return slots_[id.asInt];
  }
  const T &Get(Id id) const {
    assert(IsPresent(id));
    return slots_[id.asInt];
  }
  size_t GetSize() const { return slots_.size(); }
  bool IsPresent(Id id) const { return id.asInt < slots_.size() && live_[id.asInt]; }

  // visit every live element with its id
  template <typename F> void ForEach(F &&f) {
    for (size_t i = 0; i < slots_.size(); i++)
      if (live_[i]) f(Id(i), slots_[i]);
  }

private:
  std::vector<T> slots_;
  std::vector<bool> live_;
  std::vector<size_t> recycled_;
};

} // namespace Utils
} // namespace legit_cuda
