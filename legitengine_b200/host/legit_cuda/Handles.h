// Handles.h — move-only owning handle for rendergraph proxies (contract of LV/Handles.h:9-69):
// the destructor calls info.Reset() (which returns the proxy's pool slot to the graph) unless Detach()ed;
// only the Factory can mint attached handles.
#pragma once

#include <utility>

namespace legit_cuda {

template <typename HandleInfo, typename Factory> class UniqueHandle {
public:
  UniqueHandle() = default;
  UniqueHandle(const UniqueHandle &) = delete;
  UniqueHandle &operator=(const UniqueHandle &) = delete;
  UniqueHandle(UniqueHandle &&o) noexcept : info_(o.info_), attached_(o.attached_) { o.attached_ = false; }
  UniqueHandle &operator=(UniqueHandle &&o) noexcept {
    if (this != &o) {
      if (attached_) info_.Reset();
      info_ = o.info_;
      attached_ = o.attached_;
      o.attached_ = false;
    }
    return *this;
  }
  ~UniqueHandle() {
    if (attached_) info_.Reset();
  }
  void Detach() { attached_ = false; }
  void Reset() {
    if (attached_) info_.Reset();
    attached_ = false;
  }
  bool IsAttached() const { return attached_; }
  const HandleInfo &Get() const { return info_; }
  const HandleInfo *operator->() const { return &info_; }

private:
  friend Factory;
  explicit UniqueHandle(const HandleInfo &info) : info_(info), attached_(true) {}
  HandleInfo info_{};
  bool attached_ = false;
};

} // namespace legit_cuda
