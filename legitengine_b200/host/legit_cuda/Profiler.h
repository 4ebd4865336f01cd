// Profiler.h — per-pass CPU and GPU timing with the reference's task shape (LegitProfiler/ProfilerTask.h:36-46:
// {startTime, endTime, name, color}). The GPU profiler records a cudaEvent at the start of every pass and one at the
// end of the frame, like the reference's one-timestamp-per-pass scheme (LV/GpuProfiler.h:13-27, 97-111): a pass lasts
// until the next pass starts.
#pragma once

#include <cuda_runtime_api.h>

#include <chrono>
#include <string>
#include <vector>

namespace legit_cuda {

struct ProfilerTask {
  double startTime = 0, endTime = 0; // seconds since frame start
  std::string name;
  uint32_t color = 0;
  double GetLength() const { return endTime - startTime; }
};

class CpuProfiler {
public:
  void StartFrame() {
    tasks_.clear();
    origin_ = Clock::now();
  }
  size_t StartTask(const std::string &name, uint32_t color) {
    ProfilerTask t;
    t.name = name;
    t.color = color;
    t.startTime = Now();
    tasks_.push_back(t);
    return tasks_.size() - 1;
  }
  void EndTask(size_t id) { tasks_[id].endTime = Now(); }
  const std::vector<ProfilerTask> &GetProfilerTasks() const { return tasks_; }

private:
  using Clock = std::chrono::steady_clock;
  double Now() const { return std::chrono::duration<double>(Clock::now() - origin_).count(); }
  Clock::time_point origin_ = Clock::now();
  std::vector<ProfilerTask> tasks_;
};

class GpuProfiler {
public:
  ~GpuProfiler() {
    for (auto e : events_) cudaEventDestroy(e);
  }
  void StartFrame() {
    used_ = 0;
    names_.clear();
    colors_.clear();
  }
  void StartTask(cudaStream_t stream, const std::string &name, uint32_t color) {
    cudaEventRecord(NextEvent(), stream);
    names_.push_back(name);
    colors_.push_back(color);
  }
  void EndFrame(cudaStream_t stream) { cudaEventRecord(NextEvent(), stream); }
  // Blocks until the frame's last event has completed.
  std::vector<ProfilerTask> GatherTasks() {
    std::vector<ProfilerTask> tasks;
    if (used_ < 2) return tasks;
    cudaEventSynchronize(events_[used_ - 1]);
    for (size_t i = 0; i + 1 < used_; i++) {
      float ms0 = 0, ms1 = 0;
      cudaEventElapsedTime(&ms0, events_[0], events_[i]);
      cudaEventElapsedTime(&ms1, events_[0], events_[i + 1]);
      ProfilerTask t;
      t.startTime = ms0 * 1e-3;
      t.endTime = ms1 * 1e-3;
      t.name = names_[i];
      t.color = colors_[i];
      tasks.push_back(t);
    }
    return tasks;
  }

private:
  cudaEvent_t NextEvent() {
    if (used_ == events_.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      events_.push_back(e);
    }
    return events_[used_++];
  }
  std::vector<cudaEvent_t> events_;
  size_t used_ = 0;
  std::vector<std::string> names_;
  std::vector<uint32_t> colors_;
};

} // namespace legit_cuda
