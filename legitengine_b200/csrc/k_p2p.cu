// k_p2p.cu — peer-to-peer halo exchange primitives for the strip-sharded frame: our own kernels moving rows over
// NVLink through peer memory mapped into this process (CUDA IPC), instead of one NCCL send/recv pair per slab.
//
//   rowCopyKernel : copies up to kMaxCopies contiguous slabs in ONE launch; sources and/or destinations may be peer-GPU
//                   addresses (ld/st.global on a peer mapping travels over NVLink, local L2 is bypassed for peer lines and the
//                   L1 is flushed at every launch, so a pull always sees what the owner's previous kernels wrote).
//   signalKernel  : publishes "my stage s of frame f is complete" into flag words that live in the PEERS' memory.
//   waitKernel    : spins (one thread) until the local flag words written by the peers reach the current frame number.
// The frame number lives in device memory and is bumped by a kernel, so a whole frame — stages, signals, waits, copies —
// replays from a CUDA graph with no host involvement. Streams are in-order per GPU and no wait depends on a later signal of
// the waiting GPU itself, so the protocol cannot deadlock (see multigpu.P2PStripRenderer for the per-frame order).
#include "lgcu_kernels.h"

namespace lgcu {
namespace {

__global__ void __launch_bounds__(256) rowCopyKernel(const __grid_constant__ RowCopyArgs a) {
  // work unit = 16 bytes; slabs are laid end to end in unit space
  const uint64_t total = a.unitEnd[a.count - 1];
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
    int s = 0;
    while (u >= a.unitEnd[s]) s++;
    const uint64_t local = u - (s ? a.unitEnd[s - 1] : 0);
    const uint4 v = __ldcv(reinterpret_cast<const uint4 *>(a.src[s]) + local);
    reinterpret_cast<uint4 *>(a.dst[s])[local] = v;
  }
}

__global__ void bumpFrameKernel(uint32_t *frame) { *frame = *frame + 1u; }

__global__ void signalKernel(const __grid_constant__ FlagArgs a, const uint32_t *frame) {
  const int i = threadIdx.x;
  if (i >= a.count) return;
  __threadfence_system(); // everything this GPU wrote before (previous kernels, incl. pushes into peer memory) is visible first
  *reinterpret_cast<volatile uint32_t *>(a.flags[i]) = *frame;
}

__global__ void waitKernel(const __grid_constant__ FlagArgs a, const uint32_t *frame, int lag) {
  const int i = threadIdx.x;
  if (i >= a.count) return;
  const uint32_t want = *frame - (uint32_t)lag;
  const volatile uint32_t *f = reinterpret_cast<const volatile uint32_t *>(a.flags[i]);
  while ((int32_t)(*f - want) < 0) __nanosleep(200);
  __threadfence_system();
}

// One exchange step in one launch (lgcu_exchange): the halo transfer is fused with its own synchronisation. Every CTA waits for the
// owners' flags by itself (so the grid needs no barrier and no co-residency), the whole grid then moves the slabs over NVLink, and the
// CTA that finishes last publishes the acknowledgement. A frame of the strip protocol is 4 of these instead of 12 one-warp kernels and
// copy launches, which is worth ~20 us of launch gaps on a frame of one millisecond.
__global__ void __launch_bounds__(256) exchangeKernel(const __grid_constant__ ExchangeArgs a) {
  __shared__ uint32_t sFrame;
  __shared__ bool sLast;
  if (threadIdx.x == 0) {
    if (a.bump) *a.frame = *a.frame + 1u; // single-CTA launches only (checked on the host)
    sFrame = *reinterpret_cast<volatile uint32_t *>(a.frame);
  }
  __syncthreads();
  const uint32_t frame = sFrame;
  if (blockIdx.x == 0 && (int)threadIdx.x < a.signalBefore.count) {
    __threadfence_system(); // everything this GPU wrote before (previous kernels, incl. stores into peer memory) is visible first
    *reinterpret_cast<volatile uint32_t *>(a.signalBefore.flags[threadIdx.x]) = frame;
  }
  if ((int)threadIdx.x < a.wait.count) {
    const uint32_t want = frame - (uint32_t)a.lag;
    const volatile uint32_t *f = reinterpret_cast<const volatile uint32_t *>(a.wait.flags[threadIdx.x]);
    while ((int32_t)(*f - want) < 0) __nanosleep(100);
    __threadfence_system();
  }
  __syncthreads();
  if (a.copies.count > 0) {
    const uint64_t total = a.copies.unitEnd[a.copies.count - 1];
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
      int s = 0;
      while (u >= a.copies.unitEnd[s]) s++;
      const uint64_t local = u - (s ? a.copies.unitEnd[s - 1] : 0);
      reinterpret_cast<uint4 *>(a.copies.dst[s])[local] = __ldcv(reinterpret_cast<const uint4 *>(a.copies.src[s]) + local);
    }
  }
  if (a.signalAfter.count > 0) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) sLast = atomicAdd(a.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (sLast) {
      if (threadIdx.x == 0) *a.done = 0u; // ready for the next launch
      if ((int)threadIdx.x < a.signalAfter.count) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t *>(a.signalAfter.flags[threadIdx.x]) = frame;
      }
    }
  }
}

} // namespace

cudaError_t launchExchange(const ExchangeArgs &a, int smCount, cudaStream_t s) {
  uint64_t blocks = 1;
  if (a.copies.count > 0) {
    const uint64_t total = a.copies.unitEnd[a.copies.count - 1];
    blocks = (total + 1023) / 1024; // 4 units per thread
    if (blocks > (uint64_t)smCount * 4) blocks = (uint64_t)smCount * 4;
    if (blocks < 1) blocks = 1;
  }
  exchangeKernel<<<(unsigned)blocks, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launchRowCopies(const RowCopyArgs &a, int smCount, cudaStream_t s) {
  if (a.count <= 0) return cudaSuccess;
  const uint64_t total = a.unitEnd[a.count - 1];
  if (!total) return cudaSuccess;
  uint64_t blocks = (total + 1023) / 1024; // 4 units per thread
  if (blocks > (uint64_t)smCount * 4) blocks = (uint64_t)smCount * 4;
  rowCopyKernel<<<(unsigned)blocks, 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launchBumpFrame(uint32_t *frame, cudaStream_t s) {
  bumpFrameKernel<<<1, 1, 0, s>>>(frame);
  return cudaGetLastError();
}
cudaError_t launchSignal(const FlagArgs &a, const uint32_t *frame, cudaStream_t s) {
  if (a.count <= 0) return cudaSuccess;
  signalKernel<<<1, 32, 0, s>>>(a, frame);
  return cudaGetLastError();
}
cudaError_t launchWait(const FlagArgs &a, const uint32_t *frame, int lag, cudaStream_t s) {
  if (a.count <= 0) return cudaSuccess;
  waitKernel<<<1, 32, 0, s>>>(a, frame, lag);
  return cudaGetLastError();
}

} // namespace lgcu
