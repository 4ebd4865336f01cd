"""legitengine_b200 — B200-native (sm_100a) CUDA implementation of LegitEngine's per-pixel lighting and
screen-space GI passes (the SSVGIRenderer hot path) behind the reference's RenderGraph / RenderPassDesc API.

Layout: csrc/ = CUDA kernels + the C ABI (include/lgcu.h); host/ = C++ mirror of the reference's rendergraph,
MipBuilder, BlurBuilder and SSVGIRenderer pass list plus the headless harness; this package = thin ctypes access
for tests and the benchmark. PyTorch is used only as a device-memory / stream / torch.distributed provider.
"""
__version__ = "0.1.0"
