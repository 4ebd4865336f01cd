#!/usr/bin/env python
"""bench.py — full SSVGI GI frame (G-buffer resolve -> direct light -> mip build -> blur -> GI gather -> denoise -> final
gather) on B200, measured as BASELINE.json's metric: frame ms and Mpix/s, with the HBM roofline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 4k|8k|1080p|WxH] [--impl ours|reference]

One "step" = one frame of the workload through the C++ rendergraph harness (liblegit_cuda.so -> liblgcu.so kernels).
  value : whole-job Mpix/s with the rasterised scene (SURVEY.md §8d's per-pixel fragment buffer) already resident in HBM
          (CUDA-graph replay of the fused frame)
  e2e   : the frame as the reference's RenderFrame receives it — the SCENE (vertex / index buffers, draw list, per-object
          constants) in HOST memory: every step copies the scene host->device from pinned memory, rasterises it on the device
          (ShadowPass + GBufferRasterPass), renders the frame and copies the BGRA8 swapchain image device->host, all inside the
          timed region. value_from_mesh is the same frame with the scene resident. e2e_fragments is the other host-buffer form:
          the pre-rasterised 32 B/px fragment buffer uploaded every step (PCIe-bound).
N > 1   : one process per GPU (torchrun). Default: BASELINE configs[3] — ONE 7680x4320 frame cut into row strips with NVLink halo
          exchange ("strong"); the line also carries the same run's single-GPU 8K frame time (`single_gpu`) and the throughput mode
          of configs[4] (`replicas`: independent 4K frames per GPU, no data-path collective). --shard replicas makes that the line.
--impl reference : the reference's own SPIR-V passes on the host cores (oracle/_ref, else the C port), bounded sample.
The CPU oracle is used here ONLY for the cpu_baseline / reference legs — never on the product path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {"4k": (3840, 2160), "8k": (7680, 4320), "1080p": (1920, 1080), "512": (512, 512)}
# pass-granular algorithmic bytes per base pixel (SURVEY.md §8d / BASELINE.md §3)
BYTES_PER_PX = {"resolve": 68.0, "light": 36.0, "mips": 26.67, "blur": 42.67, "gather": 41.33, "denoise": 16.0, "final": 28.0}
FRAME_BYTES_PER_PX = 258.67
SHADOW_MAP_BYTES = 4 * 1024 * 1024
METRIC = "full_gi_frame_mpix_per_s"
SM_COUNT, ISSUE_PER_SM_PER_CLK = 148, 4  # 4 warp schedulers per SM, one warp instruction per scheduler per clock
FP32_ISSUE_PER_SM_PER_CLK = 2.76           # measured scalar FFMA issue rate, warp instructions / clock / SM (profiles/r01e_fp32_pipe_rates.txt)


def gather_ncu():
    """ncu counters of the gather kernels of the CURRENT tree on the 4K bench frame (profiles/gather_ncu.json, written by
    scripts/ncu_summary.py gather-json from the committed ncu CSV logs): DRAM bytes and executed instructions of the throughput kernel,
    and the FP32 thread-instruction count of the shader-order kernel = the pass's algorithmic FP32 work."""
    try:
        return json.loads((ROOT / "profiles" / "gather_ncu.json").read_text())
    except Exception:
        return None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default=None, help="4k | 8k | 1080p | 512 | WxH (default: 4k on one GPU, 8k for the strip-sharded frame on N > 1)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fused", choices=["fused", "passes"])
    ap.add_argument("--strict", action="store_true", help="use the shader-order parity gather kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default: min(steps, 60))")
    ap.add_argument("--frames-in-flight", type=int, default=3, help="e2e leg: frames in flight (renderer + stream + pinned swapchain per slot)")
    ap.add_argument("--shard", default=None, choices=["replicas", "strips"],
                    help="N > 1: replicas = independent frames per GPU (weak scaling, BASELINE configs[4]); strips = ONE frame cut into row strips "
                         "with NVLink halo exchange (strong scaling, BASELINE configs[3])")
    ap.add_argument("--no-present", action="store_true", help="strips: skip the composite of the swapchain strips on rank 0")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="strips: p2p = our copy/flag kernels over NVLink peer memory (CUDA IPC); nccl = torch.distributed send/recv per slab")
    ap.add_argument("--strip-input", default="fragments", choices=["fragments", "mesh"],
                    help="strips: what every rank starts from — its strip of the pre-rasterised fragment buffer (uploaded per frame in the e2e leg), or the "
                         "mesh scene, which every rank rasterises for its own rows on the device (ShadowPass + GBufferRasterPass with lgcu_rows)")
    ap.add_argument("--balance", type=int, default=5, help="strips (p2p): up to this many measure -> rebalance rounds of the strip boundaries (0 = equal rows)")
    ap.add_argument("--strip-bounds", default=None, help="strips: fixed boundaries y0,y1,...,yN (multiples of 16) instead of measuring and re-balancing")
    ap.add_argument("--unfused-exchange", action="store_true", help="strips (p2p): separate signal / wait / copy launches instead of one lgcu_exchange kernel per step (A/B)")
    ap.add_argument("--skip-extras", action="store_true", help="strips: skip the single-GPU / replica / batch legs (A/B runs)")
    ap.add_argument("--no-graph", action="store_true", help="strips: launch stages and NCCL transfers from Python every frame instead of replaying a CUDA graph")
    return ap.parse_args()


def workload_size(name: str):
    name = name or "4k"
    if name in WORKLOADS:
        return WORKLOADS[name]
    w, h = name.lower().split("x")
    return int(w), int(h)


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons during the timed region (NVML, ~20 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference / CPU legs
def cpu_frame_seconds(width: int, height: int, frames: int, warmup: int = 1):
    """Times the reference's SPIR-V passes (oracle/_ref; the C port if the reference arm is not built) on the host cores."""
    from legitengine_b200 import images, passes, scene
    from oracle import loader

    be = loader.ref() if loader.have_ref() else loader.port()
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)
    be.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    sc = scene.make_scene(0xC0FFEE, width, height)
    p = passes.make_params(width, height, sc.matrices, 0)
    fi = passes.FrameImages(width, height, images.HostImage)
    inp = passes.upload_inputs(fi, sc)
    for _ in range(warmup):
        passes.run_pass_list(be, fi, p, inp)
    times = []
    for _ in range(frames):
        t0 = time.perf_counter()
        passes.run_pass_list(be, fi, p, inp)
        times.append(time.perf_counter() - t0)
    return be.kind, int(be.num_threads()), times


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    shard = args.shard or ("strips" if world > 1 else "replicas")
    W, H = workload_size(args.workload or ("8k" if (world > 1 and shard == "strips") else "4k"))  # the workload of our arm's line at this N
    div = 4 if W * H <= 3840 * 2160 else 8  # bounded sample: 960x540 per step at 4K and 8K (~0.2 s on 16 cores)
    sw, sh = max(W // div, 64), max(H // div, 36)
    kind, cores, times = cpu_frame_seconds(sw, sh, args.steps, min(args.warmup, 1))
    sec = float(np.mean(times))
    value = sw * sh / sec / 1e6
    sample = f"{sw}x{sh} full frame (1/{div * div} of the {W}x{H} pixels) per step, same synthetic scene generator, all passes K1..K7"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if (world > 1 and shard == "strips") else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W}x{H} full GI frame, CPU sample: {sample}",
                   "what": "reference SPIR-V passes (spirv-cross C++ backend + glm) on host cores" if kind == "reference" else "plain-C port of the reference shaders on host cores"},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch

    from legitengine_b200 import abi, harness, scene

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    W, H = workload_size(args.workload)
    npx = W * H
    mode = harness.MODE_FUSED if args.mode == "fused" else harness.MODE_PASS_GRANULAR
    gi_flags = abi.GI_STRICT if args.strict else abi.GI_DEFAULT

    # synthetic rasterised scene, generated on the host into pinned memory. Every rank renders the SAME scene: the gather's cost depends
    # on what the frame shows, and with different scenes the MAX over ranks would measure the scenes, not the GPUs
    m = scene.frame_matrices(W, H)
    seed = 0xC0FFEE
    frag_host = torch.empty((H, W * 32), dtype=torch.uint8).pin_memory()
    frags = frag_host.numpy().view(abi.FRAGMENT_DTYPE).reshape(H, W)
    scene.scene_fragments(seed, W, H, m, out=frags)
    objects = scene.scene_objects(seed)
    shadow = torch.from_numpy(scene.scene_shadow_map(seed, m)).pin_memory()
    swap_host = torch.empty((H, W * 4), dtype=torch.uint8).pin_memory()

    stream = torch.cuda.Stream()
    r = harness.Renderer(W, H, stream=stream.cuda_stream)
    r.upload_fragments(frag_host.data_ptr(), W * 32)
    r.upload_objects(objects.ctypes.data, len(objects))
    r.upload_light_depth(shadow.data_ptr(), 1024)
    r.sync()

    # first frame allocates the images; per-pass GPU profile of the un-captured frame (events between passes)
    r.render_frame(mode, 0, gi_flags)
    r.sync()
    pass_ms = {}
    prof_frames = 5
    for _ in range(prof_frames):
        r.render_frame(mode, 0, gi_flags, profile=True)
        r.sync()
        for name, ms in r.profile():
            pass_ms[name] = pass_ms.get(name, 0.0) + ms / prof_frames
    passes_per_frame = r.last_pass_count()
    # the same frame with the reference's other denoiser setting (radius 2, SSVGIRenderer.h:288): per-pass times only
    pass_ms_r2 = {}
    r.render_frame(mode, 2, gi_flags)
    r.sync()
    for _ in range(prof_frames):
        r.render_frame(mode, 2, gi_flags, profile=True)
        r.sync()
        for name, ms in r.profile():
            pass_ms_r2[name] = pass_ms_r2.get(name, 0.0) + ms / prof_frames
    r.render_frame(mode, 0, gi_flags)
    r.sync()

    r.capture_frame(mode, 0, gi_flags)
    kernels_per_frame = r.captured_kernel_count()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup, sample_clocks=False):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(steps):
                fn()
            ev1.record(stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    # --- value: device-resident inputs, CUDA-graph replay
    ms_total, clocks = timed(r.replay_frame, args.steps, max(args.warmup, 3), sample_clocks=True)
    ms_per_step = ms_total / args.steps
    value = world * npx / (ms_per_step * 1e-3) / 1e6

    # per-frame distribution of the same replayed frame (median and p10 / p90), one event pair per frame
    dist_frames = 100
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(dist_frames)]
    with torch.cuda.stream(stream):
        for e0, e1 in pairs:
            e0.record(stream)
            r.replay_frame()
            e1.record(stream)
    torch.cuda.synchronize()
    per_frame = sorted(e0.elapsed_time(e1) for e0, e1 in pairs)
    frame_ms = {"median": per_frame[dist_frames // 2], "p10": per_frame[dist_frames // 10], "p90": per_frame[(dist_frames * 9) // 10], "frames": dist_frames}

    # --- e2e: host fragments in, swapchain out, every step. Like the reference's InFlightQueue (LV/PresentQueue.h:62: one command
    # buffer per in-flight frame) several frames are in flight: each slot owns a renderer (image set + captured graph), a stream and
    # a pinned swapchain buffer, so the H2D copy of frame i+1 and the D2H copy of frame i-1 overlap the kernels of frame i.
    in_flight = max(1, args.frames_in_flight)
    slots = [(r, stream, swap_host)]
    for _ in range(in_flight - 1):
        s_i = torch.cuda.Stream()
        r_i = harness.Renderer(W, H, stream=s_i.cuda_stream)
        r_i.upload_fragments(frag_host.data_ptr(), W * 32)
        r_i.upload_objects(objects.ctypes.data, len(objects))
        r_i.upload_light_depth(shadow.data_ptr(), 1024)
        r_i.render_frame(mode, 0, gi_flags)
        r_i.sync()
        r_i.capture_frame(mode, 0, gi_flags)
        slots.append((r_i, s_i, torch.empty((H, W * 4), dtype=torch.uint8).pin_memory()))
    counter = [0]

    def e2e_step():
        r_i, _, swap_i = slots[counter[0] % in_flight]
        counter[0] += 1
        r_i.upload_fragments(frag_host.data_ptr(), W * 32)
        r_i.replay_frame()
        r_i.download_swapchain(swap_i.data_ptr(), W * 4)

    def e2e_timed(steps, warmup):
        for _ in range(warmup):
            e2e_step()
        barrier()
        for _, s_i, _ in slots[1:]:
            s_i.wait_stream(stream)
        ev0.record(stream)
        for _ in range(steps):
            e2e_step()
        for _, s_i, _ in slots[1:]:
            stream.wait_stream(s_i)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    e2e_steps = args.e2e_steps or min(args.steps, 60)
    e2e_frag_ms = e2e_timed(e2e_steps, max(3, in_flight))
    e2e_frag_value = world * npx / (e2e_frag_ms / e2e_steps * 1e-3) / 1e6

    # --- the same frame from the scene in the reference's form (mesh): every slot switches to mesh input and re-captures
    mesh = scene.scene_mesh(seed)
    pinned = {}
    for name in ("vertices", "indices", "draws", "objects"):
        a = getattr(mesh, name)
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype)
        v[...] = a
        pinned[name] = (t, v)
    mesh_pinned = scene.Mesh(*(pinned[n][1] for n in ("vertices", "indices", "draws", "objects")))
    mesh_bytes = sum(pinned[n][0].numel() for n in pinned)
    for r_i, s_i, _ in slots:
        r_i.upload_mesh(mesh_pinned)
        r_i.render_frame(mode, 0, gi_flags)
        r_i.sync()
    pass_ms_mesh = {}
    for _ in range(prof_frames):
        r.render_frame(mode, 0, gi_flags, profile=True)
        r.sync()
        for name, ms in r.profile():
            pass_ms_mesh[name] = pass_ms_mesh.get(name, 0.0) + ms / prof_frames
    for r_i, _, _ in slots:
        r_i.capture_frame(mode, 0, gi_flags)
    kernels_per_frame_mesh = r.captured_kernel_count()
    mesh_ms_total, _ = timed(r.replay_frame, args.steps, max(args.warmup, 3))
    mesh_ms_per_step = mesh_ms_total / args.steps

    def e2e_step():  # noqa: F811 — scene in, swapchain out
        r_i, _, swap_i = slots[counter[0] % in_flight]
        counter[0] += 1
        r_i.upload_mesh(mesh_pinned)
        r_i.replay_frame()
        r_i.download_swapchain(swap_i.data_ptr(), W * 4)

    e2e_ms = e2e_timed(e2e_steps, max(3, in_flight))
    e2e_value = world * npx / (e2e_ms / e2e_steps * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        gather_ms = pass_ms.get("IndirectLightPass", 0.0)
        gather_bytes = BYTES_PER_PX["gather"] * npx
        frame_bytes = FRAME_BYTES_PER_PX * npx + SHADOW_MAP_BYTES
        ncu = gather_ncu() if (W, H) == (3840, 2160) and not args.strict else None
        roofline = {
            "kernel": "gi_gather (IndirectLightPass)", "bound": "hbm", "achieved": gather_bytes / (gather_ms * 1e-3) / 1e9 if gather_ms else None,
            "peak": peak, "unit": "GB/s", "frac": (gather_bytes / (gather_ms * 1e-3) / 1e9 / peak) if gather_ms else None,
            "traffic": ncu["fast"]["dram_bytes"] if ncu else None, "traffic_source": f"profiles/gather_ncu.json ({ncu['tag']})" if ncu else None,
            "algorithmic_bytes": gather_bytes, "peak_source": peak_src, "ms": gather_ms,
            "note": "the gather is bound by instruction issue and load latency, not by HBM (SURVEY.md F7): achieved = 41.33 B/px compulsory bytes / its measured time; "
                    "see gather_fp32 for the roofline that bounds it",
        }
        gather_fp32 = issue_util = None
        if ncu and gather_ms and clocks and clocks.get("sm_mhz"):
            clk = clocks["sm_mhz"] * 1e6
            fp32_peak = SM_COUNT * FP32_ISSUE_PER_SM_PER_CLK * 32 * clk / 1e12  # T thread-instructions / s
            alg = ncu["strict"]["fp32_thread_inst"]
            gather_fp32 = {"kernel": "gi_gather (IndirectLightPass)", "bound": "fp32 issue", "algorithmic_fp32_thread_inst": alg,
                           "achieved": alg / (gather_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "T thread-inst/s", "frac": alg / (gather_ms * 1e-3) / 1e12 / fp32_peak,
                           "executed_fp32_thread_inst": ncu["fast"]["fp32_thread_inst"],
                           "note": "algorithmic work = FP32 thread instructions the shader-order kernel executes on this frame (ncu smsp__sass_thread_inst_executed_op_fp32, "
                                   "profiles/gather_ncu.json) / measured time, against 148 SMs x 2.76 warp-inst/clk (measured FFMA issue rate) x 32 lanes x the SM clock "
                                   "under load; > 1 is possible because the throughput kernel needs fewer FP32 instructions than the shader's own order (affine rays, "
                                   "no per-sample unprojection, atan only on hits)"}
            issue_peak = SM_COUNT * ISSUE_PER_SM_PER_CLK * clk / 1e9
            issue_util = {"kernel": "gi_gather (IndirectLightPass)", "executed_warp_inst": ncu["fast"]["warp_inst"], "rate": ncu["fast"]["warp_inst"] / (gather_ms * 1e-3) / 1e9,
                          "issue_slots": issue_peak, "unit": "Gwarp-inst/s", "utilisation": ncu["fast"]["warp_inst"] / (gather_ms * 1e-3) / 1e9 / issue_peak,
                          "note": "a UTILISATION figure, not a roofline fraction: instructions the kernel itself executes / issue slots available; it says how "
                                  "much of the remaining time is stalls, not how good the kernel is"}
        roofline_frame = {"bound": "hbm", "achieved": frame_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": frame_bytes / (ms_per_step * 1e-3) / 1e9 / peak, "algorithmic_bytes": frame_bytes,
                          "note": "whole frame, pass-granular algorithmic bytes 258.67 B/px + 4 MiB shadow map"}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{W}x{H} full GI frame: G-buffer resolve + shadow-mapped direct light + 2x9 mip levels + 2x10 blur levels + SSVGI gather + denoise(r=0) + final gather (BASELINE configs[2] incl. resolve)"
                            + (f"; {world} GPUs render independent frames (configs[4] throughput mode)" if world > 1 else ""),
                "pass_list": "fused (K1+K2, mip+blur chain x2, K5, K6+K7)" if mode == harness.MODE_FUSED else "pass-granular (reference's 46 passes)",
                "gather_kernel": "strict" if args.strict else "fast", "cuda_graph": True, "passes_per_frame": passes_per_frame,
                "l2": "inputs larger than L2 (frame working set %.0f MB vs 126 MB L2)" % (r.allocated_bytes() / 1e6),
                "storage": "RGBA16F / RG32F / D32F / BGRA8 images exactly as the reference allocates them",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": mesh_bytes, "d2h_bytes_per_step": W * H * 4, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps, "frames_in_flight": in_flight,
                    "input": "scene as the reference holds it (vertex + index buffers, draw list, per-object constants: %d triangles), uploaded from pinned host "
                             "memory every step and rasterised on the device (ShadowPass + GBufferRasterPass); output: BGRA8 swapchain image to pinned host memory"
                             % mesh.triangle_count},
            "e2e_fragments": {"value": e2e_frag_value, "unit": "Mpix/s", "h2d_bytes_per_step": W * H * 32, "d2h_bytes_per_step": W * H * 4, "steps": e2e_steps,
                              "ms_per_step": e2e_frag_ms / e2e_steps, "frames_in_flight": in_flight,
                              "input": "pre-rasterised 32 B/px fragment buffer uploaded every step (PCIe-bound)"},
            "value_from_mesh": {"value": world * npx / (mesh_ms_per_step * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": mesh_ms_per_step,
                                "kernels_per_frame": kernels_per_frame_mesh, "note": "scene resident; frame = ShadowPass + GBufferRasterPass + the passes of `value`"},
            "gpu_launches": kernels_per_frame * args.steps,
            "kernels_per_frame": kernels_per_frame,
            "roofline": roofline,
            "gather_fp32": gather_fp32,
            "gather_issue_slot_utilisation": issue_util,
            "roofline_frame": roofline_frame,
            "frame_ms": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in frame_ms.items()},
            "pass_ms": {k: round(v, 4) for k, v in pass_ms.items()},
            "pass_ms_denoiser_radius2": dict({k: round(v, 4) for k, v in pass_ms_r2.items()},
                                             note="K6 with the 4x4 depth-guided fit (fused with K7): 24 B/px algorithmic = %.1f us at the measured HBM peak" % (24.0 * npx / (measured_peak_gbs()[0] * 1e9) * 1e6)),
            "pass_ms_mesh": {k: round(v, 4) for k, v in pass_ms_mesh.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            sw, sh = W // 2, H // 2
            kind, cores, times = cpu_frame_seconds(sw, sh, 2, 1)
            sec = float(np.mean(times))
            line["cpu_baseline"] = {"value": sw * sh / sec / 1e6, "unit": "Mpix/s", "cores": cores, "kind": kind,
                                    "sample": f"{sw}x{sh} full frame (1/4 of the pixels), mean of 2 frames after 1 warm-up, {sec:.2f} s/frame"}
        print(json.dumps(line), flush=True)
    for r_i, _, _ in slots:
        r_i.close()
    if dist is not None:
        dist.destroy_process_group()



def resident_frame_ms(W: int, H: int, seed: int, steps: int, warmup: int = 5, gi_flags: int = 0):
    """ms per frame of the fused W x H frame on THIS rank's GPU with the rasterised scene resident (CUDA-graph replay, CUDA events)."""
    import torch

    from legitengine_b200 import abi, harness, scene

    m = scene.frame_matrices(W, H)
    frags = np.empty((H, W), dtype=abi.FRAGMENT_DTYPE)
    scene.scene_fragments(seed, W, H, m, out=frags)
    objects = scene.scene_objects(seed)
    shadow = scene.scene_shadow_map(seed, m)
    stream = torch.cuda.Stream()
    r = harness.Renderer(W, H, stream=stream.cuda_stream)
    r.upload_fragments(frags.ctypes.data, frags.strides[0])
    r.upload_objects(objects.ctypes.data, len(objects))
    r.upload_light_depth(np.ascontiguousarray(shadow).ctypes.data, 1024)
    r.sync()
    r.render_frame(harness.MODE_FUSED, 0, gi_flags)
    r.sync()
    r.capture_frame(harness.MODE_FUSED, 0, gi_flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            r.replay_frame()
        e0.record(stream)
        for _ in range(steps):
            r.replay_frame()
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    r.close()
    return ms


def batch_frames_ms(W: int, H: int, seeds, repeats: int = 3, gi_flags: int = 0):
    """BASELINE configs[4] on THIS rank: one resident rasterised scene + renderer (image set, captured frame) per seed in `seeds`;
    returns the ms this GPU needs to render every one of them once (best of `repeats` passes over the batch, CUDA events)."""
    import torch

    from legitengine_b200 import abi, harness, scene

    m = scene.frame_matrices(W, H)
    stream = torch.cuda.Stream()
    frags = np.empty((H, W), dtype=abi.FRAGMENT_DTYPE)
    renderers = []
    for seed in seeds:
        scene.scene_fragments(seed, W, H, m, out=frags)
        objects = scene.scene_objects(seed)
        shadow = np.ascontiguousarray(scene.scene_shadow_map(seed, m))
        r = harness.Renderer(W, H, stream=stream.cuda_stream)
        r.upload_fragments(frags.ctypes.data, frags.strides[0])
        r.upload_objects(objects.ctypes.data, len(objects))
        r.upload_light_depth(shadow.ctypes.data, 1024)
        r.sync()
        r.render_frame(harness.MODE_FUSED, 0, gi_flags)
        r.sync()
        r.capture_frame(harness.MODE_FUSED, 0, gi_flags)
        renderers.append(r)
    best = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(repeats + 1):  # first pass = warm-up
            e0.record(stream)
            for r in renderers:
                r.replay_frame()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
    for r in renderers:
        r.close()
    return best


def run_strips(args, rank: int, world: int, local_rank: int):
    """ONE frame of the workload cut into row strips over the ranks (legitengine_b200/multigpu.py): per step every rank renders
    its strip in stages with halo exchanges over NCCL/NVLink in between, then the swapchain strips are composited on rank 0."""
    import torch
    import torch.distributed as dist

    from legitengine_b200 import abi, multigpu, scene, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H = workload_size(args.workload)
    m = scene.frame_matrices(W, H)
    seed = 0xC0FFEE
    objects = scene.scene_objects(seed)
    shadow = torch.from_numpy(scene.scene_shadow_map(seed, m)).pin_memory()
    swap_host = torch.empty((H, W * 4), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.Stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cls = multigpu.P2PStripRenderer if args.transport == "p2p" else multigpu.StripRenderer
    gi_flags = abi.GI_STRICT if args.strict else abi.GI_DEFAULT

    scene_mesh = scene.scene_mesh(seed)
    mesh = scene_mesh if args.strip_input == "mesh" else None

    def build(bounds, mesh=mesh):
        """Strip renderer for `bounds` with this rank's strip of the rasterised scene generated into pinned memory and uploaded
        (or, with `mesh`, the scene in the reference's form: every rank then rasterises its own rows in the front stage)."""
        y0, y1 = bounds[rank]
        frag_host = torch.empty((max(y1 - y0, 1) if mesh is None else 1, W * 32), dtype=torch.uint8).pin_memory()
        frags = frag_host.numpy().view(abi.FRAGMENT_DTYPE).reshape(-1, W)
        if mesh is None:
            scene.scene_fragments(seed, W, H, m, rows=(y0, y1), out=_OffsetRows(frags, y0))
        extra = {"fused_exchange": not args.unfused_exchange} if args.transport == "p2p" else {}
        sr = cls(W, H, rank, world, dist, stream=stream.cuda_stream, present=not args.no_present, bounds=bounds, **extra)
        sr.renderer.upload_objects(objects.ctypes.data, len(objects))
        sr.renderer.upload_light_depth(shadow.data_ptr(), 1024)
        ptr = frag_host.data_ptr() - y0 * W * 32  # lgh_upload_fragments takes the address of row 0
        if mesh is not None:
            sr.renderer.upload_mesh(mesh)  # the front stage then rasterises this rank's rows itself
        else:
            sr.upload_strip(ptr, W * 32)
        return sr, frag_host, ptr

    def stage_profile(sr, frames=8):
        """Per-rank GPU time between the stage marks of un-captured frames, all-gathered: {stage: [ms of rank 0, 1, ...]}."""
        n_marks = len(sr.STAGE_MARKS)
        samples = []
        for i in range(frames + 2):
            marks = []
            sr.render(gi_flags, marks=marks)
            torch.cuda.synchronize()
            if i >= 2:
                samples.append([marks[j].elapsed_time(marks[j + 1]) for j in range(n_marks - 1)])
        acc = torch.tensor(np.median(np.asarray(samples), axis=0), device="cuda", dtype=torch.float32)  # median: robust against launch jitter
        dist.barrier()
        every = [torch.zeros_like(acc) for _ in range(world)]
        dist.all_gather(every, acc)
        return {name: [round(float(e[j]), 4) for e in every] for j, name in enumerate(sr.STAGE_MARKS[1:])}

    bounds = sharding.strip_bounds(H, world)
    if args.strip_bounds:
        cuts = [int(v) for v in args.strip_bounds.split(",")]
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == H, "--strip-bounds needs world+1 increasing cuts from 0 to the height"
        bounds = [(cuts[i], cuts[i + 1]) for i in range(world)]
        args.balance = -1
    balance_log = []
    with torch.cuda.stream(stream):
        sr, frag_host, full_view_ptr = build(bounds)
        # cost-aware strips (sharding.rebalance_bounds): measure every rank's own kernel time, move the boundaries, rebuild
        # cost-aware strips: measure every rank's own kernel time, refine the frame's per-block cost profile with it
        # (sharding.refine_cost_density), cut the profile into equal parts, rebuild; in the end keep the partition whose slowest rank was
        # fastest (the un-captured stage profile is noisy at the +-2 % level, so the last partition tried is not always the best one)
        density, tried = None, []
        for it in range((args.balance + 1) if (args.transport == "p2p" and args.balance >= 0) else 0):
            prof = stage_profile(sr, frames=10)
            # what a rank adds to the frame's critical path: everyone waits for the slowest front and chains stage anyway (level 4 of the
            # chains is gathered from every rank), so it is the gather + final stage that has to be equal; the front only counts through its
            # share of the slowest rank's stage
            own = [prof["gather_final"][r] + 0.25 * prof["front"][r] for r in range(world)]
            balance_log.append({"bounds": bounds, "own_ms": [round(v, 4) for v in own]})
            tried.append((max(own), bounds))
            if it == args.balance:
                new_bounds = min(tried)[1]  # last round: go back to the best partition seen
            else:
                density = sharding.refine_cost_density(density, bounds, own, H)
                new_bounds = sharding.bounds_from_density(density, world, H)
            if new_bounds == bounds:
                if it == args.balance or min(tried)[1] == bounds:
                    break
                new_bounds = min(tried)[1]
            torch.cuda.synchronize()
            sr.close()
            del frag_host
            bounds = new_bounds
            sr, frag_host, full_view_ptr = build(bounds)
            if it == args.balance:
                break

        def barrier():
            dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, steps, warmup, sample_clocks=False):
            for _ in range(warmup):
                fn()
            barrier()
            sampler = ClockSampler(local_rank) if sample_clocks else None
            if sampler:
                sampler.start()
            ev0.record(stream)
            for _ in range(steps):
                fn()
            ev1.record(stream)
            barrier()
            clocks = sampler.stop() if sampler else None
            t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), clocks

        warm = max(args.warmup, 3)
        use_graph = not args.no_graph
        if use_graph:
            sr.capture(gi_flags)
        frame = sr.replay if use_graph else (lambda: sr.render(gi_flags))
        ms_total, clocks = timed(frame, args.steps, warm, sample_clocks=True)
        ms_per_step = ms_total / args.steps
        received = sr.received_bytes

        recv_all = torch.tensor([received], device="cuda", dtype=torch.int64)
        dist.all_reduce(recv_all)
        # per-rank, per-stage GPU time of un-captured frames (events between the stages on every rank): shows where a strip waits
        stage_ms = stage_profile(sr) if args.transport == "p2p" else None

        # --- e2e: like the single-GPU line, the frame starts from the SCENE in host memory (vertex / index buffers, draw list, per-object
        # constants uploaded by every rank every step, rasterised for the rank's own rows on the device) and ends with the composited
        # swapchain image in the presenting rank's pinned host memory
        if mesh is None:
            torch.cuda.synchronize()
            sr.release_graph()
            dist.barrier()
            sr.close()
            del frag_host
            sr, frag_host, full_view_ptr = build(bounds, scene_mesh)
            if use_graph:
                sr.capture(gi_flags)
            frame = sr.replay if use_graph else (lambda: sr.render(gi_flags))

        def e2e_step():
            sr.renderer.upload_mesh(scene_mesh)
            frame()
            if rank == 0:
                sr.renderer.download_swapchain(swap_host.data_ptr(), W * 4)

        e2e_steps = args.e2e_steps or min(args.steps, 30)
        e2e_ms, _ = timed(e2e_step, e2e_steps, 3)
    # --- the same run's single-GPU frame of this workload (rank 0 alone; the others wait), and BASELINE configs[4]'s throughput mode:
    # every GPU renders independent 4K frames (same scene on every rank, so that the MAX over ranks measures the GPUs, not the scenes)
    extra_steps = max(10, min(args.steps, 50))
    torch.cuda.synchronize()
    dist.barrier()
    if args.skip_extras:
        if rank == 0:
            print(json.dumps({"ms_per_step": ms_per_step, "value": W * H / (ms_per_step * 1e-3) / 1e6, "n_gpus": world, "strips": bounds, "fused_exchange": not args.unfused_exchange,
                              "stage_ms_per_rank": stage_ms}), flush=True)
        torch.cuda.synchronize()
        sr.release_graph()
        dist.barrier()
        sr.close()
        dist.destroy_process_group()
        return
    single_ms = resident_frame_ms(W, H, seed, extra_steps, gi_flags=gi_flags) if rank == 0 else 0.0
    dist.barrier()
    rw, rh = WORKLOADS["4k"]
    rep = torch.tensor([resident_frame_ms(rw, rh, seed, extra_steps, gi_flags=gi_flags)], device="cuda")
    rep_all = [torch.zeros_like(rep) for _ in range(world)]
    dist.all_gather(rep_all, rep)
    rep_ms = [float(t.item()) for t in rep_all]
    # BASELINE configs[4] as specified: a batch of 64 DISTINCT synthetic 4K frames (seeds 0..63) over the GPUs, G-buffers resident
    batch_total = 8 * world  # 64 on the 8 GPUs BASELINE names; 8 per GPU on smaller runs so that they stay short
    my_seeds = list(range(rank, batch_total, world))
    dist.barrier()
    bt = torch.tensor([batch_frames_ms(rw, rh, my_seeds, gi_flags=gi_flags)], device="cuda")
    bt_all = [torch.zeros_like(bt) for _ in range(world)]
    dist.all_gather(bt_all, bt)
    batch_ms = [float(t.item()) for t in bt_all]
    if rank == 0:
        npx = W * H
        peak, peak_src = measured_peak_gbs()
        frame_bytes = FRAME_BYTES_PER_PX * npx + SHADOW_MAP_BYTES * world
        fused_exchange = args.transport == "p2p" and not args.unfused_exchange
        line = {
            "metric": METRIC, "value": npx / (ms_per_step * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{W}x{H} full GI frame tile-sharded into {world} row strips with NVLink halo exchange (BASELINE configs[3])",
                "pass_list": "fused stages: front | exchange | chains | exchange | gather + final | composite on rank 0",
                "strips": bounds, "present": not args.no_present, "cuda_graph": use_graph, "strip_input": args.strip_input,
                "transport": "own copy + flag kernels over NVLink peer memory (CUDA IPC)" if args.transport == "p2p" else "NCCL send/recv per slab",
                "exchange_bytes_per_frame_all_ranks": int(recv_all.item()),
                "l2": "inputs larger than L2 (per-GPU strip working set > 126 MB at 8K / 8 GPUs)",
            },
            "clocks": clocks,
            "e2e": {"value": npx / (e2e_ms / e2e_steps * 1e-3) / 1e6, "unit": "Mpix/s",
                    "h2d_bytes_per_step": world * sum(a.nbytes for a in (scene_mesh.vertices, scene_mesh.indices, scene_mesh.draws, scene_mesh.objects)),
                    "d2h_bytes_per_step": W * H * 4,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "input": "scene as the reference holds it, uploaded by every rank from host memory every step and rasterised for the rank's own rows on the "
                             "device; output: the composited BGRA8 swapchain image from the presenting rank to pinned host memory (PCIe-bound at 8K: 133 MB per frame)"},
            # per rank and frame: frame front, chains (tail clusters + blur grid, one launch), side-pyramid pack, gather, K6+K7 — plus, with the
            # fused exchange, the four lgcu_exchange kernels (open, after front, after chains, close); the unfused / NCCL transports launch
            # more (separate signal, wait and copy kernels) and are counted by their compute kernels only
            "gpu_launches": (9 if fused_exchange else 5) * args.steps * world,
            "kernels_per_frame": 9 if fused_exchange else 5,
            "single_gpu": {"ms_per_step": single_ms, "value": npx / (single_ms * 1e-3) / 1e6, "unit": "Mpix/s", "steps": extra_steps,
                           "speedup_of_this_line": single_ms / ms_per_step,
                           "note": f"the whole {W}x{H} frame on rank 0's GPU alone, measured in this run after the strip-sharded leg"},
            "replicas": {"value": world * rw * rh / (max(rep_ms) * 1e-3) / 1e6, "unit": "Mpix/s", "scaling": "weak", "ms_per_step_per_rank": [round(v, 4) for v in rep_ms],
                         "frames_per_s": world / (max(rep_ms) * 1e-3), "steps": extra_steps,
                         "note": f"BASELINE configs[4] throughput mode: every GPU renders independent {rw}x{rh} frames, no data-path collective; MAX over ranks"},
            "throughput_batch": {"frames": batch_total, "frames_per_s": batch_total / (max(batch_ms) * 1e-3), "value": batch_total * rw * rh / (max(batch_ms) * 1e-3) / 1e6,
                                 "unit": "Mpix/s", "ms_per_rank": [round(v, 3) for v in batch_ms], "frames_per_rank": len(my_seeds),
                                 "note": f"BASELINE configs[4]: {batch_total} distinct synthetic {rw}x{rh} frames (seeds 0..{batch_total - 1}, random geometry / albedo / "
                                         f"emissive), seed s on rank s mod {world}, rasterised scenes resident, one captured frame per scene; time = MAX over ranks"},
            "stage_ms_per_rank": stage_ms,
            "balance": balance_log,
            "roofline_frame": {"bound": "hbm", "achieved": frame_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                               "frac": frame_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world), "algorithmic_bytes": frame_bytes,
                               "note": "whole frame over all GPUs, pass-granular algorithmic bytes; peak = N x measured single-GPU HBM copy bandwidth"},
        }
        print(json.dumps(line), flush=True)
    # tear-down: release the captured graph before the communicator; if NCCL still refuses to shut down cleanly, leave anyway
    torch.cuda.synchronize()
    sr.release_graph()
    dist.barrier()
    sys.stdout.flush()
    if use_graph and args.transport == "nccl":
        os._exit(0)  # a process group whose transfers were graph-captured can block in destroy_process_group
    sr.close()
    dist.destroy_process_group()


class _OffsetRows:
    """Minimal stand-in for a full-height fragment array whose storage starts at row y0 (only .ctypes.data / .strides are used)."""

    def __init__(self, frags, y0):
        self._frags, self._y0 = frags, y0
        self.strides = frags.strides

        class _C:
            data = frags.ctypes.data - y0 * frags.strides[0]

        self.ctypes = _C()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    shard = args.shard or ("strips" if world > 1 else "replicas")
    if world > 1 and shard == "strips":
        args.workload = args.workload or "8k"
        run_strips(args, rank, world, local_rank)
        return
    args.workload = args.workload or "4k"
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
