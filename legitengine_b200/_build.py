"""In-tree build of every native piece (explicit nvcc / g++ / make; outputs under legitengine_b200/lib and oracle/).

  liblgcu.so        CUDA kernels + C ABI, sm_100a only (nvcc cross-compiles without a GPU)
  liblgcu_scene.so  host-only synthetic scene + frame maths
  oracle/...        the CPU oracles (test infrastructure; building the checker is not using it)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
LIB = PKG / "lib"
OBJ = LIB / "obj"

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-Xptxas", "-v"]

# translation unit -> extra flags. "exact" units evaluate the shaders' fp32 expressions without FMA contraction.
CUDA_UNITS = {
    "k_streaming.cu": ["-fmad=false"],
    "k_chain.cu": ["-fmad=false"],
    "k_front.cu": ["-fmad=false"],
    "k_gather_strict.cu": ["-fmad=false"],
    "k_gather_fast.cu": [],
    "k_p2p.cu": [],
    "k_raster.cu": ["-fmad=false"],
    "k_aux.cu": ["-fmad=false"],
    "lgcu_api.cu": ["-fmad=false"],
    "lgcu_interop.cu": [],
}


def _run(cmd, log=None, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    res = subprocess.run([str(c) for c in cmd], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if log is not None:
        log.write(res.stdout)
    if res.returncode != 0:
        sys.stdout.write(res.stdout)
        raise RuntimeError(f"build step failed ({res.returncode}): {' '.join(str(c) for c in cmd)}")
    return res.stdout


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_cuda(force: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    OBJ.mkdir(exist_ok=True)
    # this file is a dependency too: a changed flag in CUDA_UNITS must rebuild the unit
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "lgcu.h", ROOT / "include" / "lgcu_interop.h", Path(__file__)]
    objs = []
    with open(LIB / "nvcc_ptxas.log", "a") as log:  # ptxas -v output (registers, spills) of the units compiled by this call; git-ignored
        for unit, extra in CUDA_UNITS.items():
            src, obj = CSRC / unit, OBJ / (unit + ".o")
            if force or _stale(obj, [src] + headers):
                log.write(f"==== {unit}\n")
                _run([NVCC] + NVCC_COMMON + extra + ["-c", src, "-o", obj], log=log)
            objs.append(obj)
        out = LIB / "liblgcu.so"
        if force or _stale(out, objs):
            _run([NVCC] + ARCH + ["-shared", "-o", out] + objs, log=log)
    return out


def build_scene(force: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    out = LIB / "liblgcu_scene.so"
    src = HOST / "synth_scene.cpp"
    deps = [src, HOST / "legit_cuda" / "Camera.h", CSRC / "lgcu_mat4.h", ROOT / "include" / "lgcu.h"]
    if force or _stale(out, deps):
        _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-fopenmp", "-shared", f"-I{HOST}", "-o", out, src])
    return out


def build_host(force: bool = False) -> Path:
    """liblegit_cuda.so: the C++ rendergraph / renderer mirror + headless harness (host code; links liblgcu.so)."""
    LIB.mkdir(exist_ok=True)
    out = LIB / "liblegit_cuda.so"
    src = HOST / "harness.cpp"
    deps = [src, ROOT / "include" / "lgcu.h", ROOT / "include" / "lgcu_harness.h", LIB / "liblgcu.so"] + list((HOST / "legit_cuda").glob("*.h")) + [CSRC / "lgcu_mat4.h"]
    if force or _stale(out, deps):
        _run([NVCC, "-shared", "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", f"-I{HOST}", "-o", out, src,
              f"-L{LIB}", "-llgcu", "-Xlinker", "-rpath=$ORIGIN"])
    return out


def build_oracle() -> None:
    env = dict(os.environ)
    _run(["make", "-C", ROOT / "oracle", "port"], env=env)
    if Path("/root/reference/dependencies/spirv-cross/spirv_cpp.cpp").exists():
        _run(["make", "-C", ROOT / "oracle", "-j8", "ref"], env=env)
        packed = ROOT / "oracle" / "_ref" / "bundled_sponza_mesh.npz"
        if not packed.exists():  # BASELINE configs[0]: the reference's bundled scene, packed for the GPU box (git-ignored like the .so above)
            _run([sys.executable, ROOT / "oracle" / "make_bundled_mesh.py"], env=env)


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)
    build_scene(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
