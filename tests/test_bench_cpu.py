"""CPU tests of bench.py's contract: the reference arm (`--impl reference`: the reference's own SPIR-V passes, or the C port, on the
host cores) prints ONE JSON line with the keys the driver reads, and the product arm refuses to run without a CUDA device."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_json_line():
    res = _run("--impl", "reference", "--workload", "512x288", "--steps", "2", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "full_gi_frame_mpix_per_s" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly(monkeypatch):
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the driver itself
    res = _run("--steps", "1", "--warmup", "0")
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)
