"""Gather parity diagnosis at full size: the fast and the strict kernel against the port oracle on 16-row strips of the bench frame.

Prints one JSON line per (size, kernel, strip) with the outlier fraction at the test bar (max(1e-3, one fp16 ulp)), max-abs, the
99.99th percentile and PSNR, and saves the rows of every strip (both kernels, the oracle, centre depth) to
gpurun_out/diag_gather_<W>.npz for offline analysis.  Usage: python scripts/diag_gather.py [4k] [8k]
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from legitengine_b200 import abi, harness  # noqa: E402
from tests import helpers as H  # noqa: E402

SIZES = {"4k": (3840, 2160), "8k": (7680, 4320), "1080p": (1920, 1080)}


def strips_for(h: int):
    mid = (h // 2) & ~15
    cand = [0, h // 4 & ~15, mid - 64, mid - 32, mid - 16, mid, mid + 16, mid + 32, mid + 64, (3 * h // 4) & ~15, h - 16]
    return tuple((y, y + 16) for y in sorted(set(cand)))


def stats(a, b):
    d = np.abs(a - b)
    tol = np.maximum(1e-3, H.F16_EPS * np.abs(b))
    out = (d > tol).any(axis=2)
    return {"outside": float(out.mean()), "max_abs": float(d.max()), "p9999": float(np.quantile(d.max(axis=2), 0.9999)),
            "psnr": float(H.psnr(a, b, max(1.0, float(np.abs(b).max()))))}


def main():
    names = [a for a in sys.argv[1:] if a in SIZES] or ["4k"]
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    for name in names:
        W, Hh = SIZES[name]
        strips = strips_for(Hh)
        sc, p, ref = H.oracle_frame_on_strips(0xC0FFEE, W, Hh, strips)
        r = harness.Renderer(W, Hh)
        r.upload_scene(sc)
        dump = {"strips": np.array(strips)}
        want = ref.indirectLight.level_f32(0)[..., :3]
        dump["depth"] = np.concatenate([ref.depthStencil.level_f32(0)[y0:y1, :, 0] for y0, y1 in strips])
        dump["oracle"] = np.concatenate([want[y0:y1] for y0, y1 in strips])
        for label, flags in (("fast", abi.GI_DEFAULT), ("strict", abi.GI_STRICT)):
            r.render_frame(harness.MODE_FUSED, 0, flags)
            r.sync()
            got = r.download_image("indirectLight").level_f32(0)[..., :3]
            dump[label] = np.concatenate([got[y0:y1] for y0, y1 in strips])
            for y0, y1 in strips:
                print(json.dumps({"size": name, "kernel": label, "rows": [y0, y1], **stats(got[y0:y1], want[y0:y1])}), flush=True)
        r.close()
        np.savez_compressed(out_dir / f"diag_gather_{W}.npz", **dump)


if __name__ == "__main__":
    main()
