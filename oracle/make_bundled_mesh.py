"""TEST / BENCH INFRASTRUCTURE. Packs the reference's bundled sample scene (BASELINE configs[0]: bin/data/Scenes/SponzaScene.json + the OBJ
meshes, read where they lie under /root/reference by tests/bundled_scene.py) into oracle/_ref/bundled_sponza_mesh.npz: the scene in the
reference's own form — de-duplicated vertex buffer (MeshData::Vertex, src/Scene/Mesh.h:209-216), uint32 index buffer, draw list,
per-object constants. Like oracle/_ref/libref_spirv.so the file is derived from the reference, git-ignored, and travels to the GPU box
with the snapshot; /root/reference itself does not exist there.  python oracle/make_bundled_mesh.py"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "oracle" / "_ref" / "bundled_sponza_mesh.npz"


def main() -> None:
    from tests import bundled_scene as B

    if not B.available():
        print("make_bundled_mesh: /root/reference not present - keeping", OUT if OUT.exists() else "nothing")
        return
    mesh = B.load_bundled_scene()
    v = mesh.vertices
    # the reference de-duplicates identical face corners into an index buffer (Mesh.h:39-58); per draw, so that vertexOffset stays 0
    _, first, inverse = np.unique(v.view(np.dtype((np.void, v.dtype.itemsize))), return_index=True, return_inverse=True)
    order = np.argsort(first)  # keep first-seen order
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    vertices = v[first[order]]
    indices = rank[inverse.reshape(-1)][mesh.indices].astype(np.uint32)
    assert np.array_equal(vertices[indices].view(np.uint8), v[mesh.indices].view(np.uint8))
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, vertices=vertices.view(np.uint8), indices=indices, draws=mesh.draws.view(np.uint8), objects=mesh.objects.view(np.uint8))
    print(f"make_bundled_mesh: {len(vertices)} vertices, {len(indices) // 3} triangles, {len(mesh.draws)} draws -> {OUT} ({OUT.stat().st_size / 1e6:.1f} MB)")


def load():
    """-> legitengine_b200.scene.Mesh, or None when the file has not been generated."""
    from legitengine_b200 import scene

    return scene.load_packed_mesh(OUT) if OUT.exists() else None


if __name__ == "__main__":
    main()
