"""TEST INFRASTRUCTURE: the reference's bundled sample scene (BASELINE configs[0]: bin/data/Scenes/SponzaScene.json with the OBJ meshes
under bin/data/Meshes) as a legitengine_b200.scene.Mesh, loaded from /root/reference where it lies (development container only).

Follows the reference's loaders: src/Scene/Scene.h:45-113 (meshes + objects of the scene file; objToWorld = translate(pos) [* rotate],
albedoColor / emissiveColor default to 0) and src/Scene/Mesh.h:13-60 (one vertex per face corner: pos * scale, the OBJ normal or
(1,0,0), the OBJ uv or (0,0); polygons triangulated as a fan, which is what tinyobj does for the triangles and convex quads these
files contain). The reference de-duplicates identical corners into an index buffer; that changes no triangle, so this loader keeps one
vertex per corner and an identity index buffer. Camera / light are the application defaults (src/main.cpp:166-172)."""
from __future__ import annotations

import json
import re
from pathlib import Path

import numpy as np

from legitengine_b200 import abi, scene

REFERENCE = Path("/root/reference")
SCENE_FILE = REFERENCE / "bin" / "data" / "Scenes" / "SponzaScene.json"


def available() -> bool:
    return SCENE_FILE.exists()


def _load_json_with_comments(path: Path):
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # jsoncpp accepts C comments (the scene file uses them to disable entries)
    text = re.sub(r"//[^\n]*", "", text)
    return json.loads(text)


def load_obj(path: Path, scale) -> np.ndarray:
    """-> abi.VERTEX_DTYPE array, three consecutive vertices per triangle (MeshData(filename, scale), Mesh.h:13-60)."""
    pos, nrm, uv, corners = [], [], [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            if line.startswith("v "):
                pos.append(line.split()[1:4])
            elif line.startswith("vn "):
                nrm.append(line.split()[1:4])
            elif line.startswith("vt "):
                uv.append(line.split()[1:3])
            elif line.startswith("f "):
                face = []
                for tok in line.split()[1:]:
                    parts = tok.split("/")
                    vi = int(parts[0])
                    ti = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    face.append((vi, ti, ni))
                for k in range(1, len(face) - 1):  # triangle fan
                    corners += [face[0], face[k], face[k + 1]]
    P = np.asarray(pos, dtype=np.float32)
    N = np.asarray(nrm, dtype=np.float32) if nrm else np.zeros((0, 3), np.float32)
    T = np.asarray(uv, dtype=np.float32) if uv else np.zeros((0, 2), np.float32)
    c = np.asarray(corners, dtype=np.int64)
    fix = lambda idx, n: np.where(idx < 0, idx + n, idx - 1)  # OBJ indices are 1-based; negative = relative to the end
    out = np.zeros(len(c), dtype=abi.VERTEX_DTYPE)
    out["pos"] = P[fix(c[:, 0], len(P))] * np.asarray(scale, dtype=np.float32)
    has_n, has_t = c[:, 2] != 0, c[:, 1] != 0
    out["normal"] = np.array([1.0, 0.0, 0.0], dtype=np.float32)
    if has_n.any():
        out["normal"][has_n] = N[fix(c[has_n, 2], len(N))]
    if has_t.any():
        out["uv"][has_t] = T[fix(c[has_t, 1], len(T))]
    return out


def _translate(p) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = np.asarray(p, dtype=np.float32)
    return m.T.reshape(16).copy()  # column-major like glm


def load_bundled_scene() -> scene.Mesh:
    cfg = _load_json_with_comments(SCENE_FILE)["scene"]
    meshes = {}
    for node in cfg["meshes"]:
        rel = node["filename"].replace("../data/", "")
        meshes[node["name"]] = load_obj(REFERENCE / "bin" / "data" / rel, node.get("scale", [1.0, 1.0, 1.0]))
    vertices, draws, objects = [], [], []
    first_vertex = {}
    for name, v in meshes.items():
        first_vertex[name] = sum(len(x) for x in vertices)
        vertices.append(v)
    vertices = np.concatenate(vertices)
    for node in cfg["objects"]:
        if node.get("mesh") not in meshes:
            continue  # Scene.h:88-92
        angle = node.get("angle")
        if angle is not None and float(np.linalg.norm(angle)) > 1e-3:
            raise NotImplementedError("rotated objects: not used by the bundled scene")
        obj = np.zeros((), dtype=abi.DRAW_CALL_DTYPE)
        obj["modelMatrix"] = _translate(node.get("pos", [0.0, 0.0, 0.0]))
        obj["albedoColor"] = list(node.get("albedoColor", [0.0, 0.0, 0.0])) + [0.0]   # ReadJsonVec3f of a missing node = 0; vec4(color, 0)? see below
        obj["emissiveColor"] = list(node.get("emissiveColor", [0.0, 0.0, 0.0])) + [0.0]
        # SSVGIRenderer.h:146-147: drawCallData->albedoColor = glm::vec4(albedoColor, 1.0f); emissiveColor likewise
        obj["albedoColor"][3] = 1.0
        obj["emissiveColor"][3] = 1.0
        n = len(meshes[node["mesh"]])
        draw = np.zeros((), dtype=abi.DRAW_DTYPE)
        draw["firstIndex"], draw["indexCount"], draw["vertexOffset"], draw["objectId"] = first_vertex[node["mesh"]], n, 0, len(objects)
        draws.append(draw)
        objects.append(obj)
    draws = np.array(draws, dtype=abi.DRAW_DTYPE)
    tri = 0
    for d in draws:  # lgcu_raster_prepare_draws
        d["firstTriangle"] = tri
        tri += int(d["indexCount"]) // 3
    indices = np.arange(len(vertices), dtype=np.uint32)
    return scene.Mesh(vertices, indices, draws, np.array(objects, dtype=abi.DRAW_CALL_DTYPE))
