"""Triangle-mesh file readers (OBJ, STL) returning (vertices fp64 [V,3], faces int32 [F,3]).

Stands in for `o3d.io.read_triangle_mesh` at the reference's
src/pytorch_volumetric/sdf.py:103.  Only positions and faces are read
(texture / normal indices are ignored); polygons are fan-triangulated.
"""
import struct

import numpy as np


def read_obj(path):
    """Positions (`v x y z [...]`) and faces (`f a[/t[/n]] ...`, 1-based or negative = relative to the vertices read
    so far); polygons become triangle fans; every other record (vt, vn, g, o, s, usemtl, comments) is skipped."""
    verts = []
    faces = []
    with open(path, "r", errors="ignore") as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append((float(t[1]), float(t[2]), float(t[3])))
            elif t[0] == "f":
                corner = []
                for tok in t[1:]:
                    k = int(tok.partition("/")[0])
                    corner.append(k - 1 if k > 0 else len(verts) + k)
                for j in range(2, len(corner)):
                    faces.append((corner[0], corner[j - 1], corner[j]))
    v = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    f = np.asarray(faces, dtype=np.int32).reshape(-1, 3)
    if len(f) and (f.min() < 0 or f.max() >= len(v)):
        raise RuntimeError(f"OBJ face refers to a vertex that does not exist: {path}")
    return v, f


def read_stl(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    head = raw[:512].lstrip()
    if head.startswith(b"solid") and b"facet" in raw[:4096]:
        coords = []
        for line in raw.decode(errors="ignore").splitlines():
            s = line.strip()
            if s.startswith("vertex"):
                t = s.split()
                coords.append((float(t[1]), float(t[2]), float(t[3])))
        v = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
    else:
        (n,) = struct.unpack_from("<I", raw, 80)
        rec = np.frombuffer(raw, dtype=np.dtype([("normal", "<f4", (3,)), ("v", "<f4", (9,)), ("attr", "<u2")]),
                            count=n, offset=84)
        v = rec["v"].reshape(-1, 3).astype(np.float64)
    f = np.arange(len(v), dtype=np.int32).reshape(-1, 3)
    return v, f


def read_triangle_mesh(path):
    lower = path.lower()
    if lower.endswith(".stl"):
        v, f = read_stl(path)
    elif lower.endswith(".obj"):
        v, f = read_obj(path)
    else:
        raise RuntimeError(f"Unsupported mesh format (OBJ and STL are supported): {path}")
    if len(f) == 0:
        raise RuntimeError(f"Mesh file has no faces: {path}")
    return v, f


def write_obj(path, vertices, faces):
    with open(path, "w") as fh:
        for p in np.asarray(vertices, dtype=np.float64):
            fh.write(f"v {p[0]:.17g} {p[1]:.17g} {p[2]:.17g}\n")
        for t in np.asarray(faces):
            fh.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")


def quaternion_xyzw_to_matrix(q):
    x, y, z, w = (float(c) for c in q)
    n = (x * x + y * y + z * z + w * w) ** 0.5
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def is_closed_manifold(faces):
    """True iff every directed edge (a,b) is matched by exactly one (b,a): a closed, consistently
    oriented surface.  Ray parity from outside the AABB is then always even, which is what lets the
    composed kernels skip the sign pass for points outside a sub-mesh's box."""
    f = np.asarray(faces, dtype=np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    if np.any(e[:, 0] == e[:, 1]):
        return False
    nv = int(f.max()) + 1
    fwd = e[:, 0] * nv + e[:, 1]
    rev = e[:, 1] * nv + e[:, 0]
    uf, cf = np.unique(fwd, return_counts=True)
    if np.any(cf != 1):
        return False
    return bool(np.array_equal(uf, np.unique(rev)))
