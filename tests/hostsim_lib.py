"""ctypes front end of tests/hostsim: the DEVICE header pvb_device.cuh compiled for the host (g++, see
tests/hostsim/cuda_runtime.h), so the per-thread arithmetic of the kernels can be checked in the CPU tier.
Test infrastructure only; descriptors carry host pointers (numpy / CPU torch memory)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

from pytorch_volumetric_b200 import _native as nat

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "_build", "libpvb_hostsim.so")
_SOURCES = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(_HERE, "cuda_runtime.h"),
            os.path.join(_ROOT, "pytorch_volumetric_b200", "csrc", "pvb_device.cuh"),
            os.path.join(_ROOT, "include", "pvb.h")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in _SOURCES)
        if stale:
            os.makedirs(os.path.dirname(_LIB), exist_ok=True)
            # -ffp-contract=off: every fp32 operation rounds once, like the *_rn intrinsics the header relies on
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
                            "-I", _HERE, os.path.join(_HERE, "hostsim.cpp"), "-o", _LIB], check=True)
        _lib = ctypes.CDLL(_LIB)
        _lib.sim_hash_normal.restype = ctypes.c_float
        _lib.sim_hash_normal.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32]
    return _lib


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _pts(points):
    return np.ascontiguousarray(points.detach().cpu().numpy() if torch.is_tensor(points) else points, dtype=np.float32)


def grid_lookup(desc, points, branchy=False):
    p = _pts(points); n = len(p)
    val, grad, key = np.empty(n, np.float32), np.empty((n, 3), np.float32), np.empty(n, np.int64)
    if branchy:
        lib().sim_grid_lookup_branchy(ctypes.byref(desc), _p(p), ctypes.c_longlong(n), _p(val), _p(grad))
        return val, grad
    lib().sim_grid_lookup(ctypes.byref(desc), _p(p), ctypes.c_longlong(n), _p(val), _p(grad), _p(key))
    return val, grad, key


def mesh_query(desc, points, mode=nat.PVB_MESH_DEFAULT, stats=False):
    """(dist, grad, closest, face) and, with stats=True, a dict of traversal counts per query."""
    p = _pts(points); n = len(p)
    dist, grad = np.empty(n, np.float32), np.empty((n, 3), np.float32)
    closest, face = np.empty((n, 3), np.float32), np.empty(n, np.int32)
    st = np.zeros(4, np.int64)
    lib().sim_mesh_query(ctypes.byref(desc), _p(p), ctypes.c_longlong(n), ctypes.c_uint32(mode), _p(dist), _p(grad),
                         _p(closest), _p(face), _p(st))
    if stats:
        names = ("closest_nodes", "closest_tris", "parity_nodes", "parity_tris")
        return dist, grad, closest, face, {k: float(v) / max(n, 1) for k, v in zip(names, st)}
    return dist, grad, closest, face


def parity(desc, points, dirs=None):
    p = _pts(points); n = len(p)
    px = np.empty(n, np.int32)
    pr = np.empty(n, np.int32) if dirs is not None else None
    d = _pts(dirs) if dirs is not None else None
    lib().sim_parity(ctypes.byref(desc), _p(p), ctypes.c_longlong(n), _p(d), _p(px), _p(pr))
    return px, pr


def winding(desc, points):
    p = _pts(points); n = len(p)
    w = np.empty(n, np.float32)
    lib().sim_winding(ctypes.byref(desc), _p(p), ctypes.c_longlong(n), _p(w))
    return w


def bit_reversed_order(n):
    """The visiting order the composed kernels use (pvb_kernels.cu fill_order)."""
    bits = max(0, (n - 1).bit_length())
    out = []
    for i in range(1 << bits):
        r = int(format(i, f"0{bits}b")[::-1], 2) if bits else 0
        if r < n:
            out.append(r)
    return out


def composed(descs, xforms, n_cfg, points, order=None, mesh_mode=nat.PVB_MESH_DEFAULT):
    """(val [n_cfg*P], grad [n_cfg*P,3], which [n_cfg*P]) of the composition of `descs` under xforms
    [(S*n_cfg),4,4] (sub-SDF-major), visiting the sub-SDFs in `order` (default: the kernels' bit-reversed order)."""
    p = _pts(points); n = len(p); S = len(descs)
    arr = (nat.SdfDesc * S)(*descs)
    xf = np.ascontiguousarray(xforms.detach().cpu().numpy() if torch.is_tensor(xforms) else xforms, dtype=np.float32)
    assert xf.shape == (S * n_cfg, 4, 4)
    order = bit_reversed_order(S) if order is None else list(order)
    assert sorted(order) == list(range(S))
    od = np.asarray(order, dtype=np.uint8)
    val, grad = np.empty(n_cfg * n, np.float32), np.empty((n_cfg * n, 3), np.float32)
    which = np.empty(n_cfg * n, np.int32)
    lib().sim_composed(arr, ctypes.c_int(S), _p(od), _p(xf), ctypes.c_int(n_cfg), _p(p), ctypes.c_longlong(n),
                       ctypes.c_uint32(mesh_mode), _p(val), _p(grad), _p(which))
    return val, grad, which


def sphere_desc(radius):
    import pytorch_volumetric_b200 as pv
    return pv.SphereSDF(radius).native_desc("cpu")      # the product's own descriptor; no device memory involved


def sphere(radius, points):
    p = _pts(points); n = len(p)
    val, grad = np.empty(n, np.float32), np.empty((n, 3), np.float32)
    lib().sim_sphere(ctypes.c_float(radius), _p(p), ctypes.c_longlong(n), _p(val), _p(grad))
    return val, grad


# ---------------------------------------------------------------- descriptors over host memory
def mesh_desc(obj, with_winding=False):
    """pvb_sdf_desc (MESH) of a MeshObjectFactory with the BVH / triangles / normals in host memory, filled by the
    product's own `_fill_mesh_part`.  Returns (desc, keepalive)."""
    from pytorch_volumetric_b200.sdf import _winding_moments
    nodes, tris, _ = obj._bvh_host
    st = {"nodes": torch.from_numpy(np.ascontiguousarray(nodes)), "tris": torch.from_numpy(np.ascontiguousarray(tris)),
          "fn32": torch.from_numpy(obj._face_normals.astype(np.float32)).contiguous()}
    d = nat.SdfDesc()
    d.kind = nat.PVB_KIND_MESH
    d.flags = 0
    obj._fill_mesh_part(d, st)
    if with_winding:
        st["wn"] = torch.from_numpy(np.ascontiguousarray(_winding_moments(nodes, tris)))
        d.wn_nodes = st["wn"].data_ptr()
    return d, st


def grid_desc(table_val, table_grad, ranges, bb, strategy=None, gt_obj=None):
    """pvb_sdf_desc (GRID) built by the product's CachedSDF._make_desc over CPU tensors.  Returns (desc, keepalive)."""
    from pytorch_volumetric_b200 import sdf as S
    from pytorch_volumetric_b200.voxel import GridView
    c = object.__new__(S.CachedSDF)
    c.ranges = ranges
    c.voxels = GridView(table_val, ranges, invalid_value=None)
    c.voxels_grad = table_grad
    c._table = torch.cat([table_val.reshape(-1, 1), table_grad], 1).contiguous()
    c.bb = torch.as_tensor(np.asarray(bb))
    c._cdev = torch.device("cpu")
    c.out_of_bounds_strategy = S.OutOfBoundsStrategy.BOUNDING_BOX
    c.interpolation = "nearest"
    d = c._make_desc()
    keep = [c]
    if strategy == S.OutOfBoundsStrategy.LOOKUP_GT_SDF:      # what _make_desc does when the ground truth is native
        bbf = np.asarray(bb, dtype=np.float32)
        md, st = mesh_desc(gt_obj)
        keep.append(st)
        gt_obj._fill_mesh_part(d, st)
        d.flags = (d.flags | nat.PVB_GRID_OOB_GT | (md.flags & nat.PVB_MESH_CLOSED)) & ~nat.PVB_GRID_PRUNE_OK
        for k in range(3):
            d.bb_min[k] = float(bbf[k, 0]); d.bb_max[k] = float(bbf[k, 1])
    return d, keep
