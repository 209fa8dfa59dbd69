"""ctypes binding of oracle/liborc_geom.so (see geom.c).

TEST INFRASTRUCTURE ONLY: the checker for tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Product code under
pytorch_volumetric_b200/ never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc_geom.so")
_lib = None


def build(force=False):
    """Compile geom.c with the recipe in oracle/Makefile (idempotent)."""
    src = os.path.join(_HERE, "geom.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liborc_geom.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        i64 = ctypes.c_int64
        L.orc_closest_points_brute.argtypes = [fp, ip, i64, fp, i64, fp, fp, ip]
        L.orc_closest_points_brute.restype = None
        L.orc_count_intersections_brute.argtypes = [fp, ip, i64, fp, i64, ip]
        L.orc_count_intersections_brute.restype = None
        L.orc_bvh_create.argtypes = [fp, ip, i64]
        L.orc_bvh_create.restype = ctypes.c_void_p
        L.orc_bvh_destroy.argtypes = [ctypes.c_void_p]
        L.orc_bvh_destroy.restype = None
        L.orc_closest_points_bvh.argtypes = [ctypes.c_void_p, fp, i64, fp, fp, ip]
        L.orc_closest_points_bvh.restype = None
        L.orc_count_intersections_bvh.argtypes = [ctypes.c_void_p, fp, i64, ip]
        L.orc_count_intersections_bvh.restype = None
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        L.orc_set_num_threads.restype = None
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


class TriangleSoup:
    """fp32 vertices + int32 faces with the two Embree-style queries.

    method='brute' is the parity checker; method='bvh' is the timing baseline.
    """

    def __init__(self, verts, faces):
        self.verts = np.ascontiguousarray(verts, dtype=np.float32)
        self.faces = np.ascontiguousarray(faces, dtype=np.int32)
        self._bvh = None

    def __del__(self):
        if self._bvh is not None and _lib is not None:
            _lib.orc_bvh_destroy(self._bvh)
            self._bvh = None

    def _handle(self):
        if self._bvh is None:
            self._bvh = lib().orc_bvh_create(_f(self.verts), _i(self.faces), len(self.faces))
        return self._bvh

    def closest_points(self, pts, method="brute"):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
        n = len(pts)
        closest = np.empty((n, 3), np.float32)
        d2 = np.empty(n, np.float32)
        face = np.empty(n, np.int32)
        if method == "brute":
            lib().orc_closest_points_brute(_f(self.verts), _i(self.faces), len(self.faces),
                                           _f(pts), n, _f(closest), _f(d2), _i(face))
        else:
            lib().orc_closest_points_bvh(self._handle(), _f(pts), n, _f(closest), _f(d2), _i(face))
        return closest, d2, face

    def count_intersections(self, rays, method="brute"):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = len(rays)
        cnt = np.empty(n, np.int32)
        if method == "brute":
            lib().orc_count_intersections_brute(_f(self.verts), _i(self.faces), len(self.faces),
                                                _f(rays), n, _i(cnt))
        else:
            lib().orc_count_intersections_bvh(self._handle(), _f(rays), n, _i(cnt))
        return cnt


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """OpenMP threads of the evaluators from now on."""
    lib().orc_set_num_threads(int(n))
