"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import os

import numpy as np
import torch

import workloads
from oracle import port

GOLDEN = workloads.GOLDEN


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def port_mesh(name):
    v, f = workloads.fixture_mesh(name)
    return port.MeshPort(vertices=v, faces=f, name=name)


def pv_factory(name, **kw):
    import pytorch_volumetric_b200 as pv
    v, f = workloads.fixture_mesh(name)
    return pv.MeshObjectFactory(name, mesh=(v, f), **kw)


# ---- host mirror of the kernel's deterministic ray jitter (pvb_device.cuh: mix32 / hash_normal) ----
def _mix32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7feb352d)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846ca68b)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def hash_normal(seed, idx, comp):
    idx = np.asarray(idx, dtype=np.uint64)
    lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    with np.errstate(over="ignore"):
        a = _mix32(lo * np.uint32(3) + np.uint32(comp) + np.uint32(0x9e3779b9))
        b = _mix32(hi + np.uint32(0x85ebca6b))
        h0 = _mix32(np.uint32(seed) ^ a ^ b)
        h1 = _mix32(h0 + np.uint32(0x6a09e667))
    s = ((h0 & np.uint32(0xffff)).astype(np.float32) + (h0 >> np.uint32(16)).astype(np.float32)
         + (h1 & np.uint32(0xffff)).astype(np.float32) + (h1 >> np.uint32(16)).astype(np.float32))
    return (s - np.float32(131070.0)) * np.float32(2.6428816e-05)


def ray_noise(seed, n):
    idx = np.arange(n, dtype=np.uint64)
    return np.stack([hash_normal(seed, idx, c) for c in range(3)], axis=1).astype(np.float64)


def classify_mesh_mismatch(d_gpu, g_gpu, d_ref, g_ref, tol=1e-5, coord_scale=0.1):
    """Returns (n_bad_value, n_bad_grad_unexplained, report).

    Gradient tolerance: `tol` plus the fp32 conditioning of (closest - p) / |d| -- the closest point carries a
    few ulps of the coordinate magnitude, which the division by a small |d| amplifies.  A remaining exceedance
    is 'explained' when the point sits on the |d| = 1e-3 shell (sdf.py:162) or when the two sides agree on the
    distance but pick different closest features (ties on the medial axis)."""
    dv = np.abs(d_gpu - d_ref)
    bad_v = dv > tol
    dg = np.abs(g_gpu - g_ref).max(axis=-1)
    gtol = tol + 8 * np.finfo(np.float32).eps * coord_scale / np.maximum(np.abs(d_ref), 1e-12)
    bad_g = dg > gtol
    shell = np.abs(np.abs(d_ref) - 1e-3) < 2e-6
    explained = bad_g & (shell | (dv <= tol))     # same distance, different closest feature / shell flip
    unexplained = bad_g & ~explained
    rep = dict(n=len(d_ref), bad_val=int(bad_v.sum()), bad_grad=int(bad_g.sum()), explained=int(explained.sum()),
               max_dval=float(dv.max()), max_dgrad=float(dg.max()))
    return int(bad_v.sum()), int(unexplained.sum()), rep
