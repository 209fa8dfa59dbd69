"""Signed-distance objects with the reference's Python API, computed by libpvb.so on B200.

Mirrors /root/reference/src/pytorch_volumetric/sdf.py (same class names, constructor
arguments, shapes, dtypes and exceptions) for the batched query path:

  ObjectFactory / MeshObjectFactory   sdf.py:30-214   mesh load, AABB, closest point + sign
  ObjectFrameSDF                      sdf.py:217-282  the operator interface
  SphereSDF                           sdf.py:285-299
  MeshSDF                             sdf.py:302-329
  ComposedSDF                         sdf.py:332-433
  OutOfBoundsStrategy, CachedSDF      sdf.py:436-614
  sample_mesh_points                  sdf.py:617-670

All arithmetic on the query path runs in hand-written sm_100a kernels behind the
C ABI of include/pvb.h; there is no CPU fallback.  Host tensors are accepted
(copied to the GPU, results copied back to the input's device) so that the
classes drop in for CPU callers of the reference.
"""
import abc
import ctypes
import enum
import logging
import math
import os
import typing
from functools import partial
from typing import NamedTuple, Union

import numpy as np
import torch

from . import _native as nat
from . import meshio
from .transforms import Transform3d, matrix_of
from .voxel import (GridView, VoxelGrid, get_coordinates_and_points_in_grid, get_divisible_range_by_resolution,
                    nonempty_indices, range_dtype)

logger = logging.getLogger(__name__)


class SDFQuery(NamedTuple):
    closest: torch.Tensor
    distance: torch.Tensor
    gradient: torch.Tensor
    normal: Union[torch.Tensor, None]


class TriMesh:
    """fp64 vertices + int32 triangles (+ unit face normals): what `ObjectFactory._mesh` holds."""

    def __init__(self, vertices, triangles):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
        self.triangle_normals = None

    def compute_triangle_normals(self):
        v, f = self.vertices, self.triangles
        n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        with np.errstate(invalid="ignore", divide="ignore"):
            n = n / np.linalg.norm(n, axis=1, keepdims=True)
        n[~np.isfinite(n).all(axis=1)] = (0.0, 0.0, 1.0)
        self.triangle_normals = n
        return self

    def triangle_areas(self):
        v, f = self.vertices, self.triangles
        return 0.5 * np.linalg.norm(np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]), axis=1)

    def get_center(self):
        return self.vertices.mean(axis=0)


def _as_trimesh(mesh):
    if isinstance(mesh, TriMesh):
        return mesh
    if isinstance(mesh, (tuple, list)) and len(mesh) == 2:
        return TriMesh(mesh[0], mesh[1])
    if hasattr(mesh, "vertices") and hasattr(mesh, "triangles"):     # e.g. an open3d legacy TriangleMesh
        return TriMesh(np.asarray(mesh.vertices), np.asarray(mesh.triangles))
    raise TypeError("mesh= expects (vertices, faces) or an object with .vertices and .triangles")


def _winding_moments(nodes_u8, tris):
    """Per BVH4 node and child: area-weighted centroid + bounding radius and the sum of area vectors -- the
    first-order data of the hierarchical winding-number evaluation (EXTENSION, see pvb.h PVB_MESH_WINDING).
    nodes_u8 (n,128) uint8 and tris (F,12) float32 are the arrays pvb_bvh_build returned.  -> float32 (n, 8, 4)."""
    n = nodes_u8.shape[0]
    nf = nodes_u8.view(np.float32).reshape(n, 32).astype(np.float64)
    child = nodes_u8.view(np.int32).reshape(n, 32)[:, 24:28].astype(np.int64)
    lo = np.stack([nf[:, 0:4], nf[:, 4:8], nf[:, 8:12]], axis=-1)          # (n, 4, 3)
    hi = np.stack([nf[:, 12:16], nf[:, 16:20], nf[:, 20:24]], axis=-1)
    t = tris.astype(np.float64)
    v0, v1, v2 = t[:, 0:3], t[:, 4:7], t[:, 8:11]
    avec = 0.5 * np.cross(v1 - v0, v2 - v0)                                # area vectors
    area = np.linalg.norm(avec, axis=1)
    acen = area[:, None] * (v0 + v1 + v2) / 3.0
    cs = lambda a: np.concatenate([np.zeros((1,) + a.shape[1:]), np.cumsum(a, axis=0)])   # prefix sums, leaf order
    cs_vec, cs_area, cs_cen = cs(avec), cs(area), cs(acen)
    empty = child == -2 ** 31
    leaf = (child < 0) & ~empty
    inner = child >= 0
    code = np.where(leaf, ~child, 0)
    first, cnt = code >> 2, (code & 3) + 1
    N = np.zeros((n, 4, 3)); A = np.zeros((n, 4)); C = np.zeros((n, 4, 3))
    N[leaf] = (cs_vec[(first + cnt)[leaf]] - cs_vec[first[leaf]])
    A[leaf] = (cs_area[(first + cnt)[leaf]] - cs_area[first[leaf]])
    C[leaf] = (cs_cen[(first + cnt)[leaf]] - cs_cen[first[leaf]])
    # inner children: totals of the child node; children have larger indices (BFS layout), so sweep by depth
    depth = np.zeros(n, dtype=np.int64)
    for i in range(n):                       # parents precede children
        ch = child[i][inner[i]]
        depth[ch] = depth[i] + 1
    for dlev in range(int(depth.max()), -1, -1):
        idx = np.nonzero(depth == dlev)[0]
        if dlev < depth.max():
            sel = inner[idx]
            ii, kk = np.nonzero(sel)
            cidx = child[idx[ii], kk]
            N[idx[ii], kk] = N[cidx].sum(axis=1)
            A[idx[ii], kk] = A[cidx].sum(axis=1)
            C[idx[ii], kk] = C[cidx].sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        cen = C / A[..., None]
    with np.errstate(invalid="ignore"):
        box_c = 0.5 * (lo + hi)
    bad = ~np.isfinite(cen).all(axis=-1) | (A <= 0)
    cen[bad] = np.where(np.isfinite(box_c[bad]), box_c[bad], 0.0)
    with np.errstate(invalid="ignore"):
        rad = np.linalg.norm(np.maximum(np.abs(cen - lo), np.abs(hi - cen)), axis=-1)
    rad[empty] = 0.0
    out = np.zeros((n, 8, 4), dtype=np.float32)
    out[:, 0:4, 0:3] = np.where(empty[..., None], 0.0, cen)
    out[:, 0:4, 3] = np.where(np.isfinite(rad), rad, 0.0)
    out[:, 4:8, 0:3] = np.where(empty[..., None], 0.0, N)
    return out


class _HostChunkStream:
    """Host batch in, pinned host results out, PCIe busy in both directions at once: chunk i is copied in on `s_in`
    while chunk i-1 is computed on the current stream and chunk i-2 is copied out on `s_out`; two device slots per
    direction, events order the reuse of a slot.  `tails` are the trailing shapes of the per-point outputs, e.g.
    ((), (3,)) for value + gradient; `launch(d_in, m, d_outs)` enqueues the computation of one chunk of m points on
    the current stream.

    (Measured on the pool, CachedSDF, 1e7 points: one copy-in / launch / copy-out 5.1 ms, 2M-point chunks 3.7 ms, 4M
    4.1 ms, 1M 11 ms -- small chunks are dominated by per-chunk stream/event bookkeeping on this host.)"""

    def __init__(self, device, chunk, tails):
        self.device, self.chunk, self.tails = device, chunk, tuple(tails)
        with torch.cuda.device(device):
            self.s_in, self.s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
            self.d_in = [torch.empty(chunk, 3, dtype=torch.float32, device=device) for _ in range(2)]
            self.d_out = [[torch.empty(chunk, *t, dtype=torch.float32, device=device) for t in self.tails]
                          for _ in range(2)]

    def run(self, points, launch):
        device, C = self.device, self.chunk
        src = points.detach().reshape(-1, 3)
        if src.dtype != torch.float32:
            src = src.float()
        src = src.contiguous()
        n = src.shape[0]
        with torch.cuda.device(device):
            cur = torch.cuda.current_stream(device)
            outs_h = [torch.empty(n, *t, dtype=torch.float32, pin_memory=True) for t in self.tails]
            self.s_in.wait_stream(cur)
            self.s_out.wait_stream(cur)
            ev_comp = [None, None]      # computation finished reading d_in[b] / writing d_out[b]
            ev_out = [None, None]       # copy-out finished reading d_out[b]
            for i, lo in enumerate(range(0, n, C)):
                hi = min(lo + C, n)
                m = hi - lo
                b = i & 1
                with torch.cuda.stream(self.s_in):
                    if ev_comp[b] is not None:
                        self.s_in.wait_event(ev_comp[b])
                    self.d_in[b][:m].copy_(src[lo:hi], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(self.s_in)
                cur.wait_event(ev_in)
                if ev_out[b] is not None:
                    cur.wait_event(ev_out[b])
                launch(self.d_in[b], m, self.d_out[b])
                ev_comp[b] = torch.cuda.Event()
                ev_comp[b].record(cur)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_comp[b])
                    for h, d in zip(outs_h, self.d_out[b]):
                        h[lo:hi].copy_(d[:m], non_blocking=True)
                    ev_out[b] = torch.cuda.Event()
                    ev_out[b].record(self.s_out)
            self.s_out.synchronize()
            cur.wait_stream(self.s_in)
        return outs_h


class ObjectFactory(abc.ABC):
    def __init__(self, name='', scale=1.0, vis_frame_pos=(0, 0, 0), vis_frame_rot=(0, 0, 0, 1),
                 plausible_suboptimality=0.001, mesh=None, ray_seed=0, **kwargs):
        """
        :param name: path to the mesh obj if loading from file
        :param scale: scaling factor for the mesh
        :param vis_frame_pos: position of the mesh in the object frame
        :param vis_frame_rot: quaternion rotation (xyzw) of the mesh in the object frame
        :param plausible_suboptimality: how much error to tolerate in the SDF
        :param mesh: (vertices, faces) or an object with .vertices/.triangles, given instead of a path;
        scale, vis_frame_pos and vis_frame_rot are then ignored (as in the reference)
        :param ray_seed: seed of the deterministic per-point jitter of the inside/outside ray direction
        (the reference draws it from the unseeded numpy global generator, sdf.py:149)
        """
        self.name = name
        self.scale = scale if scale is not None else 1.0
        self.vis_frame_pos = vis_frame_pos
        self.vis_frame_rot = vis_frame_rot
        self.other_load_kwargs = kwargs
        self.plausible_suboptimality = plausible_suboptimality
        self.ray_seed = int(ray_seed)

        self._mesh = _as_trimesh(mesh) if mesh is not None else None
        self._mesh_given = mesh is not None
        self._face_normals = None
        self._bvh_host = None      # (nodes uint8[n,128], tris float32[F,12], depth)
        self._dev = {}             # device -> dict of device buffers
        self._closed = None
        self._aabb = None
        self.precompute_sdf()

    def __reduce__(self):
        if self._mesh_given:
            return partial(self.__class__, scale=self.scale, vis_frame_pos=self.vis_frame_pos,
                           vis_frame_rot=self.vis_frame_rot, plausible_suboptimality=self.plausible_suboptimality,
                           mesh=(self._mesh.vertices, self._mesh.triangles), **self.other_load_kwargs), (self.name,)
        return partial(self.__class__, scale=self.scale, vis_frame_pos=self.vis_frame_pos,
                       vis_frame_rot=self.vis_frame_rot,
                       plausible_suboptimality=self.plausible_suboptimality, **self.other_load_kwargs), \
            (self.name,)

    @abc.abstractmethod
    def make_collision_obj(self, z, rgba=None):
        """Create collision object of fixed and position along x-y; returns the object ID and bounding box"""

    @abc.abstractmethod
    def get_mesh_resource_filename(self):
        """Return the path to the mesh resource file (.obj, .stl, ...)"""

    def get_mesh_high_poly_resource_filename(self):
        return self.get_mesh_resource_filename()

    def bounding_box(self, padding=0., padding_ratio=0):
        if self._aabb is None:      # the vertices never change after construction
            self._aabb = (self._mesh.vertices.min(axis=0), self._mesh.vertices.max(axis=0))
        ranges = np.stack(self._aabb, axis=1)
        extents = ranges[:, 1] - ranges[:, 0]
        ranges[:, 0] -= padding + padding_ratio * extents
        ranges[:, 1] += padding + padding_ratio * extents
        return ranges

    def center(self):
        if self._mesh is None:
            self.precompute_sdf()
        return self._mesh.get_center()

    @property
    def is_closed(self):
        if self._closed is None:
            self._closed = meshio.is_closed_manifold(self._mesh.triangles)
        return self._closed

    def precompute_sdf(self):
        if self._mesh is None:
            full_path = os.path.expanduser(self.get_mesh_high_poly_resource_filename())
            if not os.path.exists(full_path):
                raise RuntimeError(f"Expected mesh file does not exist: {full_path}")
            v, f = meshio.read_triangle_mesh(full_path)
            v = v * np.asarray(self.scale, dtype=np.float64)                       # sdf.py:105-107
            v = v @ meshio.quaternion_xyzw_to_matrix(self.vis_frame_rot).T         # sdf.py:110-112
            v = v + np.array(self.vis_frame_pos, dtype=np.float64) * self.scale    # sdf.py:113
            self._mesh = TriMesh(v, f)
        if self._bvh_host is None:
            self._mesh.compute_triangle_normals()
            self._face_normals = self._mesh.triangle_normals                       # fp64, sdf.py:119-120
            # the query structure holds fp32 vertices, like the Embree scene built from the
            # float32 tensor mesh at sdf.py:116-118
            self._bvh_host = nat.bvh_build(self._mesh.vertices.astype(np.float32), self._mesh.triangles)

    # -- device residency -------------------------------------------------------
    def _device_state(self, device):
        device = nat.compute_device(device)
        st = self._dev.get(device)
        if st is None:
            nodes, tris, depth = self._bvh_host
            if 3 * depth + 2 > 64:
                raise nat.NativeLibraryError(f"BVH too deep for the traversal stack (depth {depth})")
            st = {
                "nodes": torch.from_numpy(nodes).to(device),
                "tris": torch.from_numpy(tris).to(device),
                "fn32": torch.from_numpy(self._face_normals.astype(np.float32)).contiguous().to(device),
                "v64": torch.from_numpy(self._mesh.vertices).to(device),
                "faces": torch.from_numpy(self._mesh.triangles).to(device),
            }
            self._dev[device] = st
        return st

    def native_desc(self, device):
        """pvb_sdf_desc (kind MESH) for this object on `device`."""
        st = self._device_state(device)
        d = nat.SdfDesc()
        d.kind = nat.PVB_KIND_MESH
        d.flags = 0
        self._fill_mesh_part(d, st)
        return d

    def _winding_state(self, device):
        st = self._device_state(device)
        if "wn" not in st:
            nodes, tris, _ = self._bvh_host
            st["wn"] = torch.from_numpy(_winding_moments(nodes, tris)).to(st["nodes"].device)
        return st["wn"]

    #: set False to force the reference's diagonal ray (sdf.py:147-153) even on closed meshes
    axis_ray_when_closed = True
    #: "parity" (the reference's rule: odd number of ray crossings => inside) or "winding" -- an EXTENSION that
    #: classifies by the generalized winding number |w| > 1/2, robust on open / self-intersecting meshes
    sign_mode = "parity"

    def _fill_mesh_part(self, d, st):
        d.nodes = st["nodes"].data_ptr()
        d.n_nodes = st["nodes"].shape[0]
        d.tris = st["tris"].data_ptr()
        d.n_tris = st["tris"].shape[0]
        d.face_normals = st["fn32"].data_ptr()
        far = self.bounding_box(padding=1.0)[:, 1]                                  # sdf.py:147
        bb = self.bounding_box()
        for k in range(3):
            d.ray_far[k] = float(np.float32(far[k]))
            d.bb_min[k] = float(np.float32(bb[k, 0]))
            d.bb_max[k] = float(np.float32(bb[k, 1]))
        d.ray_seed = self.ray_seed & 0xFFFFFFFF
        if self.axis_ray_when_closed and self.is_closed:
            d.flags |= nat.PVB_MESH_CLOSED
        # slack of the "value >= distance to the AABB" bound used by the composed kernels (fp32 rounding only)
        d.prune_margin = 1e-5 * max(1.0, float(np.abs(bb).max()))

    # -- the query (sdf.py:122-172) ----------------------------------------------
    #: host batches at least this large are streamed through the GPU in chunks (_HostChunkStream); every chunk is
    #: binned and walked on its own
    host_pipeline_min_points = 1 << 22
    host_pipeline_chunk = 1 << 21

    def _do_object_frame_closest_point(self, points_in_object_frame, compute_normal=False, device=None,
                                       mode=nat.PVB_MESH_DEFAULT, want_closest=True):
        """-> (closest | None, distance, gradient, normal | None); want_closest=False (MeshSDF.__call__ needs distance
        and gradient only) skips the closest-point output and its copy to the host."""
        if torch.is_tensor(points_in_object_frame):
            dtype = points_in_object_frame.dtype
            out_device = points_in_object_frame.device
            if device is None and out_device.type == "cuda":
                device = out_device
        else:
            points_in_object_frame = np.asarray(points_in_object_frame)
            dtype = torch.float
            out_device = torch.device("cpu")
        lead = tuple(points_in_object_frame.shape[:-1])
        device = nat.compute_device(device)
        if self.sign_mode not in ("parity", "winding"):
            raise ValueError(f"sign_mode must be 'parity' or 'winding', got {self.sign_mode!r}")
        winding = self.sign_mode == "winding" and bool(mode & nat.PVB_MESH_SIGNED)
        L = nat.lib()
        if (torch.is_tensor(points_in_object_frame) and out_device.type == "cpu" and dtype == torch.float32
                and not winding and points_in_object_frame.shape[-1] == 3
                and points_in_object_frame.numel() // 3 >= self.host_pipeline_min_points
                and (self.axis_ray_when_closed and self.is_closed or not (mode & nat.PVB_MESH_SIGNED))):
            # large host batch: copy-in / tree walk / copy-out overlapped chunk by chunk.  (Only where the sign test is
            # the exact axis-aligned walk: the jittered diagonal ray of open meshes is seeded by the point's index in
            # the call, so chunking would change which ray a point gets.)
            tails = ((3,),) * bool(want_closest) + ((), (3,)) + ((3,),) * bool(compute_normal)
            key = (device, self.host_pipeline_chunk, tails)
            st = getattr(self, "_pipe_state", None)
            if st is None or st[0] != key:
                with torch.cuda.device(device):
                    st = self._pipe_state = (key, _HostChunkStream(device, self.host_pipeline_chunk, tails),
                                             nat.query_workspace(self.host_pipeline_chunk, device))
            desc = self.native_desc(device)
            ws = st[2]

            def launch(d_in, m, d_out):
                o = list(d_out)
                closest = o.pop(0) if want_closest else None
                dist, grad = o.pop(0), o.pop(0)
                normal = o.pop(0) if compute_normal else None
                nat.check(L.pvb_mesh_query(ctypes.byref(desc), nat.ptr(d_in), m, mode, nat.ptr(dist), nat.ptr(grad),
                                           nat.ptr(closest), None, nat.ptr(normal), nat.ptr(ws),
                                           ws.numel() if ws is not None else 0, nat.stream_ptr(device)),
                          "pvb_mesh_query")

            o = list(st[1].run(points_in_object_frame, launch))
            closest = o.pop(0).reshape(*lead, 3) if want_closest else None
            dist, grad = o.pop(0).reshape(lead), o.pop(0).reshape(*lead, 3)
            normal = o.pop(0).reshape(*lead, 3) if compute_normal else None
            return closest, dist, grad, normal
        with torch.cuda.device(device):
            p = nat.as_f32_points(points_in_object_frame, device)
            n = p.shape[0]
            dist = torch.empty(n, dtype=torch.float32, device=device)
            grad = torch.empty(n, 3, dtype=torch.float32, device=device)
            closest = torch.empty(n, 3, dtype=torch.float32, device=device) if want_closest else None
            normal = torch.empty(n, 3, dtype=torch.float32, device=device) if compute_normal else None
            desc = self.native_desc(device)
            face = None
            if winding:
                mode |= nat.PVB_MESH_WINDING
                desc.wn_nodes = self._winding_state(device).data_ptr()
                face = torch.empty(n, dtype=torch.int32, device=device)
            ws = nat.query_workspace(n, device)
            nat.check(L.pvb_mesh_query(ctypes.byref(desc), nat.ptr(p), n, mode, nat.ptr(dist), nat.ptr(grad),
                                       nat.ptr(closest), nat.ptr(face), nat.ptr(normal), nat.ptr(ws),
                                       ws.numel() if ws is not None else 0, nat.stream_ptr(device)),
                      "pvb_mesh_query")

        def fin(t, tail):
            return nat.deliver(t, out_device, dtype).reshape(*lead, *tail)

        return (fin(closest, (3,)) if want_closest else None, fin(dist, ()), fin(grad, (3,)),
                fin(normal, (3,)) if compute_normal else None)

    def object_frame_closest_point(self, points_in_object_frame, compute_normal=False) -> SDFQuery:
        """
        :param points_in_object_frame: N x 3 points in the object frame (arbitrary batch dimensions in front of N)
        :param compute_normal: whether to also return the surface normal at the closest point
        :return: SDFQuery(closest N x 3, signed distance N, gradient N x 3, normal N x 3 | None)
        """
        return SDFQuery(*self._do_object_frame_closest_point(points_in_object_frame, compute_normal=compute_normal))


class MeshObjectFactory(ObjectFactory):
    def __init__(self, mesh_name='', path_prefix='', **kwargs):
        self.path_prefix = path_prefix
        # strip the package:// prefix when a path prefix is given (loading the mesh manually)
        self.strip_package_prefix = path_prefix != ''
        super(MeshObjectFactory, self).__init__(mesh_name, **kwargs)

    def __reduce__(self):
        fn, args = super().__reduce__()
        return partial(fn.func, path_prefix=self.path_prefix, **fn.keywords), args

    def make_collision_obj(self, z, rgba=None):
        return None, None

    def get_mesh_resource_filename(self):
        mesh_path = self.name
        if self.strip_package_prefix:
            mesh_path = mesh_path.replace("package://", "")
        return os.path.join(self.path_prefix, mesh_path)


class ObjectFrameSDF(abc.ABC):
    @abc.abstractmethod
    def __call__(self, points_in_object_frame):
        """
        Evaluate the signed distance function at given points in the object frame
        :param points_in_object_frame: B x N x d points; located in object frame
        :return: tuple of B x N signed distance (m) and B x N x d SDF gradient pointing towards higher SDF values
        """

    @abc.abstractmethod
    def surface_bounding_box(self, padding=0., padding_ratio=0.):
        """(min,max) per dimension of the 0-level set, inflated by padding + padding_ratio * extent"""

    def outside_surface(self, points_in_object_frame, surface_level=0):
        sdf_values, _ = self.__call__(points_in_object_frame)
        return sdf_values > surface_level

    def get_voxel_view(self, voxels: VoxelGrid = None, dtype=torch.float, device='cpu') -> GridView:
        if voxels is None:
            voxels = VoxelGrid(0.01, self.surface_bounding_box(padding=0.1).cpu().numpy(), dtype=dtype, device=device)
        pts = voxels.get_voxel_center_points()
        sdf_val, sdf_grad = self.__call__(pts.unsqueeze(0))
        cached_underlying_sdf = sdf_val.reshape([len(coord) for coord in voxels.coords])
        return GridView(cached_underlying_sdf, voxels.range_per_dim,
                        invalid_value=lambda p: self.__call__(p)[0])

    def get_filtered_points(self, unary_filter, voxels: VoxelGrid = None, dtype=torch.float,
                            device='cpu') -> torch.tensor:
        model_voxels = self.get_voxel_view(voxels, dtype=dtype, device=device)
        interior = unary_filter(model_voxels.raw_data)
        flat = nonempty_indices(interior.reshape(-1))
        idx = torch.stack(torch.unravel_index(flat, tuple(model_voxels.shape)), dim=-1)
        return model_voxels.ensure_value_key(idx)

    # hook for the fused composition kernels: a pvb_sdf_desc on `device`, or None if this SDF has to be
    # evaluated through its Python __call__
    def native_desc(self, device):
        return None


def _result_like(points, lead, tensors_tails, out_device=None):
    dtype = points.dtype if torch.is_tensor(points) and points.dtype.is_floating_point else torch.float
    dev = out_device if out_device is not None else (points.device if torch.is_tensor(points) else torch.device("cpu"))
    return tuple(nat.deliver(t, dev, dtype).reshape(*lead, *tail) for t, tail in tensors_tails)


class SphereSDF(ObjectFrameSDF):
    """SDF for a geometric primitive, the sphere centered at the origin"""

    def __init__(self, radius):
        self.radius = radius

    def __call__(self, points_in_object_frame):
        lead = tuple(points_in_object_frame.shape[:-1])
        device = nat.compute_device(points_in_object_frame.device if torch.is_tensor(points_in_object_frame) else None)
        with torch.cuda.device(device):
            p = nat.as_f32_points(points_in_object_frame, device)
            n = p.shape[0]
            val = torch.empty(n, dtype=torch.float32, device=device)
            grad = torch.empty(n, 3, dtype=torch.float32, device=device)
            nat.check(nat.lib().pvb_sphere_query(float(self.radius), nat.ptr(p), n, nat.ptr(val), nat.ptr(grad),
                                                 nat.stream_ptr(device)), "pvb_sphere_query")
        return _result_like(points_in_object_frame, lead, ((val, ()), (grad, (3,))))

    def surface_bounding_box(self, padding=0., padding_ratio=0.):
        length = self.radius + padding + padding_ratio * self.radius
        return torch.tensor([[-length, length], [-length, length], [-length, length]])

    def native_desc(self, device):
        d = nat.SdfDesc()
        d.kind = nat.PVB_KIND_SPHERE
        d.radius = float(self.radius)
        for k in range(3):
            d.bb_min[k] = -float(self.radius)
            d.bb_max[k] = float(self.radius)
        return d


class MeshSDF(ObjectFrameSDF):
    """SDF from direct BVH closest-point + ray-parity queries against the mesh."""

    def __init__(self, obj_factory: ObjectFactory, vis=None):
        if vis is not None:
            raise NotImplementedError("debug drawing (vis=) is outside the query path; pass vis=None")
        self.obj_factory = obj_factory
        self.vis = vis

    def surface_bounding_box(self, **kwargs):
        return torch.tensor(self.obj_factory.bounding_box(**kwargs))

    def __call__(self, points_in_object_frame):
        of = self.obj_factory
        if (type(of).object_frame_closest_point is ObjectFactory.object_frame_closest_point
                and type(of)._do_object_frame_closest_point is ObjectFactory._do_object_frame_closest_point):
            # stock factory: distance and gradient only -- no closest-point output, no copy of it to a host caller
            _, dist, grad, _ = of._do_object_frame_closest_point(points_in_object_frame, want_closest=False)
            return dist, grad
        res = of.object_frame_closest_point(points_in_object_frame)
        return res.distance, res.gradient

    def native_desc(self, device):
        if self.obj_factory.sign_mode != "parity":
            return None      # the fused composition kernels implement the reference's parity rule only
        return self.obj_factory.native_desc(device)


def _affine_inverse(m):
    """Inverse of (n,4,4) affine matrices [A|t; 0 0 0 1] in closed form (adjugate of the 3x3)."""
    A = m[:, :3, :3]
    t = m[:, :3, 3:]
    c0 = torch.linalg.cross(A[:, :, 1], A[:, :, 2], dim=-1)
    c1 = torch.linalg.cross(A[:, :, 2], A[:, :, 0], dim=-1)
    c2 = torch.linalg.cross(A[:, :, 0], A[:, :, 1], dim=-1)
    det = (A[:, :, 0] * c0).sum(-1)
    Ainv = torch.stack([c0, c1, c2], dim=1) / det[:, None, None]
    out = torch.zeros_like(m)
    out[:, :3, :3] = Ainv
    out[:, :3, 3:] = -(Ainv @ t)
    out[:, 3, 3] = 1
    return out


class ComposedSDF(ObjectFrameSDF):
    def __init__(self, sdfs: typing.Sequence[ObjectFrameSDF], obj_frame_to_each_frame):
        """
        :param sdfs: S Object frame SDFs
        :param obj_frame_to_each_frame: [B*]S x 4 x 4 transforms (a Transform3d-like with get_matrix()) from the
        shared object frame to the frame of each SDF, flattened SDF-major when batched over configurations
        """
        self.sdfs = sdfs
        self.obj_frame_to_link_frame = None
        self._link_to_obj = None
        self.tsf_batch = None
        self._desc_cache = {}
        self.set_transforms(obj_frame_to_each_frame)

    # the reference exposes the inverted transforms as a list; build it on demand
    @property
    def link_frame_to_obj_frame(self):
        if self._link_to_obj is None and self.obj_frame_to_link_frame is not None:
            inv = _affine_inverse(matrix_of(self.obj_frame_to_link_frame))
            self._link_to_obj = [Transform3d(matrix=inv[self.ith_transform_slice(i)]) for i in range(len(self.sdfs))]
        return self._link_to_obj if self._link_to_obj is not None else []

    def surface_bounding_box(self, **kwargs):
        # transforms only the (min, max) corner pair of every sub-box, as the reference does (sdf.py:347-368)
        m_inv = _affine_inverse(matrix_of(self.obj_frame_to_link_frame))
        bounds = []
        for i, sdf in enumerate(self.sdfs):
            pts = sdf.surface_bounding_box(**kwargs)
            tsf = Transform3d(matrix=m_inv[self.ith_transform_slice(i)])
            pts = tsf.transform_points(pts.to(dtype=m_inv.dtype, device=m_inv.device).transpose(0, 1))
            if self.tsf_batch is not None and len(pts.shape) == 2:
                pts = pts.unsqueeze(0)
            bounds.append(pts)
        bounds = torch.stack(bounds)
        if self.tsf_batch is not None:
            dims = (0,) + tuple(range(2, len(bounds.shape) - 1))
        else:
            dims = tuple(range(len(bounds.shape) - 1))
        mins = bounds.amin(dim=dims)
        maxs = bounds.amax(dim=dims)
        return torch.stack((mins, maxs), dim=-1)

    def set_transforms(self, tsf, batch_dim=None):
        self.obj_frame_to_link_frame = tsf
        self._link_to_obj = None
        self._xf_dev = {}
        self.tsf_batch = batch_dim
        if tsf is not None:
            S = len(self.sdfs)
            S_tsf = matrix_of(tsf).shape[0]
            if self.tsf_batch is None and (S_tsf != S):
                self.tsf_batch = (S_tsf // S,)
            if S_tsf % S != 0:
                raise ValueError(f"{S_tsf} transforms cannot be split over {S} SDFs")

    def ith_transform_slice(self, i):
        if self.tsf_batch is None:
            return slice(i, i + 1)
        total_to_slice = math.prod(list(self.tsf_batch))
        return slice(i * total_to_slice, (i + 1) * total_to_slice)

    # -- fused path ---------------------------------------------------------------
    def _native_descs(self, device):
        key = (device, tuple(id(s) for s in self.sdfs))
        hit = self._desc_cache.get(key)
        if hit is None:
            descs = [s.native_desc(device) for s in self.sdfs]
            if any(d is None for d in descs):
                hit = (None, False, None)
            else:
                needs_mesh = any(d.kind == nat.PVB_KIND_MESH or (d.flags & nat.PVB_GRID_OOB_GT) for d in descs)
                if len(descs) > 128:
                    raise ValueError("ComposedSDF: at most 128 sub-SDFs are supported by the fused kernel")
                hit = (nat.desc_array(descs), needs_mesh, descs)
            self._desc_cache = {key: hit}
        return hit[0], hit[1]

    def _xforms_on(self, device):
        xf = self._xf_dev.get(device)
        if xf is None:
            xf = matrix_of(self.obj_frame_to_link_frame).detach().to(device=device, dtype=torch.float32).contiguous()
            self._xf_dev = {device: xf}
        return xf

    def query(self, points_in_object_frame, cfg_begin=0, cfg_count=None, return_which=False):
        """__call__ restricted to the configuration slab [cfg_begin, cfg_begin + cfg_count) (multi-GPU sharding
        over configurations).  Returns flat (cfg_count * P,) / (cfg_count * P, 3) fp32 tensors on the GPU."""
        S = len(self.sdfs)
        n_cfg = 1 if self.tsf_batch is None else math.prod(list(self.tsf_batch))
        if cfg_count is None:
            cfg_count = n_cfg - cfg_begin
        device = nat.compute_device(points_in_object_frame.device if torch.is_tensor(points_in_object_frame) else None)
        with torch.cuda.device(device):
            p = nat.as_f32_points(points_in_object_frame, device)
            P = p.shape[0]
            descs_arr, needs_mesh = self._native_descs(device)
            if descs_arr is None:
                return self._generic_query(p, cfg_begin, cfg_count, n_cfg, return_which)
            xf = self._xforms_on(device)
            val = torch.empty(cfg_count * P, dtype=torch.float32, device=device)
            grad = torch.empty(cfg_count * P, 3, dtype=torch.float32, device=device)
            which = torch.empty(cfg_count * P, dtype=torch.int32, device=device) if return_which else None
            nat.check(nat.lib().pvb_composed_query(descs_arr, S, int(needs_mesh), nat.ptr(xf), n_cfg,
                                                   cfg_begin, cfg_count, nat.ptr(p), P, nat.PVB_MESH_DEFAULT,
                                                   nat.ptr(val), nat.ptr(grad), nat.ptr(which),
                                                   nat.stream_ptr(device)), "pvb_composed_query")
        return (val, grad, which) if return_which else (val, grad)

    def query_into(self, points_in_object_frame, targets, cfg_begin=0, cfg_count=None):
        """`query` whose result slab is stored by the kernel epilogue into several full-size result buffers at once
        (pvb_composed_query_multi): `targets` is a list of (val, grad) pairs, each a float32 tensor or a raw device
        address of a buffer holding ALL n_cfg * P results -- typically this rank's buffer and the peer-mapped buffers
        of the other ranks (distributed.PeerResult).  The slab lands at element cfg_begin * P of every buffer."""
        S = len(self.sdfs)
        n_cfg = 1 if self.tsf_batch is None else math.prod(list(self.tsf_batch))
        if cfg_count is None:
            cfg_count = n_cfg - cfg_begin
        if not 1 <= len(targets) <= nat.MAX_TARGETS:
            raise ValueError(f"between 1 and {nat.MAX_TARGETS} targets per launch, got {len(targets)}")
        device = nat.compute_device(points_in_object_frame.device if torch.is_tensor(points_in_object_frame) else None)
        with torch.cuda.device(device):
            p = nat.as_f32_points(points_in_object_frame, device)
            P = p.shape[0]
            descs_arr, needs_mesh = self._native_descs(device)
            if descs_arr is None:
                raise NotImplementedError("query_into needs sub-SDFs with native descriptors (mesh / cached / sphere)")
            arr = (nat.OutTarget * len(targets))()
            for t, (tv, tg) in enumerate(targets):
                for buf, per in ((tv, 1), (tg, 3)):
                    if torch.is_tensor(buf) and (buf.dtype != torch.float32 or not buf.is_contiguous()
                                                 or buf.numel() < per * n_cfg * P):
                        raise ValueError("targets must be contiguous float32 buffers of the full (n_cfg * P) result")
                arr[t].val = (tv.data_ptr() if torch.is_tensor(tv) else int(tv)) + 4 * cfg_begin * P
                arr[t].grad = (tg.data_ptr() if torch.is_tensor(tg) else int(tg)) + 12 * cfg_begin * P
            xf = self._xforms_on(device)
            nat.check(nat.lib().pvb_composed_query_multi(descs_arr, S, int(needs_mesh), nat.ptr(xf), n_cfg, cfg_begin,
                                                         cfg_count, nat.ptr(p), P, nat.PVB_MESH_DEFAULT,
                                                         ctypes.cast(arr, ctypes.c_void_p), len(targets), None,
                                                         nat.stream_ptr(device)), "pvb_composed_query_multi")

    def query_at(self, points_in_object_frame, val_addr, grad_addr, cfg_begin=0, cfg_count=None):
        """`query` that writes the slab into caller-owned device memory: val_addr / grad_addr are the addresses of
        FULL (n_cfg * P) value / gradient buffers (float32); the slab lands at element cfg_begin * P.  Single
        destination, the kernels' fastest store path -- what the copy-engine re-assembly (gather="dma") runs before it
        pushes the slab to the peers."""
        S = len(self.sdfs)
        n_cfg = 1 if self.tsf_batch is None else math.prod(list(self.tsf_batch))
        if cfg_count is None:
            cfg_count = n_cfg - cfg_begin
        device = nat.compute_device(points_in_object_frame.device if torch.is_tensor(points_in_object_frame) else None)
        with torch.cuda.device(device):
            p = nat.as_f32_points(points_in_object_frame, device)
            P = p.shape[0]
            descs_arr, needs_mesh = self._native_descs(device)
            if descs_arr is None:
                raise NotImplementedError("query_at needs sub-SDFs with native descriptors (mesh / cached / sphere)")
            xf = self._xforms_on(device)
            nat.check(nat.lib().pvb_composed_query(descs_arr, S, int(needs_mesh), nat.ptr(xf), n_cfg, cfg_begin,
                                                   cfg_count, nat.ptr(p), P, nat.PVB_MESH_DEFAULT,
                                                   int(val_addr) + 4 * cfg_begin * P, int(grad_addr) + 12 * cfg_begin * P,
                                                   None, nat.stream_ptr(device)), "pvb_composed_query")

    def query_multicast(self, points_in_object_frame, mc_val, mc_grad, cfg_begin=0, cfg_count=None):
        """`query` whose result slab leaves through an NVLS multicast mapping (pvb_composed_query_multicast): mc_val /
        mc_grad are the multicast device addresses of the full (n_cfg * P) value / gradient buffers; one multimem.st
        per 16-byte chunk lands in every bound GPU's copy."""
        S = len(self.sdfs)
        n_cfg = 1 if self.tsf_batch is None else math.prod(list(self.tsf_batch))
        if cfg_count is None:
            cfg_count = n_cfg - cfg_begin
        device = nat.compute_device(points_in_object_frame.device if torch.is_tensor(points_in_object_frame) else None)
        with torch.cuda.device(device):
            p = nat.as_f32_points(points_in_object_frame, device)
            P = p.shape[0]
            descs_arr, needs_mesh = self._native_descs(device)
            if descs_arr is None:
                raise NotImplementedError("query_multicast needs sub-SDFs with native descriptors")
            xf = self._xforms_on(device)
            nat.check(nat.lib().pvb_composed_query_multicast(descs_arr, S, int(needs_mesh), nat.ptr(xf), n_cfg, cfg_begin,
                                                             cfg_count, nat.ptr(p), P, nat.PVB_MESH_DEFAULT,
                                                             int(mc_val) + 4 * cfg_begin * P,
                                                             int(mc_grad) + 12 * cfg_begin * P,
                                                             nat.stream_ptr(device)), "pvb_composed_query_multicast")

    def _generic_query(self, p, cfg_begin, cfg_count, n_cfg, return_which):
        """Sub-SDFs without a native descriptor (user subclasses, nested compositions): per-SDF evaluation through
        their own __call__, with the transform and the running min still on the GPU."""
        S = len(self.sdfs)
        device = p.device
        P = p.shape[0]
        M = self._xforms_on(device).reshape(S, n_cfg, 4, 4)[:, cfg_begin:cfg_begin + cfg_count]
        best = torch.full((cfg_count * P,), float("inf"), dtype=torch.float32, device=device)
        bgrad = torch.zeros(cfg_count * P, 3, dtype=torch.float32, device=device)
        bwhich = torch.full((cfg_count * P,), -1, dtype=torch.int32, device=device)
        for i, sdf in enumerate(self.sdfs):
            xf = M[i].contiguous()
            local = torch.empty(cfg_count, P, 3, dtype=torch.float32, device=device)
            nat.check(nat.lib().pvb_transform_points(nat.ptr(xf), cfg_count, nat.ptr(p), P, nat.ptr(local),
                                                     nat.stream_ptr(device)), "pvb_transform_points")
            v, g = sdf(local)
            v = v.to(device=device, dtype=torch.float32).reshape(cfg_count, P)
            g = g.to(device=device, dtype=torch.float32).reshape(cfg_count, P, 3) @ xf[:, :3, :3]
            better = (v.reshape(-1) < best) | (bwhich < 0)
            best = torch.where(better, v.reshape(-1), best)
            bgrad = torch.where(better.unsqueeze(-1), g.reshape(-1, 3), bgrad)
            bwhich = torch.where(better, torch.full_like(bwhich, i), bwhich)
        return (best, bgrad, bwhich) if return_which else (best, bgrad)

    #: host callers with at least this many result bytes get their result streamed out slab by slab
    host_result_pipeline_min_bytes = 64 << 20

    def _host_result_pipeline(self, points_in_object_frame):
        """Host points in, pinned host results out, for results so large that the device-to-host copy IS the call
        (C4: 0.45 ms of lookups, 5.8 ms of PCIe): the configurations are queried in slabs (a first small one, then 64 at
        a time -- whole 32-configuration tiles) into one device buffer, and each slab starts its copy on a side stream
        as soon as its kernel has finished, so only the first slab's lookups are not hidden behind the copy.
        Returns None when the call does not qualify (then __call__ takes the plain path)."""
        if not (torch.is_tensor(points_in_object_frame) and points_in_object_frame.device.type == "cpu"
                and points_in_object_frame.dtype == torch.float32 and self.tsf_batch is not None):
            return None
        n_cfg = math.prod(list(self.tsf_batch))
        P = points_in_object_frame.numel() // 3
        if n_cfg < 64 or 16 * n_cfg * P < self.host_result_pipeline_min_bytes:
            return None
        device = nat.compute_device(None)
        if self._native_descs(device)[0] is None:
            return None
        with torch.cuda.device(device):
            cur = torch.cuda.current_stream(device)
            p = nat.as_f32_points(points_in_object_frame, device)
            val_d = torch.empty(n_cfg * P, dtype=torch.float32, device=device)
            grad_d = torch.empty(n_cfg * P, 3, dtype=torch.float32, device=device)
            val_h = torch.empty(n_cfg * P, dtype=torch.float32, pin_memory=True)
            grad_h = torch.empty(n_cfg * P, 3, dtype=torch.float32, pin_memory=True)
            s_out = getattr(self, "_s_out", None)
            if s_out is None or s_out.device != device:
                s_out = self._s_out = torch.cuda.Stream(device)
            b = 0
            while b < n_cfg:
                c = min(32 if b == 0 else 64, n_cfg - b)
                self.query_at(p, val_d.data_ptr(), grad_d.data_ptr(), b, c)
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    val_h[b * P:(b + c) * P].copy_(val_d[b * P:(b + c) * P], non_blocking=True)
                    grad_h[b * P:(b + c) * P].copy_(grad_d[b * P:(b + c) * P], non_blocking=True)
                b += c
            s_out.synchronize()        # the device buffers may be recycled from here on
        return val_h, grad_h

    def __call__(self, points_in_object_frame):
        pts_shape = tuple(points_in_object_frame.shape)
        piped = self._host_result_pipeline(points_in_object_frame)
        if piped is not None:
            vv, gg = piped
        else:
            vv, gg = self.query(points_in_object_frame)
            dtype = points_in_object_frame.dtype if torch.is_tensor(points_in_object_frame) else torch.float
            out_device = points_in_object_frame.device if torch.is_tensor(points_in_object_frame) else "cpu"
            vv = nat.deliver(vv, out_device, dtype)
            gg = nat.deliver(gg, out_device, dtype)
        if self.tsf_batch is not None:
            # configuration batch dims first, then the query points' batch dims (sdf.py:428-431)
            vv = vv.reshape(*self.tsf_batch, *pts_shape[:-1])
            gg = gg.reshape(*self.tsf_batch, *pts_shape[:-1], 3)
        return vv, gg


class OutOfBoundsStrategy(enum.Enum):
    LOOKUP_GT_SDF = 0
    BOUNDING_BOX = 1  # always under-approximates the SDF value, but more accurate than a sphere approximation


def _fp32_ceil(x):
    f = np.float32(x)
    return f if float(f) >= x else np.nextafter(f, np.float32(np.inf))


def _fp32_floor(x):
    f = np.float32(x)
    return f if float(f) <= x else np.nextafter(f, np.float32(-np.inf))


def fast_index_band(lo, hi, n):
    """(inv_res32, idx_certain) of one grid axis for the kernels' fast voxel index.

    The kernels estimate the index as q = (p - fp32(lo)) * inv_res32 in fp32 and accept rint(q) when
    |q - rint(q)| <= idx_certain; closer to a cell boundary than that, they evaluate the reference's formula
    round((p - lo) / res) exactly (fp64 or fp32, by the dtype torch infers for the range).  `err` bounds
    |q - (p - lo) / res| in cells: rounding lo to fp32, the fp32 subtraction, the rounding of inv_res32 and the fp32
    product, with a 4x safety factor; the identity "certain => same index" is checked on adversarial inputs by
    tests/test_index_band.py.  idx_certain = -1 sends every point down the exact path."""
    res = (hi - lo) / (n - 1) if n > 1 else float("inf")
    if not (n > 1 and math.isfinite(res) and res > 0):
        return 0.0, -1.0
    scale = max(abs(lo), abs(hi)) + (hi - lo)
    err = 4.0 * scale * 2.0 ** -24 / res + (n + 4) * 2.0 ** -22
    return float(np.float32(1.0 / res)), max(0.5 - (4.0 * err + 1e-6), -1.0)


def grid_prune_margin(values, lo, hi, bb):
    """Margin m such that a BOUNDING_BOX CachedSDF never returns less than dist(q, bb) - m, for ANY query q.

    The composed kernels skip a sub-SDF when dist(q, bb) - m already exceeds the running minimum, so the bound must
    be exact, not typical.  Out of range the value IS dist(q, bb) (sdf.py:555-571).  In range it is the table entry
    of the nearest voxel c, |q - c| <= half a cell diagonal, dist(., bb) is 1-Lipschitz, and t = max over the table
    of (dist(c, bb) - value(c))+ is measured on the table itself, so the bound holds for any table contents:
    value(c) >= dist(c, bb) - t >= dist(q, bb) - half_diag - t.  (1e-5 absorbs the kernels' fp32 arithmetic.)
    values: (nx, ny, nz) table; lo / hi: per-axis range; bb: (3, 2) box the out-of-range rule measures against."""
    shape = tuple(values.shape)
    dev = values.device
    bbt = torch.as_tensor(np.asarray(bb), dtype=torch.float64, device=dev)
    per_axis = []
    for k in range(3):
        c = torch.linspace(lo[k], hi[k], shape[k], dtype=torch.float64, device=dev)
        per_axis.append(torch.clamp(torch.maximum(bbt[k, 0] - c, c - bbt[k, 1]), min=0))
    lb = torch.sqrt(per_axis[0][:, None, None] ** 2 + per_axis[1][None, :, None] ** 2 + per_axis[2][None, None, :] ** 2)
    t = float(torch.clamp(lb - values.double(), min=0).max())
    cell = [(hi[k] - lo[k]) / (shape[k] - 1) for k in range(3) if shape[k] > 1]
    return 0.5 * math.sqrt(sum(r * r for r in cell)) + t + 1e-5


class _TableStore:
    """The on-disk table cache of CachedSDF: one torch.save file holding {name: (val[nx,ny,nz], grad[Nvox,3])} for any
    number of objects -- the layout of the reference's sdf_cache.pkl (sdf.py:487-519), so existing files keep
    loading and files written here load in the reference."""

    def __init__(self, path):
        self.path = path
        self.entries = (torch.load(path) or {}) if os.path.exists(path) else {}

    def get(self, name):
        entry = self.entries.get(name)
        if entry is None:
            return None
        try:
            val, grad = entry
        except (ValueError, TypeError):
            logger.info("cached sdf invalid %s from %s, recreating", name, self.path)
            return None
        logger.info("cached sdf for %s loaded from %s", name, self.path)
        return val, grad

    def put(self, name, tables):
        self.entries[name] = tuple(t.cpu() for t in tables)
        torch.save(self.entries, self.path)
        logger.info("caching sdf for %s to %s", name, self.path)


class CachedSDF(ObjectFrameSDF):
    """SDF via nearest-voxel lookup of precomputed value and gradient tables."""

    def __init__(self, object_name, resolution, range_per_dim, gt_sdf: ObjectFrameSDF,
                 out_of_bounds_strategy=OutOfBoundsStrategy.BOUNDING_BOX,
                 device="cpu", clean_cache=False,
                 debug_check_sdf=False, cache_path="sdf_cache.pkl", interpolation="nearest"):
        """
        :param object_name: readable name of the object; combined with the resolution and range for the cache key
        :param resolution: side length of each voxel cell
        :param range_per_dim: (min, max) sequence for each dimension
        :param gt_sdf: ground truth SDF used to generate the cache and (optionally) for out-of-range queries
        :param out_of_bounds_strategy: LOOKUP_GT_SDF or BOUNDING_BOX (default)
        :param device: device results are returned on (the tables always live on the GPU)
        :param clean_cache: ignore an existing cache entry and recompute
        :param debug_check_sdf: check the generated tables against the ground truth SDF
        :param cache_path: torch.save file holding {name: (val[nx,ny,nz], grad[Nvox,3])}, same format as the reference
        :param interpolation: "nearest" (the reference's behaviour: nearest voxel value and stored gradient) or
        "trilinear" -- an EXTENSION that interpolates the value table trilinearly and returns the gradient of the
        interpolant (cell-wise finite differences); smoother, but O(resolution) away from the reference's output
        """
        if interpolation not in ("nearest", "trilinear"):
            raise ValueError(f"interpolation must be 'nearest' or 'trilinear', got {interpolation!r}")
        self.interpolation = interpolation
        self.device = device
        self.voxels = None
        self.voxels_grad = None
        self.out_of_bounds_strategy = out_of_bounds_strategy
        self.gt_sdf = gt_sdf
        self.resolution = resolution
        self._cdev = nat.compute_device(device)

        span = np.asarray(range_per_dim, dtype=np.float64)
        cells = (span[:, 1] - span[:, 0]) // resolution
        if cells.min() < 10:
            logger.warning(f"Resolution {resolution} is too high for {object_name}, only getting {cells} voxels.")
        self.ranges = get_divisible_range_by_resolution(resolution, range_per_dim)
        range_per_dim = self.ranges
        # cache key: name, resolution and the snapped range, the format existing sdf_cache.pkl files were written with
        self.name = f"{object_name} {resolution} {tuple(self.ranges)}"
        self.debug_check_sdf = debug_check_sdf

        store = _TableStore(cache_path)
        tables = None if clean_cache else store.get(self.name)
        if tables is None:
            if gt_sdf is None:
                raise RuntimeError("Cached SDF did not find the cache and requires an initialize queryable SDF")
            tables = self._build_tables(gt_sdf)
            store.put(self.name, tables)
        cached_underlying_sdf, cached_underlying_sdf_grad = tables

        val = cached_underlying_sdf.to(device=self._cdev, dtype=torch.float32)
        grad = cached_underlying_sdf_grad.to(device=self._cdev, dtype=torch.float32).reshape(-1, 3)
        self.voxels = GridView(val, range_per_dim, invalid_value=self._fallback_sdf_value_func)
        self.voxels_grad = grad
        # one 16-byte record {val, gx, gy, gz} per voxel: each lookup is a single 128-bit load
        self._table = torch.cat([val.reshape(-1, 1), grad], dim=1).contiguous()
        self.bb = self.surface_bounding_box().to(device=self._cdev)
        self._desc = self._make_desc()

    def _build_tables(self, gt_sdf):
        """(val[nx,ny,nz], grad[Nvox,3]): ONE batched ground-truth query over every voxel centre, on the GPU
        (the reference evaluates the same points through its CPU MeshSDF, sdf.py:502-505)."""
        coords, pts = get_coordinates_and_points_in_grid(self.resolution, self.ranges)
        sdf_val, sdf_grad = gt_sdf(pts.to(self._cdev))
        val = sdf_val.reshape([len(coord) for coord in coords])
        grad = sdf_grad.squeeze(0)
        if self.debug_check_sdf:
            debug_view = GridView(val, self.ranges, invalid_value=self._fallback_sdf_value_func)
            query = debug_view[pts.to(val.device)]
            assert torch.allclose(sdf_val.reshape(-1), query.reshape(-1))
        return val, grad

    # -- descriptor -----------------------------------------------------------------
    def _make_desc(self):
        d = nat.SdfDesc()
        d.kind = nat.PVB_KIND_GRID
        flags = 0
        d.table = self._table.data_ptr()
        shape = tuple(self.voxels.shape)
        lo = [float(min(r)) for r in self.ranges]
        hi = [float(max(r)) for r in self.ranges]
        fp32_index = range_dtype(self.ranges) == torch.float32
        if fp32_index:
            flags |= nat.PVB_GRID_INDEX_FP32
        bb = self.bb.detach().cpu().numpy().astype(np.float32)       # fp32 like the cast at sdf.py:556-557
        res_max = 0.0
        for k in range(3):
            d.dims[k] = shape[k]
            d.min64[k] = lo[k]
            d.res64[k] = (hi[k] - lo[k]) / (shape[k] - 1) if shape[k] > 1 else float("inf")
            lo32, hi32 = np.float32(lo[k]), np.float32(hi[k])
            d.min32[k] = float(lo32)
            d.res32[k] = float((hi32 - lo32) / np.float32(shape[k] - 1)) if shape[k] > 1 else float("inf")
            if fp32_index:
                d.valid_lo[k], d.valid_hi[k] = float(lo32), float(hi32)
            else:   # fp32 bounds equivalent to the fp64 comparison min <= double(p) <= max
                d.valid_lo[k], d.valid_hi[k] = float(_fp32_ceil(lo[k])), float(_fp32_floor(hi[k]))
            d.bb_min[k] = float(bb[k, 0])
            d.bb_max[k] = float(bb[k, 1])
            d.inv_res32[k], d.idx_certain[k] = fast_index_band(lo[k], hi[k], shape[k])
            if shape[k] > 1:
                res_max = max(res_max, d.res64[k])
        gt_native = None
        if self.out_of_bounds_strategy == OutOfBoundsStrategy.LOOKUP_GT_SDF:
            gt_native = self.gt_sdf.native_desc(self._cdev) if isinstance(self.gt_sdf, MeshSDF) else None
            if gt_native is not None:
                flags |= nat.PVB_GRID_OOB_GT
                self.gt_sdf.obj_factory._fill_mesh_part(d, self.gt_sdf.obj_factory._device_state(self._cdev))
                flags |= d.flags & nat.PVB_MESH_CLOSED
                for k in range(3):       # _fill_mesh_part rewrote the box from the mesh; keep self.bb
                    d.bb_min[k] = float(bb[k, 0])
                    d.bb_max[k] = float(bb[k, 1])
        self._gt_in_kernel = gt_native is not None
        # Pruning bound for the composed kernels: table value at the nearest voxel of q is never below
        # aabb_distance(q) - prune_margin.  t is measured on the table itself, so the bound holds for any table.
        if self.out_of_bounds_strategy == OutOfBoundsStrategy.BOUNDING_BOX:
            margin = grid_prune_margin(self.voxels.raw_data.reshape(shape), lo, hi, bb)
            if math.isfinite(margin):
                d.prune_margin = margin
                flags |= nat.PVB_GRID_PRUNE_OK
        if self.interpolation == "trilinear":
            if self.out_of_bounds_strategy != OutOfBoundsStrategy.BOUNDING_BOX:
                raise ValueError("interpolation='trilinear' supports OutOfBoundsStrategy.BOUNDING_BOX only")
            flags = (flags | nat.PVB_GRID_TRILINEAR) & ~nat.PVB_GRID_PRUNE_OK
        d.flags = flags
        return d

    def native_desc(self, device):
        device = nat.compute_device(device)
        if device != self._cdev or self.interpolation != "nearest":
            return None          # the fused composition kernels implement the reference's nearest-voxel rule only
        if self.out_of_bounds_strategy == OutOfBoundsStrategy.LOOKUP_GT_SDF and not self._gt_in_kernel:
            return None
        return self._desc

    def surface_bounding_box(self, **kwargs):
        return self.gt_sdf.surface_bounding_box(**kwargs)

    def _fallback_sdf_value_func(self, *args, **kwargs):
        sdf_val, _ = self.gt_sdf(*args, **kwargs)
        return sdf_val.to(device=self._cdev)

    def _lookup(self, points, want_val=True, want_outside=False, surface_level=0., want_index=False):
        device = self._cdev
        with nat.on_device(device):
            p = nat.as_f32_points(points, device)
            n = p.shape[0]
            val = torch.empty(n, dtype=torch.float32, device=device) if want_val else None
            grad = torch.empty(n, 3, dtype=torch.float32, device=device) if want_val else None
            outside = torch.empty(n, dtype=torch.uint8, device=device) if want_outside else None
            index = torch.empty(n, dtype=torch.int64, device=device) if want_index else None
            L = nat.lib()
            if self.interpolation == "trilinear" and (want_outside or want_index):
                # keys / occupancy always follow the reference's nearest-voxel rule; values come from the interpolant
                plain = self._desc.copy()
                plain.flags &= ~nat.PVB_GRID_TRILINEAR
                nat.check(L.pvb_grid_lookup(ctypes.byref(plain), nat.ptr(p), n, None, None, nat.ptr(outside),
                                            float(surface_level), nat.ptr(index), nat.stream_ptr(device)),
                          "pvb_grid_lookup")
                if want_val:
                    nat.check(L.pvb_grid_lookup(ctypes.byref(self._desc), nat.ptr(p), n, nat.ptr(val), nat.ptr(grad),
                                                None, 0.0, None, nat.stream_ptr(device)), "pvb_grid_lookup")
            else:
                nat.check(L.pvb_grid_lookup(ctypes.byref(self._desc), nat.ptr(p), n, nat.ptr(val), nat.ptr(grad),
                                            nat.ptr(outside), float(surface_level), nat.ptr(index),
                                            nat.stream_ptr(device)), "pvb_grid_lookup")
        return p, val, grad, outside, index

    #: host batches at least this large are streamed through the GPU in chunks (_HostChunkStream: copy-in / lookup /
    #: copy-out overlapped on three streams); smaller ones take one H2D copy, one launch, one D2H copy
    host_pipeline_min_points = 1 << 22
    host_pipeline_chunk = 1 << 21

    def _host_pipeline(self, points):
        """Host tensor in, pinned host (values, gradients) out through the chunk pipeline."""
        device = self._cdev
        st = getattr(self, "_pipe_state", None)
        if st is None or st.chunk != self.host_pipeline_chunk or st.device != device:
            st = self._pipe_state = _HostChunkStream(device, self.host_pipeline_chunk, ((), (3,)))
        L = nat.lib()

        def launch(d_in, m, d_out):
            nat.check(L.pvb_grid_lookup(ctypes.byref(self._desc), nat.ptr(d_in), m, nat.ptr(d_out[0]), nat.ptr(d_out[1]),
                                        None, 0.0, None, nat.stream_ptr(device)), "pvb_grid_lookup")

        val_h, grad_h = st.run(points, launch)
        return val_h, grad_h

    def __call__(self, points_in_object_frame):
        lead = tuple(points_in_object_frame.shape[:-1])
        gt_outside_kernel = (self.out_of_bounds_strategy == OutOfBoundsStrategy.LOOKUP_GT_SDF
                             and not self._gt_in_kernel)
        if (torch.is_tensor(points_in_object_frame) and points_in_object_frame.device.type == "cpu"
                and torch.device(self.device).type == "cpu" and not gt_outside_kernel and not self.debug_check_sdf
                and self.interpolation == "nearest"
                and points_in_object_frame.numel() // 3 >= self.host_pipeline_min_points):
            val, grad = self._host_pipeline(points_in_object_frame)
            dtype = points_in_object_frame.dtype
            if dtype != torch.float32:
                val, grad = val.to(dtype), grad.to(dtype)
            return val.reshape(lead), grad.reshape(*lead, 3)
        p, val, grad, _, index = self._lookup(points_in_object_frame, want_index=gt_outside_kernel or
                                              self.debug_check_sdf)
        if gt_outside_kernel:
            # ground-truth SDF without a native descriptor: evaluate it on the out-of-range points (host-synchronous)
            oob = index < 0
            if oob.any():
                v, g = self.gt_sdf(p[oob])
                val[oob] = v.to(device=p.device, dtype=torch.float32)
                grad[oob] = g.to(device=p.device, dtype=torch.float32)
        if self.debug_check_sdf:
            inb = index >= 0
            if self.out_of_bounds_strategy == OutOfBoundsStrategy.BOUNDING_BOX and (~inb).any():
                val_gt, grad_gt = self.gt_sdf(p[~inb])
                assert torch.all(val_gt - val[~inb] > -1e-6)     # under-approximates (sdf.py:578)
                cos = torch.cosine_similarity(grad_gt, grad[~inb], dim=-1)
                assert torch.all(cos > 0.7) and cos.mean() > 0.95
            val_gt = self._fallback_sdf_value_func(p)
            assert torch.all((torch.abs(val - val_gt) < self.resolution)[inb])
        dtype = points_in_object_frame.dtype if torch.is_tensor(points_in_object_frame) else torch.float
        return (nat.deliver(val, self.device, dtype).reshape(lead),
                nat.deliver(grad, self.device, dtype).reshape(*lead, 3))

    def outside_surface(self, points_in_object_frame, surface_level=0):
        lead = tuple(points_in_object_frame.shape[:-1])
        _, _, _, outside, _ = self._lookup(points_in_object_frame, want_val=False, want_outside=True,
                                           surface_level=surface_level)
        return nat.deliver(outside, self.device, torch.bool).reshape(lead)

    def voxel_keys(self, points_in_object_frame):
        """Ravelled nearest-voxel key per point (int64, -1 where out of the cached range): the index/occupancy
        quantity that must be bit-exact against the reference (sdf.py:537-540)."""
        lead = tuple(points_in_object_frame.shape[:-1])
        _, _, _, _, index = self._lookup(points_in_object_frame, want_val=False, want_index=True)
        return nat.deliver(index, self.device).reshape(lead)

    def get_voxel_view(self, voxels: VoxelGrid = None, dtype=torch.float, device='cpu') -> GridView:
        if voxels is None:
            return self.voxels
        pts = voxels.get_voxel_center_points()
        sdf_val, sdf_grad = self.gt_sdf(pts.unsqueeze(0))
        sdf_val = sdf_val.to(device=self._cdev)
        cached_underlying_sdf = sdf_val.reshape([len(coord) for coord in voxels.coords])
        return GridView(cached_underlying_sdf, voxels.range_per_dim, invalid_value=self._fallback_sdf_value_func)


def sample_mesh_points(obj_factory: ObjectFactory = None, num_points=100, seed=0, name="",
                       clean_cache=False, dtype=torch.float, min_init_sample_points=200,
                       dbpath='model_points_cache.pkl', device="cpu", cache=None):
    """Area-uniform surface samples and their face normals (reference sdf.py:617-670).

    Same arguments, cache layout ({name: {seed: {num_points: (points, normals, None)}}}, torch.save) and return
    value `(points, normals, cache)` as the reference.  The samples come from a counter-based GPU generator, so the
    actual point set for a given seed differs from Open3D's mt19937 stream (the reference's stream is not
    reproducible outside Open3D either); it is deterministic per (mesh, seed, num_points).
    """
    # cache layout of the reference (sdf.py:620-634, 665): {name: {seed: {num_points: (points, normals, None)}}}, kept in
    # the caller's dict when one is passed, otherwise in the torch.save file `dbpath`
    caller_owns_cache = cache is not None
    if cache is None:
        cache = torch.load(dbpath) if os.path.exists(dbpath) else {}
    per_count = cache.setdefault(name, {}).setdefault(seed, {})
    hit = None if clean_cache else per_count.get(num_points)
    if hit is not None:
        points, normals = (None if t is None else t.to(device=device, dtype=dtype) for t in hit[:2])
        return points, normals, cache

    if obj_factory is None:
        raise RuntimeError(f"Expect model points to be cached for {name} {seed} {num_points} in {dbpath}")
    if obj_factory._mesh is None:
        obj_factory.precompute_sdf()

    cdev = nat.compute_device(device)
    sample_num_points = max(min_init_sample_points, 2 * num_points)       # sdf.py:650
    points = _sample_surface(obj_factory, sample_num_points, seed, cdev)
    # dispersion: keep a random subset of the over-sampled set (sdf.py:658)
    gen = torch.Generator(device=cdev)
    gen.manual_seed(int(seed))
    keep = torch.randperm(sample_num_points, generator=gen, device=cdev)[:num_points]
    points = points[keep]
    res = obj_factory.object_frame_closest_point(points, compute_normal=True)   # sdf.py:660
    normals = res.normal

    per_count[num_points] = points.cpu(), normals.cpu(), None
    if not caller_owns_cache:
        torch.save(cache, dbpath)
    return points.to(device=device, dtype=dtype), normals.to(device=device, dtype=dtype), cache


def _sample_surface(obj_factory, n, seed, device, return_faces=False):
    """n area-uniform fp64 surface points: stratified allocation of the samples to faces by cumulative area
    (face t receives round(cum_area_fraction(t) * n) - (samples so far), the Open3D rule behind sdf.py:654), then
    (1-sqrt(r1)) v0 + sqrt(r1)(1-r2) v1 + sqrt(r1) r2 v2 inside the face."""
    areas = obj_factory._mesh.triangle_areas()
    cum = np.cumsum(areas / areas.sum())
    upto = np.minimum(np.floor(cum * n + 0.5).astype(np.int64), n)
    upto = np.maximum.accumulate(upto)
    upto[-1] = n
    st = obj_factory._device_state(device)
    with torch.cuda.device(device):
        cum_dev = torch.from_numpy(upto).to(device)
        out = torch.empty(n, 3, dtype=torch.float64, device=device)
        face = torch.empty(n, dtype=torch.int32, device=device) if return_faces else None
        nat.check(nat.lib().pvb_mesh_sample(nat.ptr(st["v64"]), nat.ptr(st["faces"]), len(areas), nat.ptr(cum_dev),
                                            n, int(seed) & 0xFFFFFFFFFFFFFFFF, nat.ptr(out), nat.ptr(face),
                                            nat.stream_ptr(device)), "pvb_mesh_sample")
    return (out, face) if return_faces else out
