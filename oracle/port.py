"""oracle/port.py -- CPU restatement of the reference's OWN code for the SDF path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Written from the
behaviour of the reference, each function citing the file:line it follows;
pinned by tests/test_oracle_pinning.py against (a) the unmodified reference
source executed over oracle/shims in the build container and (b) the golden
vectors that run produced (tests/golden/).  torch-CPU + numpy, fp32 where the
reference is fp32, fp64 where it is fp64.
"""
import math
import os

import numpy as np
import torch

from oracle import tp_open3d as o3d
from oracle.tp_multidim_indexing import TorchMultidimView
from oracle import tp_pytorch_kinematics as pk


# --------------------------------------------------------------- mesh object

class MeshPort:
    """ObjectFactory.precompute_sdf + bounding_box (sdf.py:80-89, 97-120)."""

    def __init__(self, path=None, scale=1.0, vis_frame_pos=(0, 0, 0), vis_frame_rot=(0, 0, 0, 1),
                 vertices=None, faces=None, name=None):
        self.name = name if name is not None else (path or "")
        self.scale = 1.0 if scale is None else scale
        if vertices is not None:
            mesh = o3d.TriangleMesh(vertices, faces)            # mesh= given: no scale/rot/translate (sdf.py:39-40)
        else:
            full = os.path.expanduser(path)
            if not os.path.exists(full):
                raise RuntimeError(f"Expected mesh file does not exist: {full}")
            mesh = o3d.read_triangle_mesh(full)
            S = np.eye(4)
            np.fill_diagonal(S[:3, :3], self.scale)             # sdf.py:105-107
            mesh.transform(S)
            x, y, z, w = vis_frame_rot                           # xyzw -> wxyz, sdf.py:110-111
            mesh.rotate(o3d.get_rotation_matrix_from_quaternion((w, x, y, z)), center=[0, 0, 0])
            mesh.translate(np.array(vis_frame_pos) * self.scale)  # sdf.py:113
        mesh.compute_triangle_normals()
        self.mesh = mesh
        self.face_normals = np.asarray(mesh.triangle_normals)   # fp64, sdf.py:119-120
        self.scene = o3d.RaycastingScene()
        self.scene.add_triangles(o3d.t.geometry.TriangleMesh.from_legacy(mesh))

    @property
    def vertices32(self):
        return self.scene._soup.verts

    @property
    def faces(self):
        return self.scene._soup.faces

    def bounding_box(self, padding=0., padding_ratio=0):
        lo = self.mesh.vertices.min(axis=0)
        hi = self.mesh.vertices.max(axis=0)
        ranges = np.stack([lo, hi], axis=1)
        ext = ranges[:, 1] - ranges[:, 0]
        ranges[:, 0] -= padding + padding_ratio * ext
        ranges[:, 1] += padding + padding_ratio * ext
        return ranges

    # -- ObjectFactory._do_object_frame_closest_point, sdf.py:122-172 ----------
    def closest_point(self, points, compute_normal=False, ray_noise=None):
        """points [..., N, 3] (tensor or ndarray).  Returns closest, distance,
        gradient, normal in the input dtype (sdf.py:125-132, 166).

        ray_noise: optional (M,3) array standing in for the reference's
        unseeded np.random.randn draw at sdf.py:149; None draws it from the
        numpy global generator exactly as the reference does."""
        if torch.is_tensor(points):
            dtype, device = points.dtype, points.device
            p = points.detach().cpu().numpy()
        else:
            dtype, device = torch.float, "cpu"
            p = np.asarray(points)
        lead = p.shape[:-1]
        p = p.reshape(-1, 3).astype(np.float32)

        res = self.scene.compute_closest_points(p)                       # sdf.py:134
        closest = res["points"].numpy()
        face = res["primitive_ids"].numpy()
        grad = closest - p                                               # sdf.py:139
        dist = np.linalg.norm(grad, axis=-1)                             # sdf.py:141
        nz = dist > 0
        grad[nz] = grad[nz] / dist[nz, None]                             # sdf.py:143-144

        far = self.bounding_box(padding=1.0)[:, 1]                       # sdf.py:147
        noise = np.random.randn(*p.shape) if ray_noise is None else np.asarray(ray_noise).reshape(p.shape)
        dest = (np.repeat(far[None], p.shape[0], axis=0) + 1e-4 * noise).astype(np.float32)   # sdf.py:149-150
        rays = np.concatenate([p, dest], axis=-1)                        # "destination" used as direction, sdf.py:152
        hits = self.scene.count_intersections(rays).numpy()
        inside = hits % 2 == 1                                           # sdf.py:154
        dist[inside] *= -1                                               # sdf.py:155
        grad[~inside] *= -1                                              # sdf.py:157

        shell = np.abs(dist) < 1e-3                                      # sdf.py:162
        grad[shell] = self.face_normals[face[shell]]                     # sdf.py:163-164 (fp64 -> fp32 store)

        def out(a):
            return torch.tensor(a, device=device, dtype=dtype).reshape(*lead, *a.shape[1:])

        normal = out(self.face_normals[face]) if compute_normal else None   # sdf.py:168-171
        return out(closest), out(dist), out(grad), normal


def mesh_sdf_call(mesh: MeshPort, points, ray_noise=None):
    """MeshSDF.__call__ (sdf.py:312-329): (distance, gradient) of the query."""
    _c, d, g, _n = mesh.closest_point(points, ray_noise=ray_noise)
    return d, g


class MeshSDFPort:
    def __init__(self, mesh: MeshPort):
        self.mesh = mesh

    def __call__(self, points):
        return mesh_sdf_call(self.mesh, points)

    def surface_bounding_box(self, **kw):
        return torch.tensor(self.mesh.bounding_box(**kw))


# ------------------------------------------------------------- grid helpers

def divisible_range(resolution, range_per_dim):
    """voxel.py:10-17: snap each span to a whole number of cells."""
    out = []
    for low, high in range_per_dim:
        cells = round((high - low) / resolution)
        out.append((low, low + cells * resolution))
    return out


def grid_coords_and_points(resolution, range_per_dim, dtype=torch.float, device="cpu", get_points=True):
    """voxel.py:20-25: per-axis arange then cartesian product (C order)."""
    coords = [torch.arange(lo, hi + 0.9 * resolution, resolution, dtype=dtype, device=device)
              for lo, hi in range_per_dim]
    return coords, (torch.cartesian_prod(*coords) if get_points else None)


# ---------------------------------------------------------------- CachedSDF

class CachedSDFPort:
    """CachedSDF.__init__ table build + __call__ + outside_surface
    (sdf.py:444-525, 535-571, 593-602).  No on-disk cache (out of the compute
    path); tables can be injected with `tables=(val, grad)`."""

    def __init__(self, object_name, resolution, range_per_dim, gt_sdf, tables=None, gt_lookup_oob=False):
        self.gt_sdf = gt_sdf
        self.resolution = resolution
        self.ranges = divisible_range(resolution, range_per_dim)            # sdf.py:481
        self.name = f"{object_name} {resolution} {tuple(self.ranges)}"     # sdf.py:484
        self.gt_lookup_oob = gt_lookup_oob
        if tables is None:
            coords, pts = grid_coords_and_points(resolution, self.ranges)   # sdf.py:502
            val, grad = gt_sdf(pts)                                         # sdf.py:503
            val = val.reshape([len(c) for c in coords])                     # sdf.py:504
            grad = grad.squeeze(0)                                          # sdf.py:505
        else:
            val, grad = tables
        self.voxels = TorchMultidimView(val, self.ranges, invalid_value=None)   # sdf.py:521
        self.voxels_grad = grad.squeeze()                                   # sdf.py:523
        self.bb = gt_sdf.surface_bounding_box()                             # sdf.py:525 (fp64 until first call)

    def surface_bounding_box(self, **kw):
        return self.gt_sdf.surface_bounding_box(**kw)

    def index_and_mask(self, points):
        keys = self.voxels.ensure_index_key(points)                         # sdf.py:537
        flat = self.voxels.ravel_multi_index(keys, self.voxels.shape)       # sdf.py:538
        inbound = self.voxels.get_valid_values(points)                      # sdf.py:540
        return keys, flat, inbound

    def __call__(self, points):
        keys, flat, inbound = self.index_and_mask(points)
        oob = ~inbound
        dtype = points.dtype
        val = torch.zeros(flat.shape, dtype=dtype)
        grad = torch.zeros(keys.shape, dtype=dtype)
        val[inbound] = self.voxels.raw_data[flat[inbound]].to(dtype)        # sdf.py:549
        grad[inbound] = self.voxels_grad[flat[inbound]].to(dtype)           # sdf.py:550
        p_out = points[oob]
        if self.gt_lookup_oob:                                              # sdf.py:553-554
            v, g = self.gt_sdf(p_out)
            val[oob], grad[oob] = v.to(dtype), g.to(dtype)
        else:                                                               # sdf.py:555-571
            bb = self.bb.to(dtype=dtype)
            below = bb[:, 0] - p_out
            below_on = below > 0
            below[~below_on] = 0
            above = p_out - bb[:, 1]
            above_on = above > 0
            above[~above_on] = 0
            delta = below + above
            delta[below_on] = -delta[below_on]
            dist = delta.norm(dim=-1)
            grad[oob] = delta / dist.unsqueeze(-1)
            val[oob] = dist
        return val, grad

    def outside_surface(self, points, surface_level=0):
        _keys, flat, inbound = self.index_and_mask(points)
        outside = torch.ones(flat.shape, dtype=torch.bool)                  # sdf.py:600
        outside[inbound] = self.voxels.raw_data[flat[inbound]] > surface_level   # sdf.py:601
        return outside


# --------------------------------------------------------------- SphereSDF

class SphereSDFPort:
    """sdf.py:285-299."""

    def __init__(self, radius):
        self.radius = radius

    def __call__(self, p):
        r = torch.linalg.norm(p, dim=-1)
        return r - self.radius, p / (r.unsqueeze(-1) + 1e-12)

    def surface_bounding_box(self, padding=0., padding_ratio=0.):
        L = self.radius + padding + padding_ratio * self.radius
        return torch.tensor([[-L, L], [-L, L], [-L, L]])


# -------------------------------------------------------------- ComposedSDF

class ComposedSDFPort:
    """ComposedSDF (sdf.py:332-433).  `obj_to_each` is a pk-style Transform3d
    holding S (or S*|A|, link-major) object->sub-frame transforms."""

    def __init__(self, sdfs, obj_to_each):
        self.sdfs = sdfs
        self.set_transforms(obj_to_each)

    def set_transforms(self, tsf, batch_dim=None):
        self.obj_to_link = tsf
        self.tsf_batch = batch_dim
        self.link_to_obj = []
        if tsf is None:
            return
        S = len(self.sdfs)
        n = len(tsf)
        if self.tsf_batch is None and n != S:
            self.tsf_batch = (n // S,)                     # sdf.py:378-379 (a float there; integer here)
        inv = tsf.get_matrix().inverse()                   # general 4x4 inverse, sdf.py:380
        for i in range(S):
            self.link_to_obj.append(pk.Transform3d(matrix=inv[self._slice(i)]))

    def _slice(self, i):
        if self.tsf_batch is None:
            return slice(i, i + 1)
        n = math.prod(list(self.tsf_batch))
        return slice(i * n, (i + 1) * n)

    def surface_bounding_box(self, **kw):
        """sdf.py:347-368 (transforms only the (min,max) corner pair)."""
        back = self.obj_to_link.inverse()
        per = []
        for i, s in enumerate(self.sdfs):
            corners = s.surface_bounding_box(**kw)
            corners = back[self._slice(i)].transform_points(
                corners.to(dtype=back.dtype, device=back.device).transpose(0, 1))
            if self.tsf_batch is not None and corners.dim() == 2:
                corners = corners.unsqueeze(0)
            per.append(corners)
        per = torch.stack(per)
        if self.tsf_batch is not None:
            dims = (0,) + tuple(range(2, per.dim() - 1))
        else:
            dims = tuple(range(per.dim() - 1))
        return torch.stack((per.amin(dim=dims), per.amax(dim=dims)), dim=-1)

    def __call__(self, points):
        shape = points.shape
        flat = points.reshape(-1, 3)
        S = len(self.sdfs)
        local = self.obj_to_link.transform_points(flat)             # sdf.py:399
        if self.tsf_batch is not None:
            local = local.reshape(S, *self.tsf_batch, *flat.shape)  # sdf.py:401-402
        vals, grads = [], []
        for i, s in enumerate(self.sdfs):
            v, g = s(local[i])                                      # sdf.py:407
            g = self.link_to_obj[i].transform_normals(g)            # sdf.py:409
            vals.append(v)
            grads.append(g)
        v = torch.cat(vals).reshape(S, -1)                          # sdf.py:414-418
        g = torch.cat(grads).reshape(S, -1, 3)
        which = torch.argmin(v, 0)                                  # first index on ties
        cols = torch.arange(v.shape[1])
        vv, gg = v[which, cols], g[which, cols]
        if self.tsf_batch is not None:                              # sdf.py:428-431
            vv = vv.reshape(*self.tsf_batch, *shape[:-1])
            gg = gg.reshape(*self.tsf_batch, *shape[:-1], 3)
        return vv, gg


# ----------------------------------------------------------------- RobotSDF

class RobotSDFPort:
    """RobotSDF (model_to_sdf.py:16-125).  `chain` is a pk-style serial chain;
    link_sdf_factory(MeshPort) -> SDF port object."""

    def __init__(self, chain, default_joint_config=None, path_prefix="", link_sdf_factory=MeshSDFPort,
                 mesh_loader=None):
        self.chain = chain
        self.dtype, self.device = chain.dtype, chain.device
        self.joint_names = chain.get_joint_parameter_names()
        self.link_names = []
        sdfs, offsets = [], []
        for fname in chain.get_frame_names(exclude_fixed=False):          # model_to_sdf.py:41-56
            frame = chain.find_frame(fname)
            for vis in frame.link.visuals:
                if vis.geom_type != "mesh":
                    continue
                if mesh_loader is not None:
                    mesh = mesh_loader(vis.geom_param[0], vis.geom_param[1])
                else:
                    mesh = MeshPort(os.path.join(path_prefix, vis.geom_param[0].replace("package://", "")
                                                 if path_prefix != "" else vis.geom_param[0]),
                                    scale=vis.geom_param[1], name=vis.geom_param[0])
                sdfs.append(link_sdf_factory(mesh))
                offsets.append(vis.offset)
                self.link_names.append(frame.link.name)
        self.offsets = offsets[0].stack(*offsets[1:]).to(device=self.device, dtype=self.dtype)   # model_to_sdf.py:58
        self.sdf = ComposedSDFPort(sdfs, None)
        self.set_joint_configuration(default_joint_config)

    def set_joint_configuration(self, q=None):
        M = len(self.joint_names)
        if q is None:
            q = torch.zeros(M, dtype=self.dtype, device=self.device)
        if q.dim() > 1:                                                    # model_to_sdf.py:94-98
            self.configuration_batch = q.shape[:-1]
            q = q.reshape(-1, M)
        else:
            self.configuration_batch = None
        fk = self.chain.forward_kinematics(q, end_only=False)              # model_to_sdf.py:99
        mats = torch.cat([fk[name].get_matrix() for name in self.link_names])   # link-major, :100-102,112
        off_inv = self.offsets.inverse()
        if self.configuration_batch is not None:                           # model_to_sdf.py:105-110
            m = off_inv.get_matrix()[(slice(None),) + (None,) * len(self.configuration_batch)]
            m = m.repeat(1, *self.configuration_batch, 1, 1)
            off_inv = pk.Transform3d(matrix=m.reshape(-1, 4, 4))
        self.object_to_link = off_inv.compose(pk.Transform3d(matrix=mats).inverse())   # :113
        self.sdf.set_transforms(self.object_to_link, batch_dim=self.configuration_batch)

    def surface_bounding_box(self, **kw):
        return self.sdf.surface_bounding_box(**kw)

    def __call__(self, points):
        return self.sdf(points)


def cache_link_sdf_factory_port(resolution=0.01, padding=0.1, **kw):
    """model_to_sdf.py:128-133."""

    def make(mesh: MeshPort):
        gt = MeshSDFPort(mesh)
        return CachedSDFPort(mesh.name, resolution, mesh.bounding_box(padding=padding), gt, **kw)

    return make


# ------------------------------------------------------------------ chamfer

def batch_chamfer_dist_port(world_to_object, pts_world, mesh: MeshPort = None, obj_sdf=None, scale=1000.):
    """chamfer.py:79-94."""
    pts_obj = pk.Transform3d(matrix=world_to_object).transform_points(pts_world)
    if obj_sdf is not None:
        d, _ = obj_sdf(pts_obj)
    elif mesh is not None:
        d = mesh.closest_point(pts_obj)[1]
    else:
        raise ValueError("Either obj_sdf or obj_factory must be given")
    return ((scale * d) ** 2).mean(dim=-1)


# ------------------------------------------------------- sample_mesh_points

def sample_mesh_points_port(mesh: MeshPort, num_points=100, seed=0, dtype=torch.float, min_init_sample_points=200):
    """sdf.py:639-670 without the pickle cache: seeded area-uniform samples,
    random subset, face normals of the closest faces."""
    from oracle.tp_arm_utils import rand
    with rand.SavedRNG():
        rand.seed(seed)
        o3d.utility.random.seed(seed)
        n0 = max(min_init_sample_points, 2 * num_points)                   # sdf.py:650
        pts = np.asarray(mesh.mesh.sample_points_uniformly(number_of_points=n0).points)
        pts = np.random.permutation(pts)[:num_points]                      # sdf.py:658
        normals = mesh.closest_point(pts, compute_normal=True)[3]          # sdf.py:660
    return torch.tensor(pts).to(dtype=dtype), normals.to(dtype=dtype)


# ------------------------------------------------- extension: trilinear lookup
def trilinear_lookup_port(table_val, ranges, bb, points):
    """CPU restatement of the OPT-IN extension CachedSDF(interpolation="trilinear") of pytorch_volumetric_b200
    (not reference behaviour; the reference only has the nearest-voxel rule above): value = trilinear
    interpolation of the 8 surrounding voxel values, gradient = analytic gradient of the interpolant; points
    failing all(min <= p <= max) take the AABB rule of sdf.py:555-571.  fp64 numpy."""
    val = table_val.double().numpy()
    n = np.array(val.shape)
    lo = np.array([float(min(r)) for r in ranges]); hi = np.array([float(max(r)) for r in ranges])
    res = (hi - lo) / np.maximum(n - 1, 1)
    p = points.double().numpy().reshape(-1, 3)
    inb = np.all((p >= lo) & (p <= hi), axis=1)
    u = (p - lo) / res
    c0 = np.clip(np.floor(u).astype(np.int64), 0, np.maximum(n - 2, 0))
    f = np.clip(u - c0, 0.0, 1.0)
    c1 = np.minimum(c0 + 1, n - 1)

    def v(ix, iy, iz):
        return val[ix, iy, iz]

    x0, y0, z0 = c0[:, 0], c0[:, 1], c0[:, 2]
    x1, y1, z1 = c1[:, 0], c1[:, 1], c1[:, 2]
    fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
    c00 = v(x0, y0, z0) * (1 - fz) + v(x0, y0, z1) * fz
    c01 = v(x0, y1, z0) * (1 - fz) + v(x0, y1, z1) * fz
    c10 = v(x1, y0, z0) * (1 - fz) + v(x1, y0, z1) * fz
    c11 = v(x1, y1, z0) * (1 - fz) + v(x1, y1, z1) * fz
    out_v = (c00 * (1 - fy) + c01 * fy) * (1 - fx) + (c10 * (1 - fy) + c11 * fy) * fx
    gx = ((c10 * (1 - fy) + c11 * fy) - (c00 * (1 - fy) + c01 * fy)) / res[0]
    gy = ((c01 - c00) * (1 - fx) + (c11 - c10) * fx) / res[1]
    dz00 = v(x0, y0, z1) - v(x0, y0, z0); dz01 = v(x0, y1, z1) - v(x0, y1, z0)
    dz10 = v(x1, y0, z1) - v(x1, y0, z0); dz11 = v(x1, y1, z1) - v(x1, y1, z0)
    gz = ((dz00 * (1 - fy) + dz01 * fy) * (1 - fx) + (dz10 * (1 - fy) + dz11 * fy) * fx) / res[2]
    out_g = np.stack([gx, gy, gz], axis=1)
    # out of range: distance / direction to the surface AABB
    bbn = np.asarray(bb, dtype=np.float64)
    below = np.maximum(bbn[:, 0] - p, 0); above = np.maximum(p - bbn[:, 1], 0)
    delta = np.where(bbn[:, 0] - p > 0, -(below + above), below + above)
    dist = np.linalg.norm(delta, axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        gdir = delta / dist[:, None]
    out_v = np.where(inb, out_v, dist)
    out_g = np.where(inb[:, None], out_g, gdir)
    return torch.from_numpy(out_v), torch.from_numpy(out_g), torch.from_numpy(inb)


# ------------------------------------------- extension: winding-number sign
def winding_number_port(vertices, faces, points, chunk=2000):
    """Exact generalized winding number (sum of Van Oosterom-Strackee solid angles / 4 pi), fp64, brute force.
    CPU restatement of the OPT-IN extension ObjectFactory.sign_mode = "winding" of pytorch_volumetric_b200 (not
    reference behaviour: the reference uses crossing parity, sdf.py:146-157)."""
    v = np.asarray(vertices, dtype=np.float32).astype(np.float64)      # the query structure holds fp32 vertices
    tri = v[np.asarray(faces)]
    q = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    out = np.empty(len(q))
    for s0 in range(0, len(q), chunk):
        d = tri[None] - q[s0:s0 + chunk, None, None, :]                 # (m, T, 3, 3)
        A, B, C = d[:, :, 0], d[:, :, 1], d[:, :, 2]
        la, lb, lc = (np.linalg.norm(x, axis=-1) for x in (A, B, C))
        num = np.einsum("mti,mti->mt", A, np.cross(B, C))
        den = la * lb * lc + (A * B).sum(-1) * lc + (B * C).sum(-1) * la + (C * A).sum(-1) * lb
        out[s0:s0 + chunk] = (2.0 * np.arctan2(num, den)).sum(axis=1) / (4.0 * np.pi)
    return out
