"""Grid geometry helpers and the value-range view used by CachedSDF.

Reference: src/pytorch_volumetric/voxel.py:10-25 (these two helpers define the
exact voxel lattice of the CachedSDF tables) and the third-party
`multidim_indexing.TorchMultidimView` the reference wraps its value table in
(sdf.py:521-522).  `GridView` reproduces the surface of that view that the SDF
path and its callers use (`raw_data`, `shape`, `ensure_index_key`,
`ravel_multi_index`, `get_valid_values`, `ensure_value_key`, `view[pts]`).
"""
import abc
import ctypes

import numpy as np
import torch


def get_divisible_range_by_resolution(resolution, range_per_dim):
    """Snap every (low, high) so that the span is a whole number of cells."""
    out = []
    for low, high in range_per_dim:
        n = round((high - low) / resolution)
        out.append((low, low + n * resolution))
    return out


def get_coordinates_and_points_in_grid(resolution, range_per_dim, dtype=torch.float, device='cpu', get_points=True):
    """Per-axis coordinates low, low+res, ... (<= high + 0.9 res) and their cartesian product (C order)."""
    coords = [torch.arange(low, high + 0.9 * resolution, resolution, dtype=dtype, device=device)
              for low, high in range_per_dim]
    pts = torch.cartesian_prod(*coords) if get_points else None
    return coords, pts


def range_dtype(value_ranges):
    """dtype torch infers for torch.tensor([min(r) for r in ranges]): fp64 for numpy scalars, fp32 for
    Python floats -- this decides whether the reference's index arithmetic runs in fp64 or fp32."""
    return torch.tensor([min(r) for r in value_ranges]).dtype


def nonempty_indices(flat, empty=0):
    """Ascending flat indices of the elements of a 1-D tensor that differ from `empty` -- the compaction behind
    get_known_pos_and_values / voxel_down_sample / get_filtered_points.  float32 and bool tensors on a CUDA device go
    through pvb_compact_nonempty (block counts, scan, ordered write); anything else is host-side torch."""
    if flat.is_cuda and flat.dtype in (torch.float32, torch.bool, torch.uint8) and flat.is_contiguous() \
            and 0 < flat.numel() < 2 ** 32:
        from . import _native as nat
        data = flat.view(torch.uint8) if flat.dtype == torch.bool else flat
        n = data.numel()
        dev = flat.device
        with torch.cuda.device(dev):
            L = nat.lib()
            ws = torch.empty(int(L.pvb_compact_workspace(n)), dtype=torch.uint8, device=dev)
            count = torch.zeros(1, dtype=torch.int64, device=dev)
            eb = 4 if data.dtype == torch.float32 else 1
            nat.check(L.pvb_compact_nonempty(nat.ptr(data), n, eb, float(empty), nat.ptr(ws), 0, None, nat.ptr(count),
                                             nat.stream_ptr(dev)), "pvb_compact_nonempty")
            k = int(count.item())
            out = torch.empty(k, dtype=torch.int64, device=dev)
            if k:
                nat.check(L.pvb_compact_nonempty(nat.ptr(data), n, eb, float(empty), nat.ptr(ws), k, nat.ptr(out),
                                                 nat.ptr(count), nat.stream_ptr(dev)), "pvb_compact_nonempty")
        return out
    return (flat != empty).nonzero().reshape(-1)


class GridView:
    """Dense n-d value table addressed by real-valued coordinates (nearest cell).

    Tables of float32 / bool on a CUDA device are read, written and listed by the voxel kernels of libpvb.so
    (pvb_voxel_gather / pvb_voxel_scatter, include/pvb.h) when the points are float32; host-resident tables -- the
    reference's default device -- are ordinary torch tensors and use torch indexing."""

    def __init__(self, source, value_ranges, invalid_value=-1):
        self.device = source.device
        self.dtype = source.dtype
        self.shape = source.shape
        self.dim = source.dim()
        self._d = source.reshape(-1)
        self.invalid_value = invalid_value
        self._min = torch.tensor([min(r) for r in value_ranges], device=self.device)
        self._max = torch.tensor([max(r) for r in value_ranges], device=self.device)
        self._resolution = (self._max - self._min) / (torch.tensor(self.shape, device=self.device) - 1)

    @property
    def raw_data(self):
        return self._d

    def ensure_index_key(self, key, force=False):
        return torch.round((key - self._min) / self._resolution).to(torch.long)

    def ensure_value_key(self, key, force=False):
        if key.shape[-1] == 1 and self.dim > 1:      # flat (ravelled) indices, e.g. from raw_data.nonzero()
            key = torch.stack(torch.unravel_index(key.reshape(-1), tuple(self.shape)), dim=-1)
        return key * self._resolution + self._min

    def get_valid_values(self, key):
        return torch.all((self._min <= key) & (key <= self._max), dim=-1)

    @staticmethod
    def ravel_multi_index(key, shape):
        flat = key[..., 0]
        for d in range(1, len(shape)):
            flat = flat * shape[d] + key[..., d]
        return flat

    # -- kernel path ---------------------------------------------------------------
    def _native_ok(self, pts):
        return (self._d.is_cuda and self.dim <= 3 and self._d.dtype in (torch.float32, torch.bool)
                and torch.is_tensor(pts) and pts.dtype == torch.float32 and pts.shape[-1] == self.dim
                and self._d.is_contiguous())

    def _geom(self):
        g = getattr(self, "_geom_cache", None)
        if g is None:
            lo = self._min.double().cpu().tolist()
            res = self._resolution.double().cpu().tolist()
            g = ((ctypes.c_double * self.dim)(*lo), (ctypes.c_double * self.dim)(*res),
                 (ctypes.c_int32 * self.dim)(*[int(v) for v in self.shape]),
                 1 if self._min.dtype == torch.float32 else 0)
            self._geom_cache = g
        return g

    def _data_u8(self):
        return self._d.view(torch.uint8) if self._d.dtype == torch.bool else self._d

    def __getitem__(self, pts):
        if self._native_ok(pts):
            from . import _native as nat
            lead = pts.shape[:-1]
            dev = self._d.device
            p = pts.reshape(-1, self.dim).to(dev).contiguous()
            n = p.shape[0]
            lo, res, dims, f32 = self._geom()
            data = self._data_u8()
            out = torch.empty(n, dtype=data.dtype, device=dev)
            fill = 0.0 if callable(self.invalid_value) else float(self.invalid_value)
            valid = torch.empty(n, dtype=torch.uint8, device=dev) if callable(self.invalid_value) else None
            with torch.cuda.device(dev):
                nat.check(nat.lib().pvb_voxel_gather(self.dim, lo, res, dims, f32, nat.ptr(p), n,
                                                     4 if data.dtype == torch.float32 else 1, nat.ptr(data), fill,
                                                     nat.ptr(out), nat.ptr(valid), nat.stream_ptr(dev)), "pvb_voxel_gather")
            out = out.view(torch.bool) if self._d.dtype == torch.bool else out
            if valid is not None:
                bad = valid == 0
                if bool(bad.any()):
                    out[bad] = self.invalid_value(p[bad]).to(self.dtype).reshape(-1)
            return out.reshape(lead)
        lead = pts.shape[:-1]
        p = pts.reshape(-1, pts.shape[-1]).to(self.device)
        idx = self.ensure_index_key(p)
        upper = torch.tensor(self.shape, device=self.device)
        ok = torch.all((idx >= 0) & (idx < upper), dim=-1)
        out = torch.empty(p.shape[0], dtype=self.dtype, device=self.device)
        out[ok] = self._d[self.ravel_multi_index(idx[ok], self.shape)]
        if (~ok).any():
            if callable(self.invalid_value):
                out[~ok] = self.invalid_value(p[~ok]).to(self.dtype).reshape(-1)
            else:
                out[~ok] = self.invalid_value
        return out.reshape(lead)

    def __setitem__(self, pts, value):
        if self._native_ok(pts):
            from . import _native as nat
            dev = self._d.device
            p = pts.reshape(-1, self.dim).to(dev).contiguous()
            n = p.shape[0]
            lo, res, dims, f32 = self._geom()
            data = self._data_u8()
            per_point = None
            scalar = 0.0
            if torch.is_tensor(value) and value.numel() > 1:
                per_point = value.reshape(-1).to(device=dev, dtype=self._d.dtype).contiguous()
                if per_point.numel() != n:
                    raise ValueError(f"{per_point.numel()} values for {n} points")
                per_point = per_point.view(torch.uint8) if per_point.dtype == torch.bool else per_point
            else:
                scalar = float(value.item() if torch.is_tensor(value) else value)
            with torch.cuda.device(dev):
                nat.check(nat.lib().pvb_voxel_scatter(self.dim, lo, res, dims, f32, nat.ptr(p), n,
                                                      4 if data.dtype == torch.float32 else 1, nat.ptr(per_point),
                                                      scalar, nat.ptr(data), nat.stream_ptr(dev)), "pvb_voxel_scatter")
            return
        p = pts.reshape(-1, pts.shape[-1]).to(self.device)
        idx = self.ensure_index_key(p)
        upper = torch.tensor(self.shape, device=self.device)
        ok = torch.all((idx >= 0) & (idx < upper), dim=-1)
        if torch.is_tensor(value) and value.numel() > 1:
            value = value.reshape(-1)[ok]
        self._d[self.ravel_multi_index(idx[ok], self.shape)] = value


class Voxels(abc.ABC):
    """What every voxel container offers (reference voxel.py:28-39): read, write, and list what is stored."""

    @abc.abstractmethod
    def get_known_pos_and_values(self):
        """(N x d positions, N values) of the voxels that hold something."""

    @abc.abstractmethod
    def __getitem__(self, pts):
        """Values (N) stored at positions (N x d)."""

    @abc.abstractmethod
    def __setitem__(self, pts, value):
        """Store values (N) at positions (N x d)."""


def _extent(points):
    """Per-axis (min, max) of an N x d tensor as two numpy vectors."""
    return points.min(dim=0).values.cpu().numpy(), points.max(dim=0).values.cpu().numpy()


class VoxelGrid(Voxels):
    """Dense container over a range snapped to whole cells (reference voxel.py:42-91); also the default lattice of
    ObjectFrameSDF.get_voxel_view.  Cells holding `invalid_val` (0) count as empty."""

    invalid_val = 0

    def __init__(self, resolution, range_per_dim, dtype=torch.float, device='cpu'):
        self.resolution, self.dtype, self.device = resolution, dtype, device
        self._rebuild(range_per_dim)

    def _rebuild(self, box):
        """Allocate an empty grid over `box` (snapped); the previous contents are dropped."""
        snapped = get_divisible_range_by_resolution(self.resolution, box)
        self.coords, self.pts = get_coordinates_and_points_in_grid(self.resolution, snapped, device=self.device)
        self._data = torch.zeros(tuple(len(axis) for axis in self.coords), dtype=self.dtype, device=self.device)
        self.voxels = GridView(self._data, snapped, invalid_value=self.invalid_val)
        self.range_per_dim = np.array(snapped)

    def _create_voxels(self, resolution, range_per_dim):
        """The reference's name for re-creating the grid (voxel.py:50)."""
        self.resolution = resolution
        self._rebuild(range_per_dim)

    def _regrid(self, box):
        """Move the stored voxels to a new grid over `box`."""
        pos, val = self.get_known_pos_and_values()
        self._rebuild(box)
        if pos.numel():
            self.voxels[pos] = val

    def resize_to_fit(self):
        """Tighten the range around the stored voxels, one cell of slack on every side."""
        pos, _ = self.get_known_pos_and_values()
        if pos.numel() == 0:
            return
        lo, hi = _extent(pos)
        self._regrid(np.stack((lo - self.resolution, hi + self.resolution), axis=1))

    def get_voxel_center_points(self):
        return self.pts

    def get_voxel_values(self):
        return self._data

    def get_known_pos_and_values(self):
        flat = self.voxels.raw_data
        filled = nonempty_indices(flat, self.invalid_val)           # ordered compaction kernel on CUDA grids
        cells = torch.stack(torch.unravel_index(filled, tuple(self.voxels.shape)), dim=-1)
        return self.voxels.ensure_value_key(cells), flat[filled]

    def __getitem__(self, pts):
        return self.voxels[pts]

    def __setitem__(self, pts, value):
        self.voxels[pts] = value


class ExpandingVoxelGrid(VoxelGrid):
    """VoxelGrid that grows, by whole cells, to cover whatever is written to it (reference voxel.py:94-117)."""

    def __setitem__(self, pts, value):
        if pts.numel() > 0:
            # overshoot per axis in the points' own precision, like the reference's tensor - scalar arithmetic
            box = torch.as_tensor(self.range_per_dim, dtype=pts.dtype, device=pts.device)
            under = (box[:, 0] - pts.min(dim=0).values).double().cpu().numpy()
            over = (pts.max(dim=0).values - box[:, 1]).double().cpu().numpy()
            cells_below = np.ceil(np.clip(under, 0, None) / self.resolution)
            cells_above = np.ceil(np.clip(over, 0, None) / self.resolution)
            if cells_below.any() or cells_above.any():
                grown = self.range_per_dim + np.stack((-cells_below, cells_above), axis=1) * self.resolution
                if not np.allclose(grown, self.range_per_dim):
                    self._regrid(grown)         # the contents move to the larger grid first
        super().__setitem__(pts, value)


class VoxelSet(Voxels):
    """Sparse container: an explicit list of positions with their values (reference voxel.py:120-134)."""

    def __init__(self, positions, values):
        self.positions, self.values = positions, values

    def __getitem__(self, pts):
        raise RuntimeError("Cannot get arbitrary points on a voxel set")

    def __setitem__(self, pts, value):
        width = self.positions.shape[-1]
        self.positions = torch.cat((self.positions, pts.view(-1, width)), dim=0)
        self.values = torch.cat((self.values, value))

    def get_known_pos_and_values(self):
        return self.positions, self.values


def bounds_contain_another_bounds(outer_bounds, inner_bounds):
    outer, inner = np.asarray(outer_bounds), np.asarray(inner_bounds)
    return bool((outer[:, 0] <= inner[:, 0]).all() and (outer[:, 1] >= inner[:, 1]).all())


def voxel_down_sample(points, resolution, range_per_dim=None, ignore_flat_dim=False):
    """One point per occupied cell of a grid of the given resolution: the cell centre (reference voxel.py:141-171).
    The whole N x D cloud is scattered into a boolean grid in one indexed write and the set cells are read back;
    runs on the device the points live on."""
    if points.shape[0] == 0:
        return points
    lo, hi = _extent(points)
    margin = 2 * resolution
    cloud_box = np.stack((lo - margin, hi + margin), axis=1)
    # a caller-supplied range is only honoured when it does NOT already enclose the data (reference voxel.py:153-154)
    box = cloud_box if range_per_dim is None or bounds_contain_another_bounds(range_per_dim, cloud_box) \
        else range_per_dim
    drop_last = bool(ignore_flat_dim and box[-1][0] == box[-1][1])
    if drop_last:
        flat_coordinate = box[-1][0]
        box, points = box[:-1], points[..., :-1]
    occupancy = VoxelGrid(resolution, box, device=points.device, dtype=torch.bool)
    occupancy[points] = 1
    centres = occupancy.get_known_pos_and_values()[0].to(points.dtype)
    if drop_last:
        centres = torch.nn.functional.pad(centres, (0, 1), value=float(flat_coordinate))
    return centres
