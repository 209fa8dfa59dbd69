"""Grid geometry helpers and the value-range view used by CachedSDF.

Reference: src/pytorch_volumetric/voxel.py:10-25 (these two helpers define the
exact voxel lattice of the CachedSDF tables) and the third-party
`multidim_indexing.TorchMultidimView` the reference wraps its value table in
(sdf.py:521-522).  `GridView` reproduces the surface of that view that the SDF
path and its callers use (`raw_data`, `shape`, `ensure_index_key`,
`ravel_multi_index`, `get_valid_values`, `ensure_value_key`, `view[pts]`).
"""
import abc
import copy
import math

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


class GridView:
    """Dense n-d value table addressed by real-valued coordinates (nearest cell)."""

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

    def __getitem__(self, pts):
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
        p = pts.reshape(-1, pts.shape[-1]).to(self.device)
        idx = self.ensure_index_key(p)
        upper = torch.tensor(self.shape, device=self.device)
        ok = torch.all((idx >= 0) & (idx < upper), dim=-1)
        if torch.is_tensor(value) and value.numel() > 1:
            value = value.reshape(-1)[ok]
        self._d[self.ravel_multi_index(idx[ok], self.shape)] = value


class Voxels(abc.ABC):
    """Interface of the voxel containers (reference voxel.py:28-39)."""

    @abc.abstractmethod
    def get_known_pos_and_values(self):
        """positions (N x d) and values (N) of the voxels that hold something"""

    @abc.abstractmethod
    def __getitem__(self, pts):
        """values (N) at positions (N x d)"""

    @abc.abstractmethod
    def __setitem__(self, pts, value):
        """store values (N) at positions (N x d)"""


class VoxelGrid(Voxels):
    """Dense voxel container over a snapped range (reference voxel.py:42-91); used as the default
    lattice of ObjectFrameSDF.get_voxel_view."""

    def __init__(self, resolution, range_per_dim, dtype=torch.float, device='cpu'):
        self.resolution = resolution
        self.invalid_val = 0
        self.dtype = dtype
        self.device = device
        self._create_voxels(resolution, range_per_dim)

    def _create_voxels(self, resolution, range_per_dim):
        self.range_per_dim = get_divisible_range_by_resolution(resolution, range_per_dim)
        self.coords, self.pts = get_coordinates_and_points_in_grid(resolution, self.range_per_dim, device=self.device)
        self._data = torch.zeros([len(c) for c in self.coords], dtype=self.dtype, device=self.device)
        self.voxels = GridView(self._data, self.range_per_dim, invalid_value=self.invalid_val)
        self.range_per_dim = np.array(self.range_per_dim)

    def resize_to_fit(self):
        """Shrink / move the grid so that it just contains the known voxels (one cell of slack per side)."""
        pos, val = self.get_known_pos_and_values()
        if pos.numel() == 0:
            return
        lo, hi = pos.min(dim=0).values, pos.max(dim=0).values
        box = copy.deepcopy(self.range_per_dim)
        for d in range(len(lo)):
            box[d] = (lo[d].item() - self.resolution, hi[d].item() + self.resolution)
        self._create_voxels(self.resolution, box)
        self.__setitem__(pos, val)

    def get_voxel_center_points(self):
        return self.pts

    def get_voxel_values(self):
        return self._data

    def get_known_pos_and_values(self):
        known = self.voxels.raw_data != self.invalid_val
        idx = known.nonzero()
        unravel = torch.stack(torch.unravel_index(idx.reshape(-1), tuple(self.voxels.shape)), dim=-1)
        return self.voxels.ensure_value_key(unravel), self.voxels.raw_data[idx.reshape(-1)]

    def __getitem__(self, pts):
        return self.voxels[pts]

    def __setitem__(self, pts, value):
        self.voxels[pts] = value


class ExpandingVoxelGrid(VoxelGrid):
    """VoxelGrid whose range grows, in whole cells, whenever a write falls outside it (reference voxel.py:94-117)."""

    def __setitem__(self, pts, value):
        if pts.numel() > 0:
            lo, hi = pts.min(dim=0).values, pts.max(dim=0).values
            box = copy.deepcopy(self.range_per_dim)
            for d in range(len(lo)):
                above = (hi[d] - self.range_per_dim[d][1]).item()
                below = (self.range_per_dim[d][0] - lo[d]).item()
                if above > 0:
                    box[d][1] += math.ceil(above / self.resolution) * self.resolution
                if below > 0:
                    box[d][0] -= math.ceil(below / self.resolution) * self.resolution
            if not np.allclose(box, self.range_per_dim):
                pos, val = self.get_known_pos_and_values()      # carry the contents over to the larger grid
                self._create_voxels(self.resolution, box)
                super().__setitem__(pos, val)
        return super().__setitem__(pts, value)


class VoxelSet(Voxels):
    """Explicit list of occupied positions and their values (reference voxel.py:120-134)."""

    def __init__(self, positions, values):
        self.positions = positions
        self.values = values

    def __getitem__(self, pts):
        raise RuntimeError("Cannot get arbitrary points on a voxel set")

    def __setitem__(self, pts, value):
        self.positions = torch.cat((self.positions, pts.view(-1, self.positions.shape[-1])), dim=0)
        self.values = torch.cat((self.values, value))

    def get_known_pos_and_values(self):
        return self.positions, self.values


def bounds_contain_another_bounds(outer_bounds, inner_bounds):
    outer_bounds, inner_bounds = np.asarray(outer_bounds), np.asarray(inner_bounds)
    return bool(np.all(outer_bounds[:, 0] <= inner_bounds[:, 0]) and np.all(outer_bounds[:, 1] >= inner_bounds[:, 1]))


def voxel_down_sample(points, resolution, range_per_dim=None, ignore_flat_dim=False):
    """Down-sample an N x D cloud to the centres of the occupied cells of a grid of the given resolution
    (reference voxel.py:141-171): all points are scattered into a boolean grid at once, the occupied cells are
    read back.  Runs on the device the points live on."""
    if points.shape[0] == 0:
        return points
    data_bounds = np.stack((points.min(dim=0)[0].cpu().numpy() - resolution * 2,
                            points.max(dim=0)[0].cpu().numpy() + resolution * 2)).T
    if range_per_dim is None or bounds_contain_another_bounds(range_per_dim, data_bounds):
        range_per_dim = data_bounds
    flat_z = ignore_flat_dim and range_per_dim[-1][0] == range_per_dim[-1][1]
    flat_z_val = range_per_dim[-1][0]
    if flat_z:
        range_per_dim = range_per_dim[:-1]
        points = points[..., :-1]
    grid = VoxelGrid(resolution, range_per_dim, device=points.device, dtype=torch.bool)
    grid[points] = 1
    pts, _ = grid.get_known_pos_and_values()
    pts = pts.to(points.dtype)
    if flat_z:
        pts = torch.cat((pts, torch.ones((pts.shape[0], 1), device=points.device, dtype=pts.dtype) * flat_z_val), dim=-1)
    return pts
