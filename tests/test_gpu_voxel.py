"""Voxel containers (reference voxel.py:42-171, sdf.py:248-282) against vectors produced by the UNMODIFIED
reference source over the third-party shims (oracle/make_golden.py::make_voxel_vectors -> tests/golden/ref_voxel.npz).
GPU: tables on the device, read / written / listed by pvb_voxel_gather / pvb_voxel_scatter / pvb_compact_nonempty.
CPU tier: the same assertions on host tensors (the containers are device-agnostic like the reference's)."""
import numpy as np
import pytest
import torch

from helpers import golden

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


def _rows(a):
    """rows of an (N, d) float array as a sorted list of tuples rounded to 1e-6 (order-free comparison)."""
    return sorted(map(tuple, np.round(np.asarray(a, dtype=np.float64), 6).tolist()))


@pytest.mark.parametrize("device", DEVICES)
def test_voxel_grid_vs_reference_golden(device):
    import pytorch_volumetric_b200 as pv
    z = golden("ref_voxel")
    vg = pv.VoxelGrid(0.05, [tuple(r) for r in z["vg_box"]], device=device)
    p = torch.from_numpy(z["vg_pts"]).to(device)
    val = torch.from_numpy(z["vg_val"]).to(device)
    vg[p] = val
    data = vg.get_voxel_values().cpu().numpy()
    assert data.shape == z["vg_data"].shape
    # occupancy: bit-exact (the nearest-cell index rule)
    assert np.array_equal(data != 0, z["vg_data"] != 0)
    # cells written by exactly one point hold exactly that value; cells hit by several hold one of theirs
    # (index_put with duplicate indices: which writer wins is unspecified in torch as well)
    ref_view = pv.VoxelGrid(0.05, [tuple(r) for r in z["vg_box"]])
    keys = ref_view.voxels.ensure_index_key(torch.from_numpy(z["vg_pts"]))
    shape = torch.tensor(data.shape)
    ok = ((keys >= 0) & (keys < shape)).all(-1)
    flat = ref_view.voxels.ravel_multi_index(keys[ok], data.shape).numpy()
    vals = z["vg_val"][ok.numpy()]
    uniq, counts = np.unique(flat, return_counts=True)
    single = np.isin(flat, uniq[counts == 1])
    assert np.array_equal(data.reshape(-1)[flat[single]], vals[single])
    for cell in uniq[counts > 1][:200]:
        assert data.reshape(-1)[cell] in vals[flat == cell]
    single_cells = set(uniq[counts == 1].tolist())
    # read back, including out-of-range points (-> invalid value 0); only cells with one writer are comparable
    q = torch.from_numpy(z["vg_q"]).to(device)
    got = vg[q].cpu().numpy()
    qk = ref_view.voxels.ensure_index_key(torch.from_numpy(z["vg_q"]))
    q_ok = ((qk >= 0) & (qk < shape)).all(-1).numpy()
    assert np.array_equal(got[~q_ok], z["vg_q_out"][~q_ok]) and (got[~q_ok] == 0).all()
    q_flat = ref_view.voxels.ravel_multi_index(qk[torch.from_numpy(q_ok)], data.shape).numpy()
    cmp = np.array([c in single_cells or z["vg_data"].reshape(-1)[c] == 0 for c in q_flat])
    assert np.array_equal(got[q_ok][cmp], z["vg_q_out"][q_ok][cmp])
    # listing: the same cells in the same (ascending flat index) order, positions exact up to the fp32 product
    pos, kv = vg.get_known_pos_and_values()
    assert pos.shape == z["vg_known_pos"].shape
    np.testing.assert_allclose(pos.cpu().numpy(), z["vg_known_pos"], atol=1e-6)
    # resize_to_fit: same range / shape; reads unchanged
    vg.resize_to_fit()
    np.testing.assert_allclose(np.array(vg.range_per_dim, dtype=np.float64), z["vg_fit_range"], atol=1e-6)
    assert tuple(vg.get_voxel_values().shape) == tuple(z["vg_fit_shape"])
    got2 = vg[q].cpu().numpy()
    assert np.array_equal(got2 != 0, z["vg_fit_q_out"] != 0)


@pytest.mark.parametrize("device", DEVICES)
def test_expanding_grid_vs_reference_golden(device):
    import pytorch_volumetric_b200 as pv
    z = golden("ref_voxel")
    ev = pv.ExpandingVoxelGrid(0.1, [(0, 1), (0, 1), (0, 1)], device=device)
    e1, e2 = torch.from_numpy(z["ev_p1"]).to(device), torch.from_numpy(z["ev_p2"]).to(device)
    ev[e1] = torch.tensor([4.0, 5.0], device=device)
    ev[e2] = torch.tensor([7.0, 8.0], device=device)
    np.testing.assert_allclose(np.array(ev.range_per_dim, dtype=np.float64), z["ev_range"], atol=1e-6)
    assert tuple(ev.get_voxel_values().shape) == tuple(z["ev_shape"])
    assert np.array_equal(ev[torch.cat((e1, e2))].cpu().numpy(), z["ev_read"])


@pytest.mark.parametrize("device", DEVICES)
def test_voxel_down_sample_vs_reference_golden(device):
    """tests/test_voxel_sdf.py:8-29 of the reference plus ranged / flat variants: the SET of occupied cell centres."""
    import pytorch_volumetric_b200 as pv
    z = golden("ref_voxel")
    pts = torch.from_numpy(z["ds_pts"]).to(device)
    for key, kw in (("ds_02", dict(resolution=0.2)), ("ds_007", dict(resolution=0.07)),
                    ("ds_ranged", dict(resolution=0.1, range_per_dim=z["ds_range"]))):
        red = pv.voxel_down_sample(pts, **kw)
        assert red.device.type == device
        assert red.shape == z[key].shape, key
        assert _rows(red.cpu().numpy()) == _rows(z[key]), key
    flat = torch.cat((pts[:, :2], torch.zeros(len(pts), 1, device=device)), dim=1)
    red = pv.voxel_down_sample(flat, 0.15, range_per_dim=z["ds_flat_range"], ignore_flat_dim=True)
    assert _rows(red.cpu().numpy()) == _rows(z["ds_flat"])
    # the reference's own invariant (tests/test_voxel_sdf.py:27-29)
    red = pv.voxel_down_sample(pts, 0.2)
    f = torch.sin(red[:, 0]) + 2 * torch.cos(red[:, 1])
    assert red.shape[0] < pts.shape[0] * (4 / 100) / 0.2 and torch.allclose(f, red[:, 2], atol=0.4)


@pytest.mark.gpu
def test_filtered_points_vs_reference_golden(tmp_path):
    """ObjectFrameSDF.get_filtered_points (sdf.py:273-282) on the probe's CachedSDF voxel view: the interior cells."""
    import pytorch_volumetric_b200 as pv
    from helpers import pv_factory
    z = golden("ref_voxel")
    obj = pv_factory("probe")
    shape = [int(v) for v in z["fp_shape"]]
    ranges = pv.get_divisible_range_by_resolution(0.002, obj.bounding_box(padding=0.01))
    np.testing.assert_allclose(np.array(ranges), z["fp_ranges"], atol=1e-12)
    path = str(tmp_path / "c.pkl")
    torch.save({f"probe 0.002 {tuple(ranges)}": (torch.from_numpy(z["fp_table"]).reshape(shape),
                                                  torch.zeros(int(np.prod(shape)), 3))}, path)
    cached = pv.CachedSDF("probe", 0.002, obj.bounding_box(padding=0.01), pv.MeshSDF(obj), device="cuda",
                          cache_path=path)
    inner = cached.get_filtered_points(lambda v: v < -0.001)
    assert inner.shape == z["fp_interior"].shape
    np.testing.assert_allclose(inner.cpu().numpy(), z["fp_interior"], atol=1e-6)     # same cells, same order


@pytest.mark.gpu
def test_voxel_kernels_edge_cases():
    """2-D grids, bool grids, scalar writes, empty inputs, NaN points, ordered compaction across block boundaries."""
    import pytorch_volumetric_b200 as pv
    from pytorch_volumetric_b200.voxel import nonempty_indices
    g2 = pv.VoxelGrid(0.5, [(-2, 2), (0, 3)], device="cuda")
    p = torch.tensor([[0.1, 0.2], [1.9, 2.9], [float("nan"), 1.0], [9.0, 9.0]], device="cuda")
    g2[p] = 2.0
    assert g2[p].tolist() == [2.0, 2.0, 0.0, 0.0]
    gb = pv.VoxelGrid(0.25, [(-1, 1), (-1, 1), (-1, 1)], device="cuda", dtype=torch.bool)
    gb[torch.rand(500, 3, device="cuda") * 2 - 1] = 1
    pos, val = gb.get_known_pos_and_values()
    assert val.dtype == torch.bool and bool(val.all()) and pos.shape[0] == int(gb.get_voxel_values().sum())
    gb[torch.zeros(0, 3, device="cuda")] = 1
    assert gb[torch.zeros(0, 3, device="cuda")].shape == (0,)
    big = torch.zeros(1_000_003, device="cuda")
    idx = torch.randperm(1_000_003, device="cuda")[:12345].sort().values
    big[idx] = 1.5
    assert torch.equal(nonempty_indices(big), idx)
    assert nonempty_indices(torch.zeros(5000, device="cuda")).numel() == 0
