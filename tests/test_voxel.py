"""CPU: voxel containers (reference voxel.py:28-171) -- the invariants of the reference's own
tests/test_voxel_sdf.py plus container semantics.  Their third-party index arithmetic is restated ("parity
unpinned"), so these are invariant tests, not golden comparisons."""
import numpy as np
import torch

import pytorch_volumetric_b200 as pv


def test_voxel_down_sample_reference_invariants():
    N = 100

    def f(x, y):
        return torch.sin(x) + 2 * torch.cos(y)

    x = torch.linspace(-2, 2, N)
    xx, yy = torch.meshgrid(x, x, indexing="ij")
    pts = torch.stack((xx.flatten(), yy.flatten(), f(xx, yy).flatten()), dim=-1)
    new_resolution = 0.2
    reduce_factor = (4 / N) / new_resolution
    red = pv.voxel_down_sample(pts, new_resolution)
    assert red.shape[0] < pts.shape[0] * reduce_factor                     # tests/test_voxel_sdf.py:27
    assert torch.allclose(f(red[:, 0], red[:, 1]), red[:, 2], atol=new_resolution * 2)   # :29
    # every input point is within half a cell (per axis) of some output cell centre
    d = (pts[:, None, :] - red[None, :, :]).abs().amax(-1).amin(1)
    assert d.max() <= new_resolution / 2 + 1e-5
    # idempotent on its own output
    again = pv.voxel_down_sample(red, new_resolution)
    assert again.shape == red.shape
    flat = pv.voxel_down_sample(torch.cat([pts[:, :2], torch.zeros(len(pts), 1)], dim=1), 0.2,
                                range_per_dim=np.array([[-3, 3], [-3, 3], [0, 0]]), ignore_flat_dim=True)
    assert flat.shape[1] == 3 and (flat[:, 2] == 0).all()
    assert pv.voxel_down_sample(pts[:0], 0.2).shape == (0, 3)


def test_voxel_grid_set_get_and_resize():
    g = pv.VoxelGrid(0.1, [(-1, 1), (-1, 1), (0, 0.5)])
    p = torch.tensor([[0.0, 0.0, 0.1], [0.52, -0.31, 0.4], [5.0, 5.0, 5.0]])
    g[p] = torch.tensor([1.0, 2.0, 3.0])                 # the last one is out of range and dropped
    assert torch.equal(g[p[:2]], torch.tensor([1.0, 2.0]))
    assert g[p[2:]].item() == 0                           # invalid value
    pos, val = g.get_known_pos_and_values()
    assert pos.shape == (2, 3) and sorted(val.tolist()) == [1.0, 2.0]
    assert (pos - torch.tensor([[0.0, 0.0, 0.1], [0.5, -0.3, 0.4]])).abs().max() < 1e-6
    g.resize_to_fit()
    assert g.get_voxel_values().numel() < 21 * 21 * 6
    assert torch.equal(g[p[:2]], torch.tensor([1.0, 2.0]))


def test_expanding_grid_and_voxel_set():
    g = pv.ExpandingVoxelGrid(0.1, [(0, 1), (0, 1), (0, 1)])
    g[torch.tensor([[0.5, 0.5, 0.5]])] = torch.tensor([4.0])
    g[torch.tensor([[2.03, -0.47, 0.5]])] = torch.tensor([7.0])          # outside: the grid grows in whole cells
    assert g.range_per_dim[0][1] >= 2.0 and g.range_per_dim[1][0] <= -0.4
    assert g[torch.tensor([[0.5, 0.5, 0.5]])].item() == 4.0 and g[torch.tensor([[2.03, -0.47, 0.5]])].item() == 7.0
    s = pv.VoxelSet(torch.zeros(1, 3), torch.ones(1))
    s[torch.ones(2, 3)] = torch.tensor([2.0, 3.0])
    pos, val = s.get_known_pos_and_values()
    assert pos.shape == (3, 3) and val.tolist() == [1.0, 2.0, 3.0]
    import pytest
    with pytest.raises(RuntimeError):
        s[torch.zeros(1, 3)]
