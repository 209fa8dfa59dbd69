"""CPU check of the exactness condition behind the composed kernels' pruning (pvb_kernels.cu: a sub-SDF is skipped
when dist(q, bb) - prune_margin exceeds the running minimum): with the margin `grid_prune_margin` measures on the
table, the value the REFERENCE's CachedSDF returns (oracle port on the reference-built golden tables) is never
below dist(q, bb) - margin, in range or out of range.  A violated bound would make the fused argmin differ from
sdf.py:392-433; a loose one only costs speed."""
import numpy as np
import pytest
import torch

from helpers import golden, port_mesh
from oracle import port
from pytorch_volumetric_b200.sdf import grid_prune_margin


def _aabb_distance(q, bb):
    lo, hi = torch.as_tensor(bb[:, 0]), torch.as_tensor(bb[:, 1])
    return torch.clamp(torch.maximum(lo - q.double(), q.double() - hi), min=0).norm(dim=-1)


@pytest.mark.parametrize("name", ["probe", "drill"])
def test_reference_values_respect_the_prune_bound(name, oracle_lib):
    z = golden(f"ref_cachedsdf_{name}")
    shape = tuple(int(s) for s in z["table_shape"])
    table = torch.from_numpy(z["table_val"]).reshape(shape)
    c = port.CachedSDFPort(name, float(z["resolution"]), z["range_in"], port.MeshSDFPort(port_mesh(name)),
                           tables=(table, torch.from_numpy(z["table_grad"])))
    lo = [float(r[0]) for r in c.ranges]; hi = [float(r[1]) for r in c.ranges]
    bb = np.asarray(z["bb"], dtype=np.float32)      # fp32 like the cast at sdf.py:556-557
    margin = grid_prune_margin(table, lo, hi, bb)
    cell_diag = float(z["resolution"]) * 3 ** 0.5
    assert 0.5 * cell_diag < margin < 2.5 * cell_diag, margin      # tight: about one cell, so pruning stays useful
    g = torch.Generator().manual_seed(0)
    span = torch.tensor(hi) - torch.tensor(lo)
    q = torch.cat([torch.tensor(lo) - 0.5 * span + 2.0 * span * torch.rand(300_000, 3, generator=g),   # in and out
                   torch.from_numpy(z["q"])]).float()
    v, _ = c(q)
    slack = v.double() - (_aabb_distance(q, bb.astype(np.float64)) - margin)
    assert float(slack.min()) >= 0.0, float(slack.min())
    # the bound is attained to within a cell somewhere (it is measured, not guessed)
    assert float(slack.min()) < 2.5 * cell_diag


def test_margin_adapts_to_adversarial_tables():
    """The bound holds for ANY table contents because t is measured: a table far below dist(., bb) gets a margin
    that covers it."""
    lo, hi, shape = [-1.0] * 3, [1.0] * 3, (21, 21, 21)
    bb = np.array([[-0.2, 0.2]] * 3, dtype=np.float32)
    table = torch.full(shape, -5.0)
    m = grid_prune_margin(table, lo, hi, bb)
    corner_lb = float(np.sqrt(3 * 0.8 ** 2))
    assert m >= corner_lb + 5.0
