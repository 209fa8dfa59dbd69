"""CPU check of the fast voxel-index path of the grid kernels (pvb_device.cuh grid_eval): a numpy fp32 emulation of
the device arithmetic -- fp32 estimate, magic-add rint, certainty band -- against the reference's exact formula
round((p - min) / res) (TorchMultidimView.ensure_index_key, reference sdf.py:537-549), in both dtype modes, on
uniform points and on points placed a few ulps around every cell boundary.  The kernel takes the exact path
whenever the estimate is not certain, so "certain => same index" is what makes the key bit-exact."""
import numpy as np
import pytest

from pytorch_volumetric_b200.sdf import fast_index_band

MAGIC = np.float32(12582912.0)      # 1.5 * 2^23


def device_estimate(p, lo, hi, n):
    """(index, certain) as grid_eval computes them for one axis; p float32 array."""
    inv_res32, idx_certain = fast_index_band(lo, hi, n)
    min32 = np.float32(lo)
    q = (p - min32).astype(np.float32) * np.float32(inv_res32)
    q = q.astype(np.float32)
    m = (q + MAGIC).astype(np.float32)
    k = m.view(np.int32) - np.int32(0x4B400000)
    certain = np.abs((q - (m - MAGIC).astype(np.float32)).astype(np.float32)) <= np.float32(idx_certain)
    return k, certain


def exact_index(p, lo, hi, n, fp32_mode):
    """grid_axis_index_exact: the reference formula in the dtype torch infers for the range."""
    if fp32_mode:
        lo32, hi32 = np.float32(lo), np.float32(hi)
        res32 = np.float32((hi32 - lo32) / np.float32(n - 1))
        k = np.rint(((p - lo32).astype(np.float32) / res32).astype(np.float32))
    else:
        res64 = (hi - lo) / (n - 1)
        k = np.rint((p.astype(np.float64) - lo) / res64)
    return np.clip(k.astype(np.int64), 0, n - 1)


def boundary_points(lo, hi, n, ulps=4):
    """float32 values a few ulps either side of every cell boundary lo + (k + 1/2) res."""
    res = (hi - lo) / (n - 1)
    b = (lo + (np.arange(n - 1) + 0.5) * res).astype(np.float32)
    out = [b]
    up, down = b.copy(), b.copy()
    for _ in range(ulps):
        up = np.nextafter(up, np.float32(np.inf)); down = np.nextafter(down, np.float32(-np.inf))
        out += [up.copy(), down.copy()]
    return np.concatenate(out)


CASES = [
    # lo, hi, n                      (drill x axis at res 0.005; README link at 0.02 / pad 1.0; far from the origin;
    (-0.1713, 0.1987, 75),           #  fine and far: the band closes and everything goes the exact way)
    (-1.09, 1.11, 111),
    (-0.5, 0.5, 1001),
    (100.0, 101.0, 2001),
    (1000.0, 1001.0, 1001),
    (-3.0e-3, 5.0e-3, 9),
    (0.0, 1.0, 2),
    (-7.3, 12.9, 4041),
]


@pytest.mark.parametrize("fp32_mode", [False, True])
@pytest.mark.parametrize("lo,hi,n", CASES)
def test_certain_estimate_equals_exact_formula(lo, hi, n, fp32_mode):
    rng = np.random.default_rng(n)
    span = hi - lo
    p = np.concatenate([rng.uniform(lo, hi, 400_000).astype(np.float32), boundary_points(lo, hi, n),
                        np.float32([lo, hi]), np.nextafter(np.float32([lo, hi]), np.float32([np.inf, -np.inf]))])
    inb = (p.astype(np.float64) >= lo) & (p.astype(np.float64) <= hi) if not fp32_mode \
        else (p >= np.float32(lo)) & (p <= np.float32(hi))
    p = p[inb]
    k_fast, certain = device_estimate(p, lo, hi, n)
    k_exact = exact_index(p, lo, hi, n, fp32_mode)
    wrong = certain & (k_fast.astype(np.int64) != k_exact)
    assert not wrong.any(), (p[wrong][:5], k_fast[wrong][:5], k_exact[wrong][:5])
    assert (k_fast[certain] >= 0).all() and (k_fast[certain] <= n - 1).all()
    # the band must stay narrow where precision allows it: the exact (slow) path is the rare one
    uniform_uncertain = 1.0 - certain[:min(len(p), 300_000)].mean()
    cells_of_error = 4.0 * (max(abs(lo), abs(hi)) + span) * 2.0 ** -24 / (span / (n - 1))
    if cells_of_error < 1e-3:
        assert uniform_uncertain < 0.02, uniform_uncertain


def test_degenerate_axes_take_the_exact_path():
    assert fast_index_band(0.0, 0.0, 1) == (0.0, -1.0)
    assert fast_index_band(1.0, 1.0, 5) == (0.0, -1.0)          # zero resolution
    inv, band = fast_index_band(1000.0, 1001.0, 100001)         # 1e-5 cells 1000 units from the origin: band closed
    assert band < 0


def test_random_axes_have_no_counterexample():
    """Random origins (1e-3 .. 1e3 from zero), spans (1e-3 .. 1e2) and sizes (2 .. 4000 cells), both dtype modes."""
    rng = np.random.default_rng(0)
    for _ in range(150):
        lo = float(rng.uniform(-1, 1) * 10 ** rng.uniform(-3, 3))
        hi = lo + float(10 ** rng.uniform(-3, 2))
        n = int(rng.integers(2, 4000))
        p = np.concatenate([rng.uniform(lo, hi, 20_000).astype(np.float32), boundary_points(lo, hi, n, 3)])
        for fp32_mode in (False, True):
            inb = (p >= np.float32(lo)) & (p <= np.float32(hi)) if fp32_mode \
                else (p.astype(np.float64) >= lo) & (p.astype(np.float64) <= hi)
            q = p[inb]
            k_fast, certain = device_estimate(q, lo, hi, n)
            wrong = certain & (k_fast.astype(np.int64) != exact_index(q, lo, hi, n, fp32_mode))
            assert not wrong.any(), (lo, hi, n, fp32_mode, q[wrong][:3])


def test_fp32_range_bounds_equal_the_fp64_comparison():
    """The kernels test fp32 p against host-rounded fp32 bounds instead of converting p to fp64 and comparing with
    the fp64 range (reference: torch compares the fp32 points with fp64 min / max by promotion): identical masks,
    including the fp32 neighbours of both bounds."""
    from pytorch_volumetric_b200.sdf import _fp32_ceil, _fp32_floor
    rng = np.random.default_rng(1)
    for _ in range(300):
        lo = float(rng.uniform(-1, 1) * 10 ** rng.uniform(-4, 3))
        hi = lo + float(10 ** rng.uniform(-4, 2))
        vlo, vhi = np.float32(_fp32_ceil(lo)), np.float32(_fp32_floor(hi))
        p = [np.float32(lo), np.float32(hi)]
        for seed in list(p):
            up, down = seed, seed
            for _ in range(3):
                up = np.nextafter(up, np.float32(np.inf)); down = np.nextafter(down, np.float32(-np.inf))
                p += [up, down]
        p = np.concatenate([np.array(p, dtype=np.float32), rng.uniform(lo - 1, hi + 1, 1000).astype(np.float32)])
        want = (p.astype(np.float64) >= lo) & (p.astype(np.float64) <= hi)
        got = (p >= vlo) & (p <= vhi)
        assert np.array_equal(want, got), (lo, hi)
