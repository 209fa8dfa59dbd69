"""CPU: the C oracle (oracle/geom.c) -- brute force against closed forms, BVH evaluator against brute force."""
import numpy as np
import pytest

import workloads
from oracle import _geom


def cube():
    v = np.array([[x, y, z] for x in (-1., 1.) for y in (-1., 1.) for z in (-1., 1.)], dtype=np.float32)
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6],
                  [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], dtype=np.int32)
    return v, f


def test_brute_closest_point_on_cube(oracle_lib):
    v, f = cube()
    soup = _geom.TriangleSoup(v, f)
    rng = np.random.default_rng(0)
    p = rng.uniform(-3, 3, size=(5000, 3)).astype(np.float32)
    c, d2, face = soup.closest_points(p)
    outside = np.abs(p).max(axis=1) > 1
    exp_out = np.linalg.norm(np.maximum(np.abs(p) - 1, 0), axis=1)
    exp_in = 1 - np.abs(p).max(axis=1)
    exp = np.where(outside, exp_out, exp_in)
    assert np.abs(np.sqrt(d2) - exp).max() < 1e-6
    assert (face >= 0).all() and (face < 12).all()
    # closest points lie on the cube surface
    assert np.abs(np.abs(c).max(axis=1) - 1).max() < 1e-6


def test_brute_ray_parity_on_cube(oracle_lib):
    v, f = cube()
    soup = _geom.TriangleSoup(v, f)
    rng = np.random.default_rng(1)
    p = rng.uniform(-2, 2, size=(4000, 3)).astype(np.float32)
    d = (np.array([2.0, 2.0, 2.0]) + 1e-4 * rng.standard_normal((4000, 3))).astype(np.float32)
    cnt = soup.count_intersections(np.concatenate([p, d], axis=1))
    inside = np.abs(p).max(axis=1) < 1
    assert np.array_equal(cnt % 2 == 1, inside)


@pytest.mark.parametrize("name", ["probe", "wrench", "drill"])
def test_bvh_evaluator_equals_brute(name, oracle_lib):
    v, f = workloads.fixture_mesh(name)
    soup = _geom.TriangleSoup(v, f)
    n = 3000 if name != "drill" else 1000
    p = workloads.uniform_points(n, v.min(0) - 0.03, v.max(0) + 0.03, seed=11).numpy()
    c0, d0, f0 = soup.closest_points(p, "brute")
    c1, d1, f1 = soup.closest_points(p, "bvh")
    assert np.array_equal(d0, d1) and np.array_equal(f0, f1) and np.array_equal(c0, c1)
    far = (v.max(0) + 1.0)[None] + 1e-4 * np.random.default_rng(2).standard_normal((n, 3))
    rays = np.concatenate([p, far.astype(np.float32)], axis=1)
    assert np.array_equal(soup.count_intersections(rays, "brute"), soup.count_intersections(rays, "bvh"))


def test_bumpy_sphere_is_closed_and_sized():
    from pytorch_volumetric_b200.meshio import is_closed_manifold
    v, f = workloads.bumpy_sphere(100, 51)
    assert len(f) == 10000 and is_closed_manifold(f)
    v, f = workloads.bumpy_sphere(250, 101)
    assert len(f) == 50000 and is_closed_manifold(f)
    assert not is_closed_manifold(workloads.fixture_mesh("wrench")[1])
    assert is_closed_manifold(workloads.fixture_mesh("drill")[1])
