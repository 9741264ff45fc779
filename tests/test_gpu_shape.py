"""GPU parity of particle.halo_shape / particle.sigma_projections through
the C-ABI.  The reference sums in float32 in an OpenMP-dependent order, so it is only
defined to ~1e-4; the device sums in float64 and is compared with the oracle's float64
variant at 1e-6 and with the reference-arithmetic fixture at 1e-4 (tolerances below)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from pyhalma_b200 import halo_properties
from pyhalma_b200.particle import particle
from test_oracle_shape import toy_galaxy

pytestmark = pytest.mark.gpu
WIDE_RTOL = 1e-6          # against the float64-accumulating oracle
REF_RTOL = 1e-4           # against the reference's float32 accumulation


def test_halo_shape_known_answer_and_parity():
    a, b, c = 3.0, 2.0, 1.0
    x = np.array([a, -a, 0, 0, 0, 0]); y = np.array([0, 0, b, -b, 0, 0]); z = np.array([0, 0, 0, 0, c, -c])
    e = particle.halo_shape(1, 6, y, z, x, np.ones(6))
    assert e.dtype == np.float32 and e.shape == (3,)
    np.testing.assert_allclose(e, np.array([a, b, c]) / np.sqrt(3.0), rtol=1e-6)
    rng = np.random.default_rng(2)
    for n in (1, 2, 255, 256, 257, 10_000, 1_000_003):
        p = rng.normal(size=(3, n)) * np.array([[3e-3], [2e-3], [1e-3]])
        m = rng.uniform(0.5e6, 2e6, n)
        got = particle.halo_shape(1, n, *p, m)
        want = O.halo_shape(1, n, *p, m, wide=True)
        np.testing.assert_allclose(got, want, rtol=WIDE_RTOL, atol=1e-12)
    with pytest.raises(ValueError):
        particle.halo_shape(1, 5, np.zeros(4), np.zeros(5), np.zeros(5), np.zeros(5))
    # size-independent properties: exact scaling by powers of two, invariance to particle order
    n = 300_000
    p = rng.normal(size=(3, n)) * np.array([[3.0], [2.0], [1.0]])
    m = rng.uniform(1, 2, n)
    e1 = particle.halo_shape(1, n, *p, m)
    np.testing.assert_array_equal(particle.halo_shape(1, n, *(4 * p), 1024 * m), 4 * e1)
    perm = rng.permutation(n)
    np.testing.assert_allclose(particle.halo_shape(1, n, *p[:, perm], m[perm]), e1, rtol=1e-6)
    np.testing.assert_allclose(e1, [3.0, 2.0, 1.0], rtol=0.02)


def test_sigma_projections_parity_and_errors():
    pl, x, y, z, vx, vy, vz, m = toy_galaxy(40_000, 30_000, seed=9)
    c = (3.0, -7.0, 11.0)
    ll = 0.5e-3
    for n_cell in (25, 24, 7, 1):
        grid = (np.arange(n_cell) - n_cell // 2) * ll
        r05 = (3e-3, 2.4e-3, 1.8e-3)
        args = (1, len(pl), grid, n_cell, pl + 1, x, y, z, vx, vy, vz, m, *c, *r05, ll)
        got = particle.sigma_projections(*args)
        assert isinstance(got, tuple) and len(got) == 5 and all(isinstance(v, float) for v in got)
        np.testing.assert_allclose(got, O.sigma_projections(*args, wide=True), rtol=WIDE_RTOL)
        np.testing.assert_allclose(got, O.sigma_projections(*args), rtol=REF_RTOL * 10)
    # grids that are not strictly increasing take the linear minloc scan
    for grid in ((np.arange(25) - 12)[::-1] * ll, np.repeat((np.arange(13) - 6) * 2 * ll, 2)[:25],
                 np.random.default_rng(4).permutation(25) * ll - 12 * ll):
        args = (1, len(pl), grid, 25, pl + 1, x, y, z, vx, vy, vz, m, *c, *r05, ll)
        np.testing.assert_allclose(particle.sigma_projections(*args), O.sigma_projections(*args, wide=True),
                                   rtol=WIDE_RTOL)
    grid = np.array([-1.0, 0.0, 1.0])
    v = np.array([10.0, -10.0])
    xx = np.array([0.1, -0.1]); yy = np.array([0.05, 0.0]); zz = np.array([0.0, 0.1])
    out = particle.sigma_projections(1, 2, grid, 3, [1, 2], xx, yy, zz, v, v, v, [1.0, 1.0], 0, 0, 0, 1, 1, 1, 0.5)
    np.testing.assert_allclose(out[:3], [10.0, 10.0, 10.0], rtol=1e-6)
    assert out[3] == 0.0 and out[4] == 0.0
    out = particle.sigma_projections(1, 0, grid, 3, np.zeros(0, np.int32), xx, yy, zz, v, v, v, [1.0, 1.0],
                                     0, 0, 0, 1, 1, 1, 0.5)
    assert out == (0.0, 0.0, 0.0, 0.0, 0.0)
    with pytest.raises(IndexError):
        particle.sigma_projections(1, 1, grid, 3, [3], xx, yy, zz, v, v, v, [1.0, 1.0], 0, 0, 0, 1, 1, 1, 0.5)
    with pytest.raises(ValueError):
        particle.sigma_projections(1, 2, grid, 3, [1], xx, yy, zz, v, v, v, [1.0, 1.0], 0, 0, 0, 1, 1, 1, 0.5)


def test_shape_sigma_golden_through_the_reference_wrappers(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "shape_sigma.npz")))
    pl, c, r = g["part_list"], g["com"], float(g["rad05"])
    abc = halo_properties.halo_shape_fortran(pl, g["st_x"], g["st_y"], g["st_z"], g["st_mass"], *c, r)
    np.testing.assert_allclose(abc, g["abc"], rtol=REF_RTOL)
    sig = halo_properties.sigma_projections_fortran(
        g["grid"], int(g["n_cell"]), pl, g["st_x"], g["st_y"], g["st_z"], g["st_vx"], g["st_vy"], g["st_vz"],
        *g["vb"], g["st_mass"], *c, r, 0.8 * r, 0.6 * r, float(g["ll"]))
    np.testing.assert_allclose(sig, g["sigma"], rtol=REF_RTOL)
