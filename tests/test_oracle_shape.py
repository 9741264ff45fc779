"""Known-answer tests of the oracle's halo_shape / sigma_projections restatement
(particle_subroutines.f90:12-461) and of the golden fixture written by the
reference's own wrappers (halo_properties.py:781-812, :852-866)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

f32 = np.float32


def numpy_sigma(grid, part_list0, x, y, z, vx, vy, vz, m, c, r05, ll):
    """Independent float64 numpy restatement of particle_subroutines.f90:217-461."""
    grid = np.asarray(grid, np.float32)
    n = len(grid)
    q = np.asarray(part_list0)
    d = [f32(a)[q] - f32(cc) for a, cc in zip((x, y, z), c)]
    cell = [np.argmin(np.abs(grid[None, :] - dd[:, None]), axis=1) for dd in d]      # first minimum
    v = [f32(a)[q].astype(np.float64) for a in (vx, vy, vz)]
    mm = f32(m)[q].astype(np.float64)
    pairs = [(1, 2), (0, 2), (0, 1)]                       # (first, second) subscripts of the x, y, z maps
    s05, vs, lam = [], [], []
    for a, (i, j) in enumerate(pairs):
        k = cell[i] + n * cell[j]
        sd = np.bincount(k, mm, n * n)
        cnt = np.bincount(k, None, n * n)
        vcm = np.bincount(k, v[a] * mm, n * n)
        vcm = np.where(sd != 0, vcm / np.where(sd != 0, sd, 1), vcm)
        sig = np.bincount(k, (v[a] - vcm[k]) ** 2, n * n)
        sig = np.where(cnt != 0, np.sqrt(sig / np.where(cnt != 0, cnt, 1)), sig)
        dist = np.sqrt(d[i] * d[i] + d[j] * d[j])            # float32 like the Fortran
        inside = dist < f32(r05[a])
        s05.append(sig[k][inside].mean() if inside.any() else 0.0)
        gi, gj = np.meshgrid(grid, grid, indexing="ij")
        rbin = np.sqrt(gi * gi + gj * gj).ravel(order="F")   # index first + n * second
        sel = rbin < f32(r05[a]) + f32(2) * f32(ll)
        rb = rbin.astype(np.float64)
        sv = (vcm ** 2 * sd)[sel].sum()
        ss = (sig ** 2 * sd)[sel].sum()
        up = (sd * rb * np.abs(vcm))[sel].sum()
        dn = (sd * rb * np.sqrt(vcm ** 2 + sig ** 2))[sel].sum()
        vs.append(np.sqrt(sv / ss) if ss > 0 else 0.0)
        lam.append(up / dn if dn > 0 else 0.0)
    return np.array(s05 + [sum(vs) / 3, sum(lam) / 3])


def toy_galaxy(n_glob=3000, n=2000, seed=5):
    rng = np.random.default_rng(seed)
    pl = np.sort(rng.choice(n_glob, n, replace=False))
    x = 3.0 + rng.normal(0, 3e-3, n_glob)
    y = -7.0 + rng.normal(0, 2e-3, n_glob)
    z = 11.0 + rng.normal(0, 1e-3, n_glob)
    m = rng.uniform(0.5e6, 2e6, n_glob)
    vx = rng.normal(0, 60, n_glob) - 2.0e4 * (y + 7.0)
    vy = rng.normal(0, 60, n_glob) + 2.0e4 * (x - 3.0)
    vz = rng.normal(0, 40, n_glob)
    return pl, x, y, z, vx, vy, vz, m


def test_jacobi_eigenvalues_against_numpy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.normal(size=(3, 3)) * 10.0 ** rng.integers(-6, 6)
        s = f32(a @ a.T)
        got = np.sort(O.diagonalise(s).astype(np.float64))
        want = np.linalg.eigvalsh(s.astype(np.float64))
        # convergence criterion is sum|off-diagonal| < 1e-4 * sum|elements| (:57): the eigenvalue
        # error is second order in the remaining off-diagonal part
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6 * np.abs(want).max())
    d = O.diagonalise(np.diag([3.0, 1.0, 2.0]))
    assert np.array_equal(d, f32([3.0, 1.0, 2.0]))           # diagonal input: no rotation, order kept


def test_halo_shape_known_answer_and_order():
    a, b, c = 3.0, 2.0, 1.0
    x = np.array([a, -a, 0, 0, 0, 0])
    y = np.array([0, 0, b, -b, 0, 0])
    z = np.array([0, 0, 0, 0, c, -c])
    m = np.ones(6)
    for perm in ((x, y, z), (z, x, y), (y, z, x)):
        e = O.halo_shape(1, 6, *perm, m)
        assert e.dtype == np.float32
        np.testing.assert_allclose(e, np.array([a, b, c]) / np.sqrt(3.0), rtol=1e-6)    # largest first (:129-157)
    # mass scaling leaves the axes unchanged, a length scale multiplies them
    rng = np.random.default_rng(1)
    p = rng.normal(size=(3, 500)) * np.array([[3.0], [2.0], [1.0]])
    mm = rng.uniform(1, 2, 500)
    e1 = O.halo_shape(1, 500, *p, mm, wide=True)
    e2 = O.halo_shape(1, 500, *(4.0 * p), 1024.0 * mm, wide=True)
    np.testing.assert_array_equal(e2, 4.0 * e1)              # powers of two: exact
    w = np.linalg.eigvalsh(np.einsum("n,in,jn->ij", f32(mm).astype(float), f32(p).astype(float),
                                     f32(p).astype(float)) / f32(mm).astype(float).sum())
    np.testing.assert_allclose(np.sort(e1.astype(float) ** 2), w, rtol=1e-5)
    np.testing.assert_allclose(O.halo_shape(1, 500, *p, mm), e1, rtol=1e-4)      # float32 sums drift


def test_halo_shape_shape_errors():
    with pytest.raises(ValueError):
        O.halo_shape(1, 5, np.zeros(4), np.zeros(5), np.zeros(5), np.zeros(5))


def test_sigma_projections_against_numpy_restatement():
    pl, x, y, z, vx, vy, vz, m = toy_galaxy()
    c = (3.0, -7.0, 11.0)
    ll = 0.5e-3
    for n_cell in (25, 24, 7):                                # odd, even and coarse grids
        grid = (np.arange(n_cell) - n_cell // 2) * ll
        r05 = (3e-3, 2.4e-3, 1.8e-3)
        got = O.sigma_projections(1, len(pl), grid, n_cell, pl + 1, x, y, z, vx, vy, vz, m, *c, *r05, ll, wide=True)
        want = numpy_sigma(grid, pl, x, y, z, vx, vy, vz, m, c, r05, ll)
        np.testing.assert_allclose(got, want, rtol=2e-6)
        seq = O.sigma_projections(1, len(pl), grid, n_cell, pl + 1, x, y, z, vx, vy, vz, m, *c, *r05, ll)
        np.testing.assert_allclose(seq, want, rtol=2e-4)       # the reference's float32 sums


def test_sigma_projections_known_answers():
    grid = np.array([-1.0, 0.0, 1.0])
    # two particles in the central cell with opposite line-of-sight velocities: mean 0, sigma = v
    x = np.array([0.1, -0.1]); y = np.array([0.05, 0.0]); z = np.array([0.0, 0.1])
    v = np.array([10.0, -10.0])
    out = O.sigma_projections(1, 2, grid, 3, [1, 2], x, y, z, v, v, v, [1.0, 1.0], 0, 0, 0, 1, 1, 1, 0.5)
    np.testing.assert_allclose(out[:3], [10.0, 10.0, 10.0], rtol=1e-6)
    assert out[3] == 0.0 and out[4] == 0.0                      # no ordered motion
    # one particle: dispersion 0 -> sumSigma = 0 -> V_sigma stays 0 (:397), lambda = 1 off-centre
    out = O.sigma_projections(1, 1, grid, 3, [1], [1.0], [1.0], [1.0], [5.0], [5.0], [5.0], [2.0],
                              0, 0, 0, 5, 5, 5, 0.5)
    assert out[:4] == (0.0, 0.0, 0.0, 0.0) and out[4] == pytest.approx(1.0)
    # nobody inside R05: the three means stay 0 (:356-366)
    out = O.sigma_projections(1, 2, grid, 3, [1, 2], x, y, z, v, v, v, [1.0, 1.0], 0, 0, 0, 1e-3, 1e-3, 1e-3, 0.5)
    assert out[:3] == (0.0, 0.0, 0.0)
    # ties go to the first grid point (minloc): d = 0.5 sits between grid[1] and grid[2]
    out_a = O.sigma_projections(1, 2, grid, 3, [1, 2], [0.5, 0.0], [0.0, 0.0], [0.0, 0.0], v, v, v, [1.0, 1.0],
                                0, 0, 0, 2, 2, 2, 0.5)
    np.testing.assert_allclose(out_a[:3], [10.0, 10.0, 10.0], rtol=1e-6)   # both in the central cell
    # empty list
    out = O.sigma_projections(1, 0, grid, 3, np.zeros(0, np.int32), x, y, z, v, v, v, [1.0, 1.0],
                              0, 0, 0, 1, 1, 1, 0.5)
    assert out == (0.0, 0.0, 0.0, 0.0, 0.0)
    with pytest.raises(IndexError):
        O.sigma_projections(1, 1, grid, 3, [3], x, y, z, v, v, v, [1.0, 1.0], 0, 0, 0, 1, 1, 1, 0.5)
    with pytest.raises(IndexError):
        O.sigma_projections(1, 1, grid, 3, [0], x, y, z, v, v, v, [1.0, 1.0], 0, 0, 0, 1, 1, 1, 0.5)


def test_shape_sigma_golden_from_reference_wrappers(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "shape_sigma.npz")))
    pl = g["part_list"]
    c = g["com"]
    abc = O.halo_shape_fortran(pl, g["st_x"], g["st_y"], g["st_z"], g["st_mass"], *c, float(g["rad05"]))
    np.testing.assert_array_equal(f32(abc), g["abc"])
    assert abc[0] >= abc[1] >= abc[2] > 0
    np.testing.assert_allclose(abc, [3e-3, 2e-3, 1e-3], rtol=0.05)           # the generator's axes
    r = float(g["rad05"])
    args = (g["grid"], int(g["n_cell"]), pl, g["st_x"], g["st_y"], g["st_z"], g["st_vx"], g["st_vy"], g["st_vz"],
            *g["vb"], g["st_mass"], *c, r, 0.8 * r, 0.6 * r, float(g["ll"]))
    sig = O.sigma_projections_fortran(*args)
    np.testing.assert_array_equal(f32(sig), g["sigma"])
    wide = O.sigma_projections_fortran(*args, wide=True)
    np.testing.assert_allclose(wide, g["sigma"], rtol=1e-4)
    assert 0 < sig[4] < 1 and sig[3] > 0                                    # a rotating system
