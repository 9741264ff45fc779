"""Known-answer tests pinning the CPU oracle (SURVEY.md §8c items i-vii).

The reference has no tests for this path; these are the pins for the kernel restatement
of fortran_modules/particle_subroutines.f90:466-556.
"""
import numpy as np
import pytest

from oracle import oracle as O
from pyhalma_b200 import synth

f32 = np.float32


def be(m, x, y, z, tx, ty, tz, variant="f32seq", ncores=2):
    return O.brute_force_binding_energy(ncores, len(m), f32(m), f32(x), f32(y), f32(z), len(tx),
                                        f32(tx), f32(ty), f32(tz), variant=variant)


def test_two_body_exact():
    # r = sqrt(3^2 + 4^2 + 12^2) = 13 exactly; m / r = 26 / 13 = 2
    out = be([26.0], [3.0], [4.0], [12.0], [0.0], [0.0], [0.0])
    assert out.dtype == np.float32 and out[0] == f32(2.0)
    # symmetric pair, both directions
    out = be([26.0, 39.0], [3.0, 0.0], [4.0, 0.0], [12.0, 0.0], [3.0, 0.0], [4.0, 0.0], [12.0, 0.0])
    assert out[0] == f32(3.0) and out[1] == f32(2.0)


def test_shared_coordinate_excludes_pair():
    # particle_subroutines.f90:499-501: sharing only x drops the pair
    m, x, y, z = [5.0, 7.0, 11.0], [1.0, 1.0, 2.0], [0.0, 3.0, 4.0], [0.0, 4.0, 12.0]
    out = be(m, x, y, z, x, y, z)
    # pair (0,1) shares x -> excluded both ways; (0,2) r=sqrt(1+16+144)=sqrt(161); (1,2) r=sqrt(1+1+64)
    r02 = np.sqrt(f32(161.0)); r12 = np.sqrt(f32(66.0))
    assert out[0] == f32(11.0) / r02
    assert out[1] == f32(11.0) / r12
    assert out[2] == f32(f32(5.0) / r02 + f32(7.0) / r12)   # in-order float32 sum, j ascending


def test_duplicate_particle_no_inf():
    m, x, y, z = [1.0, 1.0, 4.0], [0.5, 0.5, 1.5], [0.25, 0.25, 2.25], [0.125, 0.125, 2.125]
    out = be(m, x, y, z, x, y, z)
    assert np.all(np.isfinite(out))
    assert out[0] == out[1] == f32(4.0) / f32(3.0)          # the duplicate is dropped, r = 3
    assert out[2] == f32(f32(1.0) / f32(3.0) + f32(1.0) / f32(3.0))


def test_lattice_gas_massive_exclusion():
    n = 8
    g = (np.arange(n) + 0.5) * 0.25
    X, Y, Z = (a.ravel() for a in np.meshgrid(g, g, g, indexing="ij"))
    m = np.ones(n ** 3)
    # a pair survives only if it differs in all three coordinates: (n-1)^3 sources per target
    excluded = O.count_excluded(f32(X), f32(Y), f32(Z), f32(X), f32(Y), f32(Z))
    assert excluded == n ** 3 * (n ** 3 - (n - 1) ** 3)
    out = be(m, X, Y, Z, X, Y, Z)
    # brute-force numpy float64 with the same predicate
    dx = X[None, :] - X[:, None]; dy = Y[None, :] - Y[:, None]; dz = Z[None, :] - Z[:, None]
    ok = (dx != 0) & (dy != 0) & (dz != 0)
    with np.errstate(divide="ignore"):
        ref = np.where(ok, 1.0 / np.sqrt(dx * dx + dy * dy + dz * dz), 0.0).sum(axis=1)
    np.testing.assert_allclose(out, ref, rtol=2e-6)


def test_plummer_analytic_potential():
    # statistical check: Phi(r) = M / sqrt(r^2 + a^2) (G = 1 units of the kernel output)
    rng = np.random.default_rng(7)
    p = synth.plummer_stars(10_000, 2e-3, 1e6, rng, centre=(0, 0, 0), interloper_frac=0)
    out = be(p.mass, p.x, p.y, p.z, p.x[:500], p.y[:500], p.z[:500], variant="f64acc")
    r = np.sqrt(p.x[:500] ** 2 + p.y[:500] ** 2 + p.z[:500] ** 2)
    ana = 1e10 / np.sqrt(r * r + 4e-6)
    inner = r < 6e-3
    assert np.median(np.abs(out[inner] / ana[inner] - 1)) < 0.03


def test_omp_equals_serial_bit_for_bit():
    # the OpenMP array reduction is the same in-order float sum
    c = synth.config1(1500, 1500, n_dm=200)
    s, g = c.stars, c.gas
    args = (f32(np.concatenate((g.mass, s.mass))), f32(np.concatenate((g.x, s.x))),
            f32(np.concatenate((g.y, s.y))), f32(np.concatenate((g.z, s.z))))
    ser = O.serial_brute_force_binding_energy(len(args[0]), *args, len(s), f32(s.x), f32(s.y), f32(s.z))
    for nc in (1, 3, 8):
        par = O.brute_force_binding_energy(nc, len(args[0]), *args, len(s), f32(s.x), f32(s.y), f32(s.z))
        assert np.array_equal(ser.view(np.uint32), par.view(np.uint32))


def test_fma_contraction_is_pinned():
    # the compiler's contraction == fma(dz,dz, fma(dx,dx, dy*dy)) written out
    c = synth.config1(2000, 2000, n_dm=100)
    s, g = c.stars, c.gas
    for src in (s, g):
        a = be(src.mass, src.x, src.y, src.z, s.x, s.y, s.z, "f32seq")
        b = be(src.mass, src.x, src.y, src.z, s.x, s.y, s.z, "f32seq_fma")
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_f64acc_close_to_f32seq_small_n():
    c = synth.config1(3000, 3000, n_dm=100)
    s = c.stars
    a = be(s.mass, s.x, s.y, s.z, s.x, s.y, s.z, "f32seq")
    b = be(s.mass, s.x, s.y, s.z, s.x, s.y, s.z, "f64acc")
    assert b.dtype == np.float64
    assert np.max(np.abs(a / b - 1)) < 2e-5     # float32 in-order drift at N=3e3


def test_wrapper_casts_and_empty():
    # halo_gas.py:166-187: float64 in, float32 cast, float32 out; ntest == 0 -> np.array([])
    rng = np.random.default_rng(3)
    m, x, y, z = rng.uniform(1, 2, 50), rng.normal(3, 1e-3, 50), rng.normal(-7, 1e-3, 50), rng.normal(11, 1e-3, 50)
    out = O.brute_force_binding_energy_fortran(m, x, y, z, x[:10], y[:10], z[:10])
    assert out.dtype == np.float32 and out.shape == (10,)
    direct = be(m, x, y, z, x[:10], y[:10], z[:10])
    assert np.array_equal(out, direct)
    empty = O.brute_force_binding_energy_fortran(m, x, y, z, [], [], [])
    assert empty.shape == (0,) and empty.dtype == np.float64
    with pytest.raises(ValueError):
        O.brute_force_binding_energy(1, 60, f32(m), f32(x), f32(y), f32(z), 10, f32(x[:10]), f32(y[:10]), f32(z[:10]))


def test_nan_and_inf_semantics():
    # NaN coordinates compare unequal (contribute NaN) unless another coordinate excludes the pair
    out = be([1.0, 1.0], [np.nan, 5.0], [1.0, 0.0], [1.0, 7.0], [0.0], [0.0], [0.0])
    assert np.isnan(out[0])
    out = be([1.0, 1.0], [np.nan, 5.0], [0.0, 3.0], [1.0, 7.0], [0.0], [0.0], [0.0])   # y equal -> dropped
    assert out[0] == f32(1.0) / np.sqrt(f32(25 + 9 + 49))
    out = be([1.0], [np.inf], [1.0], [1.0], [0.0], [0.0], [0.0])
    assert out[0] == 0.0


def test_energy_step_promotions():
    # float32 chain for the potential term, float64 kinetic term, <= 0 is bound
    be32 = f32([1.0e12, 2.5e11, 0.0])
    v = np.array([100.0, 10.0, 5.0])
    E = O.energy_step(be32, v, v * 0, v * 0, 0.0, 0.0, 0.0, 9.0)
    G32 = f32(O.G_const())
    pe = f32(f32(-be32 * G32) * f32(9.0))
    assert E.dtype == np.float64
    np.testing.assert_array_equal(E, 0.5 * v * v + pe.astype(np.float64))
    assert (E <= 0).tolist() == [True, True, False]
    # E == 0 counts as bound (halo_properties.py:359, halo_gas.py:476)
    E0 = O.energy_step(f32([0.0]), [0.0], [0.0], [0.0], 0.0, 0.0, 0.0, 2.0)
    assert E0[0] == 0.0 and bool((E0 <= 0)[0])


def test_unbind_loop_fixed_point_properties():
    c = synth.config1(1200, 800, n_dm=150)
    s, g, d = c.stars, c.gas, c.dm
    r = O.unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, pre=[g.pos_mass()], post=[d.pos_mass()],
                      kappa=9.0)
    assert r.n_iter >= 2 and r.n_bound_history[-1] == r.n_bound_history[-2]
    assert np.array_equal(np.flatnonzero(r.mask), r.idx) and np.all(np.diff(r.idx) > 0)
    # idempotence: unbinding the bound set again with the same externals removes nothing
    i = r.idx
    r2 = O.unbind_halo(s.x[i], s.y[i], s.z[i], s.vx[i], s.vy[i], s.vz[i], s.mass[i],
                       pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0)
    assert r2.n_iter == 1 and r2.mask.all()
    np.testing.assert_allclose(r2.vb, r.vb, rtol=1e-13)
    # first pass == the reference's one-pass function (halo_properties.py:333-359)
    M = O.total_mass(np.arange(len(s)), s.mass)
    vb = O.CM_velocity(M, np.arange(len(s)), s.vx, s.vy, s.vz, s.mass)
    bound, _, _ = O.escape_velocity_unbinding(
        (g.x, g.y, g.z, g.mass), (s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass), (d.x, d.y, d.z, d.mass), vb, 3.0)
    r1 = O.unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, pre=[g.pos_mass()], post=[d.pos_mass()],
                       kappa=9.0, max_iter=1)
    assert np.array_equal(r1.mask, bound) and r1.n_iter == 1
    # degenerate inputs
    e = O.unbind_halo([], [], [], [], [], [], [])
    assert e.n_iter == 0 and e.mass == 0.0 and e.vb == (0., 0., 0.)
    lone = O.unbind_halo([0.], [0.], [0.], [1.], [0.], [0.], [1e6], vb_fixed=(0., 0., 0.))
    assert not lone.mask[0] and lone.n_iter == 1       # Phi = 0, KE > 0 -> unbound
