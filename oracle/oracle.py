"""CPU oracle for the pyHALMA unbinding hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (pyhalma_b200/) never does.

PARITY STATUS.  The reference has no tests or golden vectors for this path and its
Fortran kernel cannot be compiled in this image (no gfortran), so the kernel
arithmetic is "parity unpinned" by the reference itself.  The pins
are: tests/test_oracle_kat.py (known answers) and tests/golden/*.npz, which were
written by tests/golden/make_golden.py running the reference's OWN Python drivers
with `fortran_modules.particle` bound to this oracle.

All citations are relative to /root/reference/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhalma_oracle.so")
_lib = None

_F32P = ctypes.POINTER(ctypes.c_float)
_F64P = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile oracle/halma_oracle.c with the reference's flags (oracle/Makefile)."""
    src = os.path.join(_HERE, "halma_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libhalma_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64 = ctypes.c_int64
        for name, outp in (("oracle_potential_f32seq", _F32P),
                           ("oracle_potential_f32seq_fma", _F32P),
                           ("oracle_potential_f64acc", _F64P)):
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_int, i64, _F32P, _F32P, _F32P, _F32P, i64, _F32P, _F32P,
                           _F32P, outp]
        L.oracle_potential_f32seq_serial.restype = None
        L.oracle_potential_f32seq_serial.argtypes = [i64, _F32P, _F32P, _F32P, _F32P, i64,
                                                     _F32P, _F32P, _F32P, _F32P]
        L.oracle_count_excluded.restype = i64
        L.oracle_count_excluded.argtypes = [i64, _F32P, _F32P, _F32P, i64, _F32P, _F32P, _F32P]
        I32P = ctypes.POINTER(ctypes.c_int32)
        L.oracle_halo_shape.restype = None
        L.oracle_halo_shape.argtypes = [i64, _F32P, _F32P, _F32P, _F32P, ctypes.c_int, _F32P]
        L.oracle_diagonalise.restype = None
        L.oracle_diagonalise.argtypes = [_F32P, _F32P]
        L.oracle_sigma_projections.restype = None
        L.oracle_sigma_projections.argtypes = ([i64, _F32P, ctypes.c_int, I32P] + [_F32P] * 7 + [ctypes.c_float] * 7
                                               + [ctypes.c_int, _F32P])
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_wtime.restype = ctypes.c_double
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _f32c(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError("expected a 1-D array")
    return a


def _p32(a: np.ndarray):
    return a.ctypes.data_as(_F32P)


# --------------------------------------------------------------------------------------
# f2py-level entry points: particle.particle.brute_force_binding_energy
# (fortran_modules/particle_subroutines.f90:466-514, serial twin :517-556)
# --------------------------------------------------------------------------------------
def _check_f2py_shapes(ntotal, arrays_total, ntest, arrays_test):
    # f2py raises when a depend(ntotal)/depend(ntest) array is shorter than the size
    # argument (particle_subroutines.f90:485-488).
    for a in arrays_total:
        if a.shape[0] != int(ntotal):
            raise ValueError("0-th dimension must be fixed to %d but got %d" % (ntotal, a.shape[0]))
    for a in arrays_test:
        if a.shape[0] != int(ntest):
            raise ValueError("0-th dimension must be fixed to %d but got %d" % (ntest, a.shape[0]))


def brute_force_binding_energy(ncores, ntotal, total_mass, total_x, total_y, total_z,
                               ntest, test_x, test_y, test_z, *, variant: str = "f32seq"):
    """Positional twin of the f2py function (particle_subroutines.f90:466-469).

    variant: "f32seq" (bit-faithful), "f32seq_fma" (contraction spelled out) or
    "f64acc" (same terms, double accumulator; returns float64).
    """
    tm, tx, ty, tz = map(_f32c, (total_mass, total_x, total_y, total_z))
    sx, sy, sz = map(_f32c, (test_x, test_y, test_z))
    _check_f2py_shapes(ntotal, (tm, tx, ty, tz), ntest, (sx, sy, sz))
    L = lib()
    if variant == "f64acc":
        out = np.zeros(int(ntest), dtype=np.float64)
        L.oracle_potential_f64acc(int(ncores), int(ntotal), _p32(tm), _p32(tx), _p32(ty), _p32(tz),
                                  int(ntest), _p32(sx), _p32(sy), _p32(sz),
                                  out.ctypes.data_as(_F64P))
        return out
    out = np.zeros(int(ntest), dtype=np.float32)
    fn = {"f32seq": L.oracle_potential_f32seq, "f32seq_fma": L.oracle_potential_f32seq_fma}[variant]
    fn(int(ncores), int(ntotal), _p32(tm), _p32(tx), _p32(ty), _p32(tz), int(ntest), _p32(sx),
       _p32(sy), _p32(sz), _p32(out))
    return out


def serial_brute_force_binding_energy(ntotal, total_mass, total_x, total_y, total_z,
                                      ntest, test_x, test_y, test_z):
    """particle_subroutines.f90:517-556."""
    tm, tx, ty, tz = map(_f32c, (total_mass, total_x, total_y, total_z))
    sx, sy, sz = map(_f32c, (test_x, test_y, test_z))
    _check_f2py_shapes(ntotal, (tm, tx, ty, tz), ntest, (sx, sy, sz))
    out = np.zeros(int(ntest), dtype=np.float32)
    lib().oracle_potential_f32seq_serial(int(ntotal), _p32(tm), _p32(tx), _p32(ty), _p32(tz),
                                         int(ntest), _p32(sx), _p32(sy), _p32(sz), _p32(out))
    return out


# --------------------------------------------------------------------------------------
# §8f-4: particle.halo_shape (particle_subroutines.f90:160-214) and
# particle.sigma_projections (:217-461), f2py signatures.
# --------------------------------------------------------------------------------------
def diagonalise(matrix3x3):
    m = np.ascontiguousarray(matrix3x3, dtype=np.float32).reshape(3, 3)
    out = np.zeros(3, np.float32)
    lib().oracle_diagonalise(_p32(m.reshape(-1)), _p32(out))
    return out


def halo_shape(ncore, npart, x, y, z, mass, *, wide: bool = False):
    x, y, z, mass = map(_f32c, (x, y, z, mass))
    _check_f2py_shapes(npart, (x, y, z, mass), npart, ())
    out = np.zeros(3, np.float32)
    lib().oracle_halo_shape(int(npart), _p32(x), _p32(y), _p32(z), _p32(mass), int(wide), _p32(out))
    return out


def sigma_projections(ncore, npart, grid, n_cell, part_list, st_x, st_y, st_z, st_vx, st_vy, st_vz, st_mass,
                      cx, cy, cz, R05x, R05y, R05z, ll, *, wide: bool = False):
    """part_list is 1-based, exactly what the reference hands to Fortran (halo_properties.py:787)."""
    grid = _f32c(grid)
    pl = np.ascontiguousarray(part_list, dtype=np.int32) - 1
    arrs = [_f32c(a) for a in (st_x, st_y, st_z, st_vx, st_vy, st_vz, st_mass)]
    if len(pl) != int(npart) or len(grid) != int(n_cell):
        raise ValueError("0-th dimension must be fixed")
    if len(pl) and (pl.min() < 0 or pl.max() >= len(arrs[0])):
        raise IndexError("part_list out of range")
    out = np.zeros(5, np.float32)
    lib().oracle_sigma_projections(int(npart), _p32(grid), int(n_cell), pl.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                   *[_p32(a) for a in arrs], float(cx), float(cy), float(cz), float(R05x),
                                   float(R05y), float(R05z), float(ll), int(wide), _p32(out))
    return tuple(out)


def halo_shape_fortran(part_list, st_x, st_y, st_z, st_mass, cx, cy, cz, RAD05, *, wide=False):
    """python_scripts/halo_properties.py:852-866."""
    x = np.float32(st_x[part_list] - cx)
    y = np.float32(st_y[part_list] - cy)
    z = np.float32(st_z[part_list] - cz)
    mass = np.float32(st_mass[part_list])
    e = halo_shape(np.int32(max_threads()), np.int32(len(part_list)), x, y, z, mass, wide=wide)
    return e[0], e[1], e[2]


def sigma_projections_fortran(grid, n_cell, part_list, st_x, st_y, st_z, st_vx, st_vy, st_vz, vx, vy, vz, st_mass,
                              cx, cy, cz, R05x, R05y, R05z, ll, *, wide=False):
    """python_scripts/halo_properties.py:781-812: float32 casts, velocities relative to the
    bulk velocity, part_list + 1 for Fortran."""
    return sigma_projections(np.int32(max_threads()), np.int32(len(part_list)), np.float32(grid), np.int32(n_cell),
                             np.int32(1 + part_list), np.float32(st_x), np.float32(st_y), np.float32(st_z),
                             np.float32(st_vx - vx), np.float32(st_vy - vy), np.float32(st_vz - vz),
                             np.float32(st_mass), np.float32(cx), np.float32(cy), np.float32(cz), np.float32(R05x),
                             np.float32(R05y), np.float32(R05z), np.float32(ll), wide=wide)


class _ParticleNamespace:
    """Stands in for the f2py module object: `particle.particle.<routine>`
    (python_scripts/halo_gas.py:6,182)."""
    brute_force_binding_energy = staticmethod(brute_force_binding_energy)
    serial_brute_force_binding_energy = staticmethod(serial_brute_force_binding_energy)
    halo_shape = staticmethod(halo_shape)
    sigma_projections = staticmethod(sigma_projections)


class particle_module:  # noqa: N801  (mirrors the f2py module name)
    particle = _ParticleNamespace


def count_excluded(total_x, total_y, total_z, test_x, test_y, test_z) -> int:
    tx, ty, tz = map(_f32c, (total_x, total_y, total_z))
    sx, sy, sz = map(_f32c, (test_x, test_y, test_z))
    return int(lib().oracle_count_excluded(len(tx), _p32(tx), _p32(ty), _p32(tz), len(sx),
                                           _p32(sx), _p32(sy), _p32(sz)))


# --------------------------------------------------------------------------------------
# a3: python_scripts/halo_gas.py:164-187 (and serial twin :191-215)
# --------------------------------------------------------------------------------------
def brute_force_binding_energy_fortran(total_mass, total_x, total_y, total_z, test_x, test_y,
                                       test_z, *, variant: str = "f32seq", ncores: Optional[int] = None):
    ntotal = np.int32(len(total_mass))          # :167
    ntest = np.int32(len(test_x))               # :168
    if ntest == 0:                              # :169-170
        return np.array([])
    casts = [np.asarray(a).astype(np.float32) for a in
             (total_mass, total_x, total_y, total_z, test_x, test_y, test_z)]   # :172-178
    nc = np.int32(max_threads() if ncores is None else ncores)                  # :181
    return brute_force_binding_energy(nc, ntotal, *casts[:4], ntest, *casts[4:], variant=variant)


# --------------------------------------------------------------------------------------
# a6: python_scripts/halo_properties.py:16-60  (float64 mass-weighted sums)
# --------------------------------------------------------------------------------------
def total_mass(part_list, st_mass) -> float:
    return float(np.sum(np.asarray(st_mass, dtype=np.float64)[part_list]))        # :17-24


def center_of_mass(part_list, st_x, st_y, st_z, st_mass):
    m = np.asarray(st_mass, dtype=np.float64)[part_list]                          # :26-43
    M = float(np.sum(m))
    if M > 0:
        return (float(np.sum(m * np.asarray(st_x, dtype=np.float64)[part_list]) / M),
                float(np.sum(m * np.asarray(st_y, dtype=np.float64)[part_list]) / M),
                float(np.sum(m * np.asarray(st_z, dtype=np.float64)[part_list]) / M), M)
    return 0., 0., 0., 0.


def CM_velocity(M, part_list, st_vx, st_vy, st_vz, st_mass):  # noqa: N802
    m = np.asarray(st_mass, dtype=np.float64)[part_list]                          # :45-60
    if M > 0.:
        return (float(np.sum(m * np.asarray(st_vx, dtype=np.float64)[part_list]) / M),
                float(np.sum(m * np.asarray(st_vy, dtype=np.float64)[part_list]) / M),
                float(np.sum(m * np.asarray(st_vz, dtype=np.float64)[part_list]) / M))
    return 0., 0., 0.


# --------------------------------------------------------------------------------------
# a7 constants (halo_gas.py:459-465, halo_properties.py:345-351)
# --------------------------------------------------------------------------------------
def G_const() -> float:  # noqa: N802
    g = 4.3 * 1e-3     # (km/s)^2 pc/Msun
    g *= 1e-6          # (km/s)^2 Mpc/Msun
    return g


COLD_T = 5 * 1e4       # halo_gas.py:479-480


def energy_step(be32: np.ndarray, vx, vy, vz, vbx, vby, vbz, kappa) -> np.ndarray:
    """Total specific energy with the reference's dtype promotions.

    Stars  halo_properties.py:342-357 (kappa = factor_v**2):
        be = -be ; be *= G ; be *= factor_v**2      (all float32, two roundings)
        ke = 0.5*((vx-vbx)**2 + (vy-vby)**2 + (vz-vbz)**2)   (float64, no FMA)
        E  = ke + be                                 (float32 promoted to float64)
    Gas    halo_gas.py:456-471 (kappa = 2.): the same chain with `be *= 2.`.
    """
    be = np.array(be32, dtype=np.float32, copy=True)
    be = -be
    be *= G_const()
    be *= kappa
    assert be.dtype == np.float32
    ke = 0.5 * ((np.asarray(vx, np.float64) - vbx) ** 2 + (np.asarray(vy, np.float64) - vby) ** 2
                + (np.asarray(vz, np.float64) - vbz) ** 2)
    return ke + be


# --------------------------------------------------------------------------------------
# a4: python_scripts/halo_properties.py:282-361, from :333 on (inputs already gathered:
# the AMR/DM gather at :289-326 is SURVEY §2 row 7, out of scope).
# --------------------------------------------------------------------------------------
def escape_velocity_unbinding(gas, stars, dm, vb, factor_v, *, variant: str = "f32seq"):
    """gas/dm = (x, y, z, mass); stars = (x, y, z, vx, vy, vz, mass); vb = (vx, vy, vz).

    Returns (bound, be32, E).  Source order is gas, stars, DM (:333-336).
    """
    gx, gy, gz, gm = gas
    sx, sy, sz, svx, svy, svz, sm = stars
    dx, dy, dz, dmass = dm
    tot_x = np.concatenate((gx, sx, dx))
    tot_y = np.concatenate((gy, sy, dy))
    tot_z = np.concatenate((gz, sz, dz))
    tot_m = np.concatenate((gm, sm, dmass))
    be = brute_force_binding_energy_fortran(tot_m, tot_x, tot_y, tot_z, sx, sy, sz, variant=variant)
    if len(sx) == 0:
        return np.zeros(0, bool), np.zeros(0, np.float32), np.zeros(0)
    be32 = np.float32(be)        # f64acc variant: one rounding to the f2py output dtype
    E = energy_step(be32, svx, svy, svz, vb[0], vb[1], vb[2], factor_v ** 2)
    return E <= 0., be32, E      # :359


# --------------------------------------------------------------------------------------
# a5: python_scripts/halo_gas.py:285-492.  Class order of the float32 `+=` is gas-self,
# DM (heavy then light when num_dm_species > 1), stars.  Sampling (:306-321 etc.) uses
# the global numpy RNG exactly like the reference, so a seeded run reproduces it.
# --------------------------------------------------------------------------------------
@dataclass
class RPSResult:
    total_gas_mass: float
    frac_cold_gas_mass: float
    unbound_cold_gas_mass: float
    unbound_hot_gas_mass: float
    bound: np.ndarray = field(repr=False, default=None)
    be32: np.ndarray = field(repr=False, default=None)
    energy: np.ndarray = field(repr=False, default=None)

    def as_tuple(self):
        return (self.total_gas_mass, self.frac_cold_gas_mass, self.unbound_cold_gas_mass,
                self.unbound_hot_gas_mass)


def _class_sum(be_acc, src_m, src_x, src_y, src_z, tx, ty, tz, lim, variant):
    """One source class of RPS / most_bound_particle: exact below `lim`, else the
    reference's with-replacement subsample rescaled by n/nsample (halo_gas.py:306-328)."""
    n = len(src_x)
    if n > lim:
        nsample = np.max([lim, int(0.01 * n)])
        sample = np.random.choice(np.arange(n), nsample, replace=True)
        be = brute_force_binding_energy_fortran(src_m[sample], src_x[sample], src_y[sample],
                                                src_z[sample], tx, ty, tz, variant=variant)
        be = np.float32(be) * n / nsample
        be_acc += be
    elif n > 0:
        be = brute_force_binding_energy_fortran(src_m, src_x, src_y, src_z, tx, ty, tz,
                                                variant=variant)
        be_acc += np.float32(be)


def split_dm_species(dm_x, dm_y, dm_z, dm_mass, mass_dm_part):
    """halo_gas.py:340-366: heavy ("mandatory", l<=1) vs light DM particles."""
    heavy = dm_mass >= 0.9 * (mass_dm_part / 8)
    light = np.logical_not(heavy)
    return ((dm_x[heavy], dm_y[heavy], dm_z[heavy], dm_mass[heavy]),
            (dm_x[light], dm_y[light], dm_z[light], dm_mass[light]))


def rps_potential(gas_x, gas_y, gas_z, gas_mass, dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z,
                  st_mass, BRUTE_FORCE_LIM, mass_dm_part, num_dm_species, *, variant="f32seq"):
    """float32 sum over the source classes, halo_gas.py:299-450."""
    ngas = len(gas_x)
    be = np.zeros((ngas,), dtype=np.float32)                                        # :301
    _class_sum(be, gas_mass, gas_x, gas_y, gas_z, gas_x, gas_y, gas_z, BRUTE_FORCE_LIM, variant)
    if num_dm_species > 1:                                                          # :337
        heavy, light = split_dm_species(dm_x, dm_y, dm_z, dm_mass, mass_dm_part)
        if len(heavy[0]) > 0:                                                       # :353 (never sampled)
            be += np.float32(brute_force_binding_energy_fortran(heavy[3], heavy[0], heavy[1],
                                                                heavy[2], gas_x, gas_y, gas_z,
                                                                variant=variant))
        _class_sum(be, light[3], light[0], light[1], light[2], gas_x, gas_y, gas_z,
                   BRUTE_FORCE_LIM, variant)
    else:                                                                           # :396
        _class_sum(be, dm_mass, dm_x, dm_y, dm_z, gas_x, gas_y, gas_z, BRUTE_FORCE_LIM, variant)
    _class_sum(be, st_mass, st_x, st_y, st_z, gas_x, gas_y, gas_z, BRUTE_FORCE_LIM, variant)
    return be


def rps_masses(gas_mass, gas_temp, total_energy):
    """halo_gas.py:475-492."""
    unbound = total_energy > 0.
    bound = total_energy <= 0.
    cold = gas_temp < COLD_T
    hot = gas_temp >= COLD_T
    total_gas_mass = np.sum(gas_mass)
    cold_bound = np.sum(gas_mass[cold * bound])
    frac = cold_bound / total_gas_mass if total_gas_mass != 0. else 0.
    return (total_gas_mass, frac, np.sum(gas_mass[unbound * cold]), np.sum(gas_mass[unbound * hot]),
            bound)


def RPS(gas_x, gas_y, gas_z, gas_vx, gas_vy, gas_vz, gas_mass, gas_temp,  # noqa: N802
        dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z, st_mass, vx, vy, vz, BRUTE_FORCE_LIM,
        mass_dm_part, num_dm_species, *, variant: str = "f32seq") -> RPSResult:
    be = rps_potential(gas_x, gas_y, gas_z, gas_mass, dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z,
                       st_mass, BRUTE_FORCE_LIM, mass_dm_part, num_dm_species, variant=variant)
    E = energy_step(be, gas_vx, gas_vy, gas_vz, vx, vy, vz, 2.)                     # :456-471
    tot, frac, ucold, uhot, bound = rps_masses(gas_mass, gas_temp, E)
    return RPSResult(tot, frac, ucold, uhot, bound, be, E)


# --------------------------------------------------------------------------------------
# §8f-1: python_scripts/halo_gas.py:498-634  most_bound_particle (targets = stars;
# class order gas, DM heavy, DM light, stars; the DM split is unconditional here, :548).
# --------------------------------------------------------------------------------------
def most_bound_potential(gas_x, gas_y, gas_z, gas_mass, dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z,
                         st_mass, BRUTE_FORCE_LIM, mass_dm_part, *, variant="f32seq"):
    nst = len(st_x)
    be = np.zeros((nst,), dtype=np.float32)
    _class_sum(be, gas_mass, gas_x, gas_y, gas_z, st_x, st_y, st_z, BRUTE_FORCE_LIM, variant)
    heavy, light = split_dm_species(dm_x, dm_y, dm_z, dm_mass, mass_dm_part)
    if len(heavy[0]) > 0:
        be += np.float32(brute_force_binding_energy_fortran(heavy[3], heavy[0], heavy[1], heavy[2],
                                                            st_x, st_y, st_z, variant=variant))
    _class_sum(be, light[3], light[0], light[1], light[2], st_x, st_y, st_z, BRUTE_FORCE_LIM, variant)
    _class_sum(be, st_mass, st_x, st_y, st_z, st_x, st_y, st_z, BRUTE_FORCE_LIM, variant)
    return be


def most_bound_particle(gas_x, gas_y, gas_z, gas_mass, dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z,
                        st_mass, st_oripa, BRUTE_FORCE_LIM, mass_dm_part, *, variant="f32seq"):
    be = most_bound_potential(gas_x, gas_y, gas_z, gas_mass, dm_x, dm_y, dm_z, dm_mass, st_x, st_y,
                              st_z, st_mass, BRUTE_FORCE_LIM, mass_dm_part, variant=variant)
    k = np.argmin(-be)                                                              # :627-632
    return st_x[k], st_y[k], st_z[k], st_oripa[k]


# --------------------------------------------------------------------------------------
# the fixed-point iteration of the reference's one-pass functions.
# Iteration 1 is the reference (a4 / a5); iterations >= 2 re-apply it with
# part_list <- part_list[bound] and, when not fixed, vb <- CM velocity of the bound set.
# --------------------------------------------------------------------------------------
@dataclass
class UnbindResult:
    mask: np.ndarray          # bool[N] over the original members
    idx: np.ndarray           # ascending member indices (== flatnonzero(mask))
    be32: np.ndarray          # float32[N]: sum m/r at the last pass the particle took part in
    energy: np.ndarray        # float64[N]: total energy at that pass
    n_iter: int
    mass: float
    com: tuple
    vb: tuple
    n_bound_history: list
    pairs: int                # interactions evaluated (targets x sources, summed over passes)


SourceGroup = Sequence[np.ndarray]     # (mass, x, y, z)


def unbind_halo(x, y, z, vx, vy, vz, mass, *, pre: Sequence[SourceGroup] = (),
                post: Sequence[SourceGroup] = (), split_classes: bool = False, kappa: float = 9.0,
                vb_fixed=None, max_iter: int = 64, variant: str = "f32seq") -> UnbindResult:
    """Iterative unbinding of one halo.

    Members (x..mass, float64 like the reference's particle arrays) are both targets
    and sources.  External source groups are fixed.  Two layouts cover the reference:

    * stellar (a4): split_classes=False, pre=[gas], post=[dm]; one float32 in-order sum
      over concat(pre..., members, post...) (halo_properties.py:333-339); kappa=factor_v**2;
      vb recomputed per pass with CM_velocity unless vb_fixed is given.
    * gas (a5): split_classes=True, pre=[], post=[dm(, dm_light), stars]; every group is
      summed separately in float32 and added in float32 in order (halo_gas.py:301-450);
      kappa=2; vb_fixed = the stellar halo's bulk velocity (halo_gas.py:468).
    """
    cols = [np.asarray(a, dtype=np.float64) for a in (x, y, z, vx, vy, vz, mass)]
    x, y, z, vx, vy, vz, mass = cols
    N = len(x)
    idx = np.arange(N)
    be_out = np.zeros(N, np.float32)
    e_out = np.zeros(N, np.float64)
    hist = [N]
    it = 0
    pairs = 0
    n_ext = sum(len(g[0]) for g in pre) + sum(len(g[0]) for g in post)
    while len(idx) > 0 and it < max_iter:
        mx, my, mz, mm = x[idx], y[idx], z[idx], mass[idx]
        if vb_fixed is None:
            M = total_mass(idx, mass)
            vb = CM_velocity(M, idx, vx, vy, vz, mass)
        else:
            vb = tuple(float(v) for v in vb_fixed)
        if split_classes:
            be = np.zeros(len(idx), np.float32)
            for g in list(pre) + [(mm, mx, my, mz)] + list(post):
                if len(g[0]) > 0:
                    be += np.float32(brute_force_binding_energy_fortran(g[0], g[1], g[2], g[3],
                                                                        mx, my, mz, variant=variant))
        else:
            groups = list(pre) + [(mm, mx, my, mz)] + list(post)
            tot = [np.concatenate([np.asarray(g[k], np.float64) for g in groups]) for k in range(4)]
            be = np.float32(brute_force_binding_energy_fortran(tot[0], tot[1], tot[2], tot[3],
                                                               mx, my, mz, variant=variant))
        pairs += len(idx) * (len(idx) + n_ext)
        E = energy_step(be, vx[idx], vy[idx], vz[idx], vb[0], vb[1], vb[2], kappa)
        bound = E <= 0.
        be_out[idx] = be
        e_out[idx] = E
        it += 1
        new_idx = idx[bound]
        changed = len(new_idx) != len(idx)
        idx = new_idx
        hist.append(len(idx))
        if not changed:
            break
    mask = np.zeros(N, bool)
    mask[idx] = True
    cx, cy, cz, M = center_of_mass(idx, x, y, z, mass)
    vb_out = CM_velocity(M, idx, vx, vy, vz, mass) if vb_fixed is None else tuple(map(float, vb_fixed))
    return UnbindResult(mask, idx, be_out, e_out, it, M, (cx, cy, cz), vb_out, hist, pairs)


def energy_margin(E: np.ndarray, be32: np.ndarray, kappa: float) -> np.ndarray:
    """|E| / max(KE, |PE|): the relative distance from the E = 0 boundary used to bin mask
    differences (SURVEY.md §7 hard part 1; north_star's 1e-6 band)."""
    pe = np.abs((np.float32(np.asarray(be32, np.float32) * np.float32(G_const()))
                 * np.float32(kappa)).astype(np.float64))
    ke = np.abs(E + pe)
    scale = np.maximum(ke, pe)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(scale > 0, np.abs(E) / scale, np.inf)
    return r
