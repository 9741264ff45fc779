"""Drop-in for the f2py extension module `fortran_modules.particle`.

The reference reaches the Fortran routines as `particle.particle.<routine>(...)`
(python_scripts/halo_gas.py:6,182,208).  This module offers the same attribute path, the
same positional signatures (fortran_modules/particle_subroutines.f90:466-469, :517-519;
the `!f2py depend(...)` directives at :485-488 keep ntotal / ntest as required arguments)
and the same return value (a fresh float32 array of length ntest), computed on the B200.

`ncores` is accepted and ignored.  Shape mismatches raise ValueError like the f2py shim.
`halo_shape` and `sigma_projections` (particle_subroutines.f90:160-214, :217-461), the other
two routines of the module, run on the GPU as well, so the swap needs no
Fortran toolchain.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _as_f32(a, n, name):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError("%s: expected a rank-1 array" % name)
    if a.shape[0] != int(n):
        raise ValueError("%s: 0-th dimension must be fixed to %d but got %d" % (name, int(n), a.shape[0]))
    return a


def _potential(total_mass, total_x, total_y, total_z, ntotal, test_x, test_y, test_z, ntest, mode, device):
    ntotal, ntest = int(ntotal), int(ntest)
    if ntotal < 0 or ntest < 0:
        raise ValueError("negative size")
    tm = _as_f32(total_mass, ntotal, "total_mass")
    tx = _as_f32(total_x, ntotal, "total_x")
    ty = _as_f32(total_y, ntotal, "total_y")
    tz = _as_f32(total_z, ntotal, "total_z")
    sx = _as_f32(test_x, ntest, "test_x")
    sy = _as_f32(test_y, ntest, "test_y")
    sz = _as_f32(test_z, ntest, "test_z")
    out = np.zeros(ntest, dtype=np.float32)
    mode = _lib.default_mode() if mode is None else _lib.mode_code(mode)
    _lib.check(_lib.lib().halma_potential_f32(device, mode, tm.ctypes.data, tx.ctypes.data, ty.ctypes.data,
                                              tz.ctypes.data, ntotal, sx.ctypes.data, sy.ctypes.data,
                                              sz.ctypes.data, ntest, out.ctypes.data))
    return out


class particle:  # noqa: N801  (the Fortran MODULE name inside the f2py extension)
    """`particle.particle`: namespace of the Fortran module `particle`."""

    @staticmethod
    def brute_force_binding_energy(ncores, ntotal, total_mass, total_x, total_y, total_z,
                                   ntest, test_x, test_y, test_z, *, mode=None, device=0):
        """particle_subroutines.f90:466-514 on the GPU; returns float32[ntest]."""
        del ncores
        return _potential(total_mass, total_x, total_y, total_z, ntotal, test_x, test_y, test_z, ntest,
                          mode, device)

    @staticmethod
    def serial_brute_force_binding_energy(ntotal, total_mass, total_x, total_y, total_z,
                                          ntest, test_x, test_y, test_z, *, mode=None, device=0):
        """particle_subroutines.f90:517-556 (same arithmetic as the OpenMP routine)."""
        return _potential(total_mass, total_x, total_y, total_z, ntotal, test_x, test_y, test_z, ntest,
                          mode, device)

    @staticmethod
    def halo_shape(ncore, npart, x, y, z, mass, *, device=0):
        """particle_subroutines.f90:160-214: semi-axes a >= b >= c (float32[3]) of the
        mass-weighted second-moment tensor of the centred positions."""
        del ncore
        npart = int(npart)
        xs = [_as_f32(a, npart, n) for a, n in ((x, "x"), (y, "y"), (z, "z"), (mass, "mass"))]
        out = np.zeros(3, dtype=np.float32)
        _lib.check(_lib.lib().halma_halo_shape_f32(device, *[a.ctypes.data for a in xs], npart, out.ctypes.data))
        return out

    @staticmethod
    def sigma_projections(ncore, npart, grid, n_cell, part_list, st_x, st_y, st_z, st_vx, st_vy, st_vz, st_mass,
                          cx, cy, cz, R05x, R05y, R05z, ll, *, device=0):
        """particle_subroutines.f90:217-461.  part_list is 1-based, as the reference's wrapper
        passes it (halo_properties.py:787).  Returns (SIG_1D_x_05, SIG_1D_y_05, SIG_1D_z_05,
        V_sigma, lambda) as Python floats, like f2py."""
        del ncore
        npart, n_cell = int(npart), int(n_cell)
        g = _as_f32(grid, n_cell, "grid")
        pl = np.ascontiguousarray(part_list, dtype=np.int32)
        if pl.ndim != 1 or pl.shape[0] != npart:
            raise ValueError("part_list: 0-th dimension must be fixed to %d but got %d" % (npart, pl.shape[0]))
        pl = pl - np.int32(1)
        n_all = len(st_x)
        arrs = [_as_f32(a, n_all, n) for a, n in ((st_x, "st_x"), (st_y, "st_y"), (st_z, "st_z"), (st_vx, "st_vx"),
                                                  (st_vy, "st_vy"), (st_vz, "st_vz"), (st_mass, "st_mass"))]
        if npart and (pl.min() < 0 or pl.max() >= n_all):
            raise IndexError("part_list out of range")
        out = np.zeros(5, dtype=np.float32)
        _lib.check(_lib.lib().halma_sigma_projections_f32(
            device, npart, g.ctypes.data, n_cell, pl.ctypes.data, n_all, *[a.ctypes.data for a in arrs],
            float(cx), float(cy), float(cz), float(R05x), float(R05y), float(R05z), float(ll), out.ctypes.data))
        return tuple(float(v) for v in out)
