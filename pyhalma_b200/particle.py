"""Drop-in for the f2py extension module `fortran_modules.particle`.

The reference reaches the Fortran routines as `particle.particle.<routine>(...)`
(python_scripts/halo_gas.py:6,182,208).  This module offers the same attribute path, the
same positional signatures (fortran_modules/particle_subroutines.f90:466-469, :517-519;
the `!f2py depend(...)` directives at :485-488 keep ntotal / ntest as required arguments)
and the same return value (a fresh float32 array of length ntest), computed on the B200.

`ncores` is accepted and ignored.  Shape mismatches raise ValueError like the f2py shim.
`halo_shape` / `sigma_projections` (particle_subroutines.f90:12-461) are outside the
replaced path; when the original f2py module is importable they are forwarded to it.
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
    def _forward(name):
        try:
            import importlib
            orig = importlib.import_module("fortran_modules._particle_f2py")
        except Exception as exc:  # pragma: no cover - needs the reference build
            raise NotImplementedError(
                "particle.%s is outside the replaced hot path; keep the original f2py build as "
                "fortran_modules/_particle_f2py to forward it (INTEGRATION.md)" % name) from exc
        return getattr(orig.particle, name)

    @staticmethod
    def halo_shape(*args, **kw):  # pragma: no cover
        return particle._forward("halo_shape")(*args, **kw)

    @staticmethod
    def sigma_projections(*args, **kw):  # pragma: no cover
        return particle._forward("sigma_projections")(*args, **kw)
