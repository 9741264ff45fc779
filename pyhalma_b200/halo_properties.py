"""Host-side mirror of the hot-path functions of python_scripts/halo_properties.py.

  total_mass / center_of_mass / CM_velocity      halo_properties.py:16-60
  escape_velocity_unbinding_fortran              halo_properties.py:282-361
  escape_velocity_unbinding                      the same from :333 on, for inputs that
                                                 are already gathered, optionally iterated
  halo_shape_fortran                             halo_properties.py:852-866
  sigma_projections_fortran                      halo_properties.py:781-812
"""
from __future__ import annotations

import numpy as np

from . import halo_gas
from .particle import particle
from .unbind import unbind_halo


def total_mass(part_list, st_mass):
    return float(np.sum(np.asarray(st_mass, np.float64)[part_list]))


def center_of_mass(part_list, st_x, st_y, st_z, st_mass):
    m = np.asarray(st_mass, np.float64)[part_list]
    M = float(np.sum(m))
    if M > 0:
        return (float(np.sum(m * np.asarray(st_x)[part_list]) / M), float(np.sum(m * np.asarray(st_y)[part_list]) / M),
                float(np.sum(m * np.asarray(st_z)[part_list]) / M), M)
    return 0., 0., 0., 0.


def CM_velocity(M, part_list, st_vx, st_vy, st_vz, st_mass):  # noqa: N802
    m = np.asarray(st_mass, np.float64)[part_list]
    if M > 0.:
        return (float(np.sum(m * np.asarray(st_vx)[part_list]) / M), float(np.sum(m * np.asarray(st_vy)[part_list]) / M),
                float(np.sum(m * np.asarray(st_vz)[part_list]) / M))
    return 0., 0., 0.


def escape_velocity_unbinding(gas, stars, dm, vb, factor_v, *, max_iter=1, recompute_vb=False, mode=None,
                              device=0):
    """gas / dm = (x, y, z, mass); stars = (x, y, z, vx, vy, vz, mass); vb = bulk velocity.

    One pass with the given vb is halo_properties.py:333-361 (sources in the order gas,
    stars, DM; kappa = factor_v**2).  max_iter > 1 iterates it; recompute_vb=True then
    re-derives vb from the bound set each pass (CM_velocity).  Returns the UnbindResult."""
    gx, gy, gz, gm = gas
    sx, sy, sz, svx, svy, svz, sm = stars
    dx, dy, dz, dmass = dm
    return unbind_halo(sx, sy, sz, svx, svy, svz, sm, pre=[(gm, gx, gy, gz)], post=[(dmass, dx, dy, dz)],
                       kappa=factor_v ** 2, vb_fixed=None if recompute_vb else vb, max_iter=max_iter,
                       mode=mode, device=device)


def escape_velocity_unbinding_fortran(rete, L, ncoarse, grid_data, gas_data, masclet_dm_data, cx, cy, cz,
                                      vx, vy, vz, Rmax, part_list, st_x, st_y, st_z, st_vx, st_vy, st_vz,
                                      st_mass, factor_v, rho_B, *, mass_to_sun=None, mode=None, device=0):
    """Reference signature (halo_properties.py:282-286).  The gathers at :289-326 (DM inside
    Rmax, AMR gas -> particles through halo_gas.AMRgrid_to_particles, gas mass * rete**3) are
    host-side numpy like the reference; everything from :333 on runs on the GPU.
    mass_to_sun: masclet_framework.units.mass_to_sun (:293) when that package is importable, else it must be given."""
    if mass_to_sun is None:
        from . import gather
        mass_to_sun = gather.default_mass_to_sun()
    dm_x, dm_y, dm_z = masclet_dm_data[0], masclet_dm_data[1], masclet_dm_data[2]
    dm_mass = masclet_dm_data[3] * mass_to_sun
    inside = np.sqrt((dm_x - cx) ** 2 + (dm_y - cy) ** 2 + (dm_z - cz) ** 2) < Rmax
    dm = (dm_x[inside], dm_y[inside], dm_z[inside], dm_mass[inside])
    stars = tuple(np.asarray(a)[part_list] for a in (st_x, st_y, st_z, st_vx, st_vy, st_vz, st_mass))
    gas_x, gas_y, gas_z, _, _, _, gas_mass, _ = halo_gas.AMRgrid_to_particles(
        L, ncoarse, grid_data, gas_data, Rmax, cx, cy, cz, rho_B)
    inside = np.sqrt((gas_x - cx) ** 2 + (gas_y - cy) ** 2 + (gas_z - cz) ** 2) < Rmax
    gas = (gas_x[inside], gas_y[inside], gas_z[inside], gas_mass[inside] * rete ** 3)
    res = escape_velocity_unbinding(gas, stars, dm, (vx, vy, vz), factor_v, mode=mode, device=device)
    return res.mask


def halo_shape_fortran(part_list, st_x, st_y, st_z, st_mass, cx, cy, cz, RAD05, *, device=0):
    """Semi-axes (a, b, c) of the particle list, halo_properties.py:852-866."""
    x = np.float32(st_x[part_list] - cx)
    y = np.float32(st_y[part_list] - cy)
    z = np.float32(st_z[part_list] - cz)
    mass = np.float32(st_mass[part_list])
    eigenvalues = particle.halo_shape(np.int32(1), np.int32(len(part_list)), x, y, z, mass, device=device)
    return eigenvalues[0], eigenvalues[1], eigenvalues[2]


def sigma_projections_fortran(grid, n_cell, part_list, st_x, st_y, st_z, st_vx, st_vy, st_vz, vx, vy, vz, st_mass,
                              cx, cy, cz, R05x, R05y, R05z, ll, *, device=0):
    """halo_properties.py:781-812: float32 casts, velocities relative to (vx, vy, vz),
    part_list + 1 for the Fortran-style index base."""
    return particle.sigma_projections(
        np.int32(1), np.int32(len(part_list)), np.float32(grid), np.int32(n_cell), np.int32(1 + part_list),
        np.float32(st_x), np.float32(st_y), np.float32(st_z), np.float32(st_vx - vx), np.float32(st_vy - vy),
        np.float32(st_vz - vz), np.float32(st_mass), np.float32(cx), np.float32(cy), np.float32(cz),
        np.float32(R05x), np.float32(R05y), np.float32(R05z), np.float32(ll), device=device)
