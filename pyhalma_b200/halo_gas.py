"""Host-side mirror of the hot-path functions of python_scripts/halo_gas.py.

Same names, argument order and return values as the reference, so the parity tests read
like calls into the reference:

  brute_force_binding_energy_fortran          halo_gas.py:164-187
  serial_brute_force_binding_energy_fortran   halo_gas.py:191-215
  RPS                                         halo_gas.py:285-492
  most_bound_particle                         halo_gas.py:498-634
  AMRgrid_to_particles                        halo_gas.py:56-141   (the gather before the path,
  st_gas_dm_particles_inside                  halo_gas.py:223-277   SURVEY.md §8f-3)

All potential sums run on the GPU through libhalma_unbind (no CPU fallback).  The
Monte-Carlo subsampling above BRUTE_FORCE_LIM (halo_gas.py:306-321 and seven more sites)
stays on the host and draws from the global numpy RNG in the reference's order, so a
seeded call reproduces the reference's sample.
"""
from __future__ import annotations

import numpy as np

from . import _lib, gather
from .particle import particle
from .unbind import G_CONST, UnbindPlan

COLD_T = 5 * 1e4          # halo_gas.py:479-480


def brute_force_binding_energy_fortran(total_mass, total_x, total_y, total_z, test_x, test_y, test_z,
                                       *, mode=None, device=0):
    ntotal_f90 = np.int32(len(total_mass))
    ntest_f90 = np.int32(len(test_x))
    if ntest_f90 == 0:
        return np.array([])                     # halo_gas.py:169-170 (float64 empty, like the reference)
    cast = [np.asarray(a).astype(np.float32) for a in
            (total_mass, total_x, total_y, total_z, test_x, test_y, test_z)]
    return particle.brute_force_binding_energy(np.int32(1), ntotal_f90, *cast[:4], ntest_f90, *cast[4:],
                                               mode=mode, device=device)


def serial_brute_force_binding_energy_fortran(total_mass, total_x, total_y, total_z, test_x, test_y, test_z,
                                              *, mode=None, device=0):
    return brute_force_binding_energy_fortran(total_mass, total_x, total_y, total_z, test_x, test_y, test_z,
                                              mode=mode, device=device)


def _class_sum(acc, src_m, src_x, src_y, src_z, tx, ty, tz, lim, mode, device):
    """One source class: exact below `lim`, else the reference's with-replacement subsample
    rescaled by n / nsample (halo_gas.py:306-328)."""
    n = len(src_x)
    if n > lim:
        nsample = np.max([lim, int(0.01 * n)])
        sample = np.random.choice(np.arange(n), nsample, replace=True)
        be = brute_force_binding_energy_fortran(src_m[sample], src_x[sample], src_y[sample], src_z[sample],
                                                tx, ty, tz, mode=mode, device=device)
        be = be * n / nsample
        acc += be
    elif n > 0:
        acc += brute_force_binding_energy_fortran(src_m, src_x, src_y, src_z, tx, ty, tz, mode=mode,
                                                  device=device)


def _split_dm(dm_x, dm_y, dm_z, dm_mass, mass_dm_part):
    heavy = dm_mass >= 0.9 * (mass_dm_part / 8)            # halo_gas.py:340-350
    light = np.logical_not(heavy)
    return ((dm_x[heavy], dm_y[heavy], dm_z[heavy], dm_mass[heavy]),
            (dm_x[light], dm_y[light], dm_z[light], dm_mass[light]))


def _energy(be32, vx, vy, vz, vbx, vby, vbz, kappa):
    be = -be32
    be *= G_CONST
    be *= kappa
    ke = 0.5 * ((vx - vbx) ** 2 + (vy - vby) ** 2 + (vz - vbz) ** 2)
    return ke + be


def _rps_sources(dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z, st_mass, mass_dm_part, num_dm_species):
    """External classes of RPS in the reference's order: DM (heavy, light) then stars."""
    groups = []
    if num_dm_species > 1:
        heavy, light = _split_dm(dm_x, dm_y, dm_z, dm_mass, mass_dm_part)
        groups += [(heavy[3], heavy[0], heavy[1], heavy[2], False),      # never sampled (:353)
                   (light[3], light[0], light[1], light[2], True)]
    else:
        groups += [(dm_mass, dm_x, dm_y, dm_z, True)]
    groups += [(st_mass, st_x, st_y, st_z, True)]
    return groups


def RPS(gas_x, gas_y, gas_z, gas_vx, gas_vy, gas_vz, gas_mass, gas_temp,  # noqa: N802
        dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z, st_mass, vx, vy, vz, BRUTE_FORCE_LIM,
        mass_dm_part, num_dm_species, *, mode=None, device=0, fused=True, max_iter=1, return_mask=False):
    """Bound / unbound gas masses of one halo (halo_gas.py:285-492).

    fused=True evaluates all source classes, the energy step and the mask in one
    device-resident plan when no class needs subsampling; otherwise the classes are summed
    by separate kernel calls exactly like the reference.  max_iter > 1 iterates the gas
    self-term to a fixed point; the reference is max_iter = 1.
    """
    ngas = len(gas_x)
    ext = _rps_sources(dm_x, dm_y, dm_z, dm_mass, st_x, st_y, st_z, st_mass, mass_dm_part, num_dm_species)
    sampled = ngas > BRUTE_FORCE_LIM or any(len(g[0]) > BRUTE_FORCE_LIM and g[4] for g in ext)
    mode = _lib.default_mode() if mode is None else _lib.mode_code(mode)
    if ngas > 0 and fused and not sampled:
        offs = np.array([0, ngas], np.int64)
        with UnbindPlan(offs, [np.array([0, len(g[0])], np.int64) for g in ext], mode=mode,
                        split_classes=True, vb_fixed=True, max_iter=max_iter, kappa=2.0, device=device) as plan:
            plan.upload_members(gas_x, gas_y, gas_z, gas_vx, gas_vy, gas_vz, gas_mass)
            for k, g in enumerate(ext):
                plan.upload_group(k, g[0], g[1], g[2], g[3])
            plan.set_vb([vx, vy, vz])
            plan.run()
            res = plan.download(idx=False)
        total_energy = res.energy
        bound = res.mask.astype(bool)
        unbound = np.logical_not(bound) & ~np.isnan(total_energy)
    else:
        if max_iter != 1:
            raise ValueError("iterating RPS needs the fused path (no subsampling)")
        binding_energy = np.zeros((ngas,), dtype=np.float32)
        _class_sum(binding_energy, gas_mass, gas_x, gas_y, gas_z, gas_x, gas_y, gas_z, BRUTE_FORCE_LIM,
                   mode, device)
        for g in ext:
            _class_sum(binding_energy, g[0], g[1], g[2], g[3], gas_x, gas_y, gas_z,
                       BRUTE_FORCE_LIM if g[4] else np.inf, mode, device)
        total_energy = _energy(binding_energy, gas_vx, gas_vy, gas_vz, vx, vy, vz, 2.)
        unbound = total_energy > 0.
        bound = total_energy <= 0.
    cold = gas_temp < COLD_T
    hot = gas_temp >= COLD_T
    total_gas_mass = np.sum(gas_mass)
    cold_bound_gas_mass = np.sum(gas_mass[cold * bound])
    frac_cold_gas_mass = cold_bound_gas_mass / total_gas_mass if total_gas_mass != 0. else 0.
    unbound_cold_gas_mass = np.sum(gas_mass[unbound * cold])
    unbound_hot_gas_mass = np.sum(gas_mass[unbound * hot])
    out = (total_gas_mass, frac_cold_gas_mass, unbound_cold_gas_mass, unbound_hot_gas_mass)
    return out + (bound,) if return_mask else out


def most_bound_particle(gas_x, gas_y, gas_z, gas_mass, dm_x, dm_y, dm_z, dm_mass,
                        st_x, st_y, st_z, st_mass, st_oripa, BRUTE_FORCE_LIM, mass_dm_part,
                        *, mode=None, device=0, fused=True):
    """Position and id of the star at the potential minimum (halo_gas.py:498-634).  Class
    order: gas, heavy DM, light DM, stars; the DM split is unconditional here (:548).

    fused=True (and no class above BRUTE_FORCE_LIM): one device-resident plan evaluates the
    four classes and the arg-min on the GPU; otherwise one kernel call per class and a host
    arg-min, like the reference."""
    nst = len(st_x)
    mode = _lib.default_mode() if mode is None else _lib.mode_code(mode)
    heavy, light = _split_dm(dm_x, dm_y, dm_z, dm_mass, mass_dm_part)
    sampled = (len(gas_x) > BRUTE_FORCE_LIM or len(light[0]) > BRUTE_FORCE_LIM or nst > BRUTE_FORCE_LIM)
    if fused and not sampled and nst > 0:
        ext = [(gas_mass, gas_x, gas_y, gas_z), (heavy[3], heavy[0], heavy[1], heavy[2]),
               (light[3], light[0], light[1], light[2])]
        zeros = np.zeros(nst)
        with UnbindPlan(np.array([0, nst], np.int64), [np.array([0, len(g[0])], np.int64) for g in ext], mode=mode,
                        n_pre=3, split_classes=True, vb_fixed=True, max_iter=1, kappa=1.0, device=device) as plan:
            plan.upload_members(st_x, st_y, st_z, zeros, zeros, zeros, st_mass)
            for k, g in enumerate(ext):
                plan.upload_group(k, g[0], g[1], g[2], g[3])
            plan.set_vb([0., 0., 0.])
            plan.run()
            k = plan.download(mask=False, be=False, energy=False, idx=False).halos[0].most_bound
        return st_x[k], st_y[k], st_z[k], st_oripa[k]
    binding_energy = np.zeros((nst,), dtype=np.float32)
    _class_sum(binding_energy, gas_mass, gas_x, gas_y, gas_z, st_x, st_y, st_z, BRUTE_FORCE_LIM, mode, device)
    _class_sum(binding_energy, heavy[3], heavy[0], heavy[1], heavy[2], st_x, st_y, st_z, np.inf, mode, device)
    _class_sum(binding_energy, light[3], light[0], light[1], light[2], st_x, st_y, st_z, BRUTE_FORCE_LIM,
               mode, device)
    _class_sum(binding_energy, st_mass, st_x, st_y, st_z, st_x, st_y, st_z, BRUTE_FORCE_LIM, mode, device)
    binding_energy = -binding_energy
    k = np.argmin(binding_energy)
    return st_x[k], st_y[k], st_z[k], st_oripa[k]


def AMRgrid_to_particles(L, ncoarse, grid_data, gas_data, Rrps, cx, cy, cz, rho_B, *, device=0):  # noqa: N802
    """python_scripts/halo_gas.py:56-141: the gas cells (level >= 1, not refined, not overlapped)
    whose centres lie strictly inside the box of half-width Rrps, as pseudo-particles
    x, y, z, vx, vy, vz (km/s), mass (comoving: no rete**3 yet), temp -- float64, in the
    reference's order.  The snapshot is uploaded on first use and stays on the GPU."""
    snap = gather.snapshot_for(L, ncoarse, grid_data, gas_data, None, None, 1.0, device)
    return _box_gas(snap, cx, cy, cz, Rrps, rho_B)


def _box_gas(snap, cx, cy, cz, Rrps, rho_B):
    # the box selection alone, made on the device (halma_snapshot_gather_box)
    return snap.gather_box(cx, cy, cz, Rrps, rho_B, 1.0)


def parallel_inside(array_x, array_y, array_z, R, cx, cy, cz):
    """python_scripts/halo_gas.py:216-218."""
    return np.sqrt((array_x - cx) ** 2 + (array_y - cy) ** 2 + (array_z - cz) ** 2) < R


def st_gas_dm_particles_inside(rete, L, ncoarse, grid_data, gas_data, masclet_dm_data, masclet_st_data,
                               st_kdtree, dm_kdtree, cx, cy, cz, R, rho_B, *, mass_to_sun=None, device=0):
    """python_scripts/halo_gas.py:223-277, same arguments and the same 17-tuple:
    gas x, y, z, vx, vy, vz, mass (x rete**3), temp; DM x, y, z, mass; stars x, y, z, mass, oripa.
    One device call per halo against the resident snapshot.  The KD-trees are accepted for
    signature compatibility and not used: the ball query is a brute-force pass over the
    resident particles (24 B per particle at HBM speed).  Gas is bit-identical to the
    reference; DM and stars are the same set in ascending index order."""
    del st_kdtree, dm_kdtree
    if mass_to_sun is None:
        mass_to_sun = gather.default_mass_to_sun()
    snap = gather.snapshot_for(L, ncoarse, grid_data, gas_data, masclet_dm_data, masclet_st_data, mass_to_sun,
                               device)
    return snap.gather(cx, cy, cz, R, rho_B, rete)
