"""The potential stage of pyHALMA's halo loop without leaving the GPU (pyHALMA.py:1023-1067).

For every halo the reference gathers gas / DM / stars inside FACTOR_R12_POT x R1/2
(halo_gas.st_gas_dm_particles_inside), then runs halo_gas.RPS and
halo_gas.most_bound_particle on the gathered arrays -- three to nine f2py calls and as many
host round trips.  Here the snapshot is resident (gather.Snapshot), the gather leaves its
result in HBM, and the two unbinding plans read it there: per halo only the four RPS masses
and the most bound star come back to the host.

Exact all-pairs (no BRUTE_FORCE_LIM subsampling: the GPU does not need it).  The mass sums
are ordered float64 reductions on the device; the reference sums the same numbers with
numpy's pairwise sum, so they agree to ~1e-15 relative, not bitwise.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .gather import Snapshot
from .halo_gas import COLD_T
from .unbind import UnbindPlan


def _one(n):
    return np.array([0, n], np.int64)


def rps_on_device(g, vx, vy, vz, num_dm_species, *, mode=None, device=0, max_iter=1):
    """halo_gas.RPS (halo_gas.py:285-492) on a DeviceGather: (total gas mass, cold bound
    fraction, unbound cold mass, unbound hot mass)."""
    if num_dm_species > 1 and not g.split:
        raise ValueError("num_dm_species > 1 needs a gather with dm_heavy_min")
    if num_dm_species <= 1 and g.split:
        raise ValueError("num_dm_species = 1 sums the DM as one class: gather without dm_heavy_min")
    if g.n_gas == 0:
        return 0.0, 0.0, 0.0, 0.0
    mode = _lib.default_mode() if mode is None else _lib.mode_code(mode)
    ext = [g.dm_source()] + ([g.dm_source(light=True)] if num_dm_species > 1 else []) + [g.star_source()]
    with UnbindPlan(_one(g.n_gas), [_one(n) for n, _ in ext], mode=mode, split_classes=True, vb_fixed=True,
                    max_iter=max_iter, kappa=2.0, device=device) as plan:
        plan.upload_members_raw(*[g.col(0, k) for k in range(7)])
        plan.upload_temp_raw(g.col(0, 7), COLD_T)
        for k, (n, cols) in enumerate(ext):
            plan.upload_group_raw(k, *cols)
        plan.set_vb([vx, vy, vz])
        plan.run()
        h = plan.download(mask=False, be=False, energy=False, idx=False).halos[0]
    total = h.mass_initial
    return total, (h.cold_bound_mass / total if total != 0.0 else 0.0), h.unbound_cold_mass, h.unbound_hot_mass


def most_bound_on_device(g, *, mode=None, device=0):
    """halo_gas.most_bound_particle (halo_gas.py:498-634) on a DeviceGather made with
    dm_heavy_min (the DM split is unconditional there, :548): x, y, z, id of the star with the
    deepest potential."""
    if not g.split:
        raise ValueError("most_bound_particle splits the DM by species: gather with dm_heavy_min")
    if g.n_st == 0:
        raise ValueError("no star inside the radius")          # the reference's argmin of an empty array raises too
    mode = _lib.default_mode() if mode is None else _lib.mode_code(mode)
    ext = [g.gas_source(), g.dm_source(), g.dm_source(light=True)]
    zeros = np.zeros(g.n_st)
    with UnbindPlan(_one(g.n_st), [_one(n) for n, _ in ext], mode=mode, n_pre=3, split_classes=True, vb_fixed=True,
                    max_iter=1, kappa=1.0, device=device) as plan:
        z = zeros.ctypes.data
        plan.upload_members_raw(g.col(3, 0), g.col(3, 1), g.col(3, 2), z, z, z, g.col(3, 3))
        for k, (n, cols) in enumerate(ext):
            plan.upload_group_raw(k, *cols)
        plan.set_vb([0.0, 0.0, 0.0])
        plan.run()
        k = plan.download(mask=False, be=False, energy=False, idx=False).halos[0].most_bound
    x, y, z_, _, sid = g.snap.fetch_star(k)
    return x, y, z_, sid


def halo_potential_stage(snap: Snapshot, cx, cy, cz, R, rho_B, rete, vx, vy, vz, mass_dm_part, num_dm_species, *,
                         rps=True, most_bound=True, mode=None, device=0):
    """pyHALMA.py:1023-1067 for one halo: returns (RPS 4-tuple or None, most-bound 4-tuple or None)."""
    heavy_min = 0.9 * (mass_dm_part / 8)                                  # halo_gas.py:342,347
    out_rps = out_mb = None
    g = None
    if rps:
        g = snap.gather_device(cx, cy, cz, R, rho_B, rete, dm_heavy_min=heavy_min if num_dm_species > 1 else None)
        out_rps = rps_on_device(g, vx, vy, vz, num_dm_species, mode=mode, device=device)
    if most_bound:
        if g is None or not g.split:
            g = snap.gather_device(cx, cy, cz, R, rho_B, rete, dm_heavy_min=heavy_min)
        if g.n_st > 0:
            out_mb = most_bound_on_device(g, mode=mode, device=device)
    return out_rps, out_mb
