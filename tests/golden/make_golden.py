"""Generate tests/golden/*.npz by running the REFERENCE's own Python drivers.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

What is real and what is stubbed:
  * REAL: /root/reference/python_scripts/halo_gas.py (RPS, most_bound_particle,
    brute_force_binding_energy_fortran) and halo_properties.py
    (escape_velocity_unbinding_fortran, center_of_mass, CM_velocity, total_mass) are
    imported and executed unmodified.
  * STUB: `masclet_framework` and `matplotlib` (absent, un-vendored; only imported at
    module top, and `units.mass_to_sun` = 1 so DM masses are already in Msun), and
    `halo_gas.AMRgrid_to_particles` (needs MASCLET AMR files) which is
    replaced by a function returning the synthetic gas particles.
  * ORACLE: `fortran_modules.particle` -- the f2py module that cannot be built here (no
    Fortran compiler) -- is bound to oracle/oracle.py's C restatement.  Every kernel call
    is recorded, so the fixtures also pin the order and sizes of the calls.

The fixtures therefore pin the reference's driver semantics (class order, dtype
promotions, energy step, mask, mass sums, argmin, sampling RNG use) around the kernel;
the kernel arithmetic itself stays "parity unpinned" (see oracle/halma_oracle.c).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from pyhalma_b200 import synth  # noqa: E402

CALLS = []
CALL_INPUTS = []          # filled only while RECORD_INPUTS is on (call_sequences())
RECORD_INPUTS = False


def _record(ntotal, ntest, out, args):
    CALLS.append((int(ntotal), int(ntest), out.copy()))
    if RECORD_INPUTS:
        # exactly what reached the f2py boundary: dtypes as the reference's wrapper made them (halo_gas.py:172-178)
        CALL_INPUTS.append([np.array(a) for a in args])


class _RecordingParticle:
    @staticmethod
    def brute_force_binding_energy(ncores, ntotal, tm, tx, ty, tz, ntest, sx, sy, sz):
        out = O.brute_force_binding_energy(ncores, ntotal, tm, tx, ty, tz, ntest, sx, sy, sz)
        _record(ntotal, ntest, out, (tm, tx, ty, tz, sx, sy, sz))
        return out

    @staticmethod
    def serial_brute_force_binding_energy(ntotal, tm, tx, ty, tz, ntest, sx, sy, sz):
        out = O.serial_brute_force_binding_energy(ntotal, tm, tx, ty, tz, ntest, sx, sy, sz)
        _record(ntotal, ntest, out, (tm, tx, ty, tz, sx, sy, sz))
        return out

    @staticmethod
    def halo_shape(ncore, npart, x, y, z, mass):
        return O.halo_shape(ncore, npart, x, y, z, mass)

    @staticmethod
    def sigma_projections(*args):
        return O.sigma_projections(*args)


def install_stubs():
    mf = types.ModuleType("masclet_framework")
    for sub in ("units", "tools", "particles", "particle2grid"):
        m = types.ModuleType("masclet_framework." + sub)
        setattr(mf, sub, m)
        sys.modules["masclet_framework." + sub] = m
    mf.units.mass_to_sun = 1.0
    # un-vendored masclet_framework.tools helpers the gather calls (halo_gas.py:91-92), restated
    # from their published behaviour in oracle/gather.py (see its header)
    from oracle import gather as OG
    mf.tools.which_patches_inside_box = OG.which_patches_inside_box
    mf.tools.create_vector_levels = OG.create_vector_levels
    sys.modules["masclet_framework"] = mf
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    fm = types.ModuleType("fortran_modules")
    fm.__path__ = []
    part = types.ModuleType("fortran_modules.particle")
    part.particle = _RecordingParticle
    fm.particle = part
    sys.modules["fortran_modules"] = fm
    sys.modules["fortran_modules.particle"] = part
    sys.path.insert(0, REF)


def pack_calls():
    d = {"n_calls": np.int64(len(CALLS))}
    for k, (nt, ns, out) in enumerate(CALLS):
        d["call%d_ntotal" % k] = np.int64(nt)
        d["call%d_ntest" % k] = np.int64(ns)
        d["call%d_be" % k] = out
    CALLS.clear()
    return d


def small_case(seed, two_species, dm_mass=8e7 / 8):
    rng = np.random.default_rng(seed)
    stars = synth.plummer_stars(260, 2 * synth.KPC, 1e6, rng)
    synth.add_coincident_pairs(stars, 3, rng)
    gas = synth.lattice_gas(420, synth.CELL, rng, m_total=6e7)
    dm = synth.dm_cloud(310, 5 * synth.KPC, dm_mass, rng, two_species=two_species)
    return stars, gas, dm


def inputs_dict(stars, gas, dm):
    d = {}
    for name, p, keys in (("st", stars, ("x", "y", "z", "vx", "vy", "vz", "mass")),
                          ("gas", gas, ("x", "y", "z", "vx", "vy", "vz", "mass", "temp")),
                          ("dm", dm, ("x", "y", "z", "mass"))):
        for k in keys:
            d["%s_%s" % (name, k)] = getattr(p, k)
    return d


def main():
    install_stubs()
    from python_scripts import halo_gas, halo_properties  # the reference, unmodified

    mass_dm_part = 8e7
    for tag, two_species, lim, seed in (("rps_one_dm", False, 5000, 11),
                                        ("rps_two_dm", True, 5000, 12),
                                        ("rps_sampled", True, 100, 13)):
        stars, gas, dm = small_case(seed, two_species)
        part_list = np.arange(len(stars))
        cx, cy, cz, M = halo_properties.center_of_mass(part_list, stars.x, stars.y, stars.z, stars.mass)
        vb = halo_properties.CM_velocity(M, part_list, stars.vx, stars.vy, stars.vz, stars.mass)
        num_dm_species = 2 if two_species else 1
        np.random.seed(4242)
        out = halo_gas.RPS(gas.x, gas.y, gas.z, gas.vx, gas.vy, gas.vz, gas.mass, gas.temp,
                           dm.x, dm.y, dm.z, dm.mass, stars.x, stars.y, stars.z, stars.mass,
                           vb[0], vb[1], vb[2], lim, mass_dm_part, num_dm_species)
        d = inputs_dict(stars, gas, dm)
        d.update(pack_calls())
        d.update(rps_out=np.array(out, dtype=np.float64), vb=np.array(vb), com=np.array([cx, cy, cz]),
                 M=np.float64(M), lim=np.int64(lim), mass_dm_part=np.float64(mass_dm_part),
                 num_dm_species=np.int64(num_dm_species), np_seed=np.int64(4242))
        np.random.seed(4242)
        mb = halo_gas.most_bound_particle(gas.x, gas.y, gas.z, gas.mass, dm.x, dm.y, dm.z, dm.mass,
                                          stars.x, stars.y, stars.z, stars.mass,
                                          np.arange(len(stars)) + 1000, lim, mass_dm_part)
        mbc = pack_calls()
        d.update({"mb_" + k: v for k, v in mbc.items()})
        d.update(mb_out=np.array(mb, dtype=np.float64))
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **d)
        print(tag, "RPS ->", out, " most_bound ->", mb)

    # stellar one-pass unbinding, halo_properties.py:282-361
    stars, gas, dm = small_case(21, False, dm_mass=8e7 / 512)   # light DM: mixed bound/unbound mask
    rete, rho_B, Rmax, factor_v = 0.8, 1.0, 1.0, 3.0
    # :326 multiplies the gathered gas mass by rete**3 in place: hand it mass / rete**3
    gas_mass_pre = gas.mass / rete ** 3
    halo_gas.AMRgrid_to_particles = lambda L, ncoarse, grid_data, gas_data, R, cx, cy, cz, rho: (
        gas.x.copy(), gas.y.copy(), gas.z.copy(), None, None, None, gas_mass_pre.copy(), None)
    # global star arrays larger than the halo, gathered through part_list (:302-308)
    rng = np.random.default_rng(5)
    n_glob = 400
    part_list = np.sort(rng.choice(n_glob, len(stars), replace=False))
    glob = {k: rng.normal(0, 1, n_glob) for k in ("x", "y", "z", "vx", "vy", "vz", "mass")}
    for k in glob:
        glob[k][part_list] = getattr(stars, k)
    cx, cy, cz, M = halo_properties.center_of_mass(part_list, glob["x"], glob["y"], glob["z"], glob["mass"])
    vb = halo_properties.CM_velocity(M, part_list, glob["vx"], glob["vy"], glob["vz"], glob["mass"])
    # DM further than Rmax is dropped at :289-299: add two far particles that must be ignored
    dmx = np.concatenate((dm.x, [cx + 5.0, cx - 7.0]))
    dmy = np.concatenate((dm.y, [cy, cy]))
    dmz = np.concatenate((dm.z, [cz, cz]))
    dmm = np.concatenate((dm.mass, [1e15, 1e15]))
    bound = halo_properties.escape_velocity_unbinding_fortran(
        rete, 40.0, 128, None, None, (dmx, dmy, dmz, dmm), cx, cy, cz, vb[0], vb[1], vb[2], Rmax,
        part_list, glob["x"], glob["y"], glob["z"], glob["vx"], glob["vy"], glob["vz"], glob["mass"],
        factor_v, rho_B)
    d = inputs_dict(stars, gas, dm)
    d.update(pack_calls())
    # the gas mass the kernel actually saw (after the in-place rete**3 scaling at :326)
    d["gas_mass_seen"] = gas_mass_pre * rete ** 3
    d.update(bound=np.asarray(bound), vb=np.array(vb), com=np.array([cx, cy, cz]), M=np.float64(M),
             factor_v=np.float64(factor_v), part_list=part_list)
    np.savez_compressed(os.path.join(HERE, "stellar_onepass.npz"), **d)
    print("stellar_onepass bound", int(np.sum(bound)), "of", len(bound))


def shape_and_sigma():
    """halo_properties.halo_shape_fortran (:852-866) and sigma_projections_fortran (:781-812),
    the reference's own wrappers, on a triaxial rotating toy galaxy."""
    from python_scripts import halo_properties
    rng = np.random.default_rng(31)
    n_glob, n = 6000, 4500
    part_list = np.sort(rng.choice(n_glob, n, replace=False))
    st_x = 3.0 + rng.normal(0, 3e-3, n_glob)
    st_y = -7.0 + rng.normal(0, 2e-3, n_glob)
    st_z = 11.0 + rng.normal(0, 1e-3, n_glob)
    st_mass = rng.uniform(0.5e6, 2e6, n_glob)
    st_vx = 150.0 + rng.normal(0, 60, n_glob) - 2.0e4 * (st_y + 7.0)
    st_vy = -80.0 + rng.normal(0, 60, n_glob) + 2.0e4 * (st_x - 3.0)
    st_vz = 40.0 + rng.normal(0, 40, n_glob)
    cx, cy, cz, M = halo_properties.center_of_mass(part_list, st_x, st_y, st_z, st_mass)
    vx, vy, vz = halo_properties.CM_velocity(M, part_list, st_vx, st_vy, st_vz, st_mass)
    rad05 = 3e-3
    abc = halo_properties.halo_shape_fortran(part_list, st_x, st_y, st_z, st_mass, cx, cy, cz, rad05)
    ll = 0.5e-3
    n_cell = 25
    grid = (np.arange(n_cell) - n_cell // 2) * ll
    sig = halo_properties.sigma_projections_fortran(grid, n_cell, part_list, st_x, st_y, st_z, st_vx, st_vy, st_vz,
                                                    vx, vy, vz, st_mass, cx, cy, cz, rad05, 0.8 * rad05, 0.6 * rad05, ll)
    np.savez_compressed(os.path.join(HERE, "shape_sigma.npz"), part_list=part_list, st_x=st_x, st_y=st_y, st_z=st_z,
                        st_vx=st_vx, st_vy=st_vy, st_vz=st_vz, st_mass=st_mass, com=np.array([cx, cy, cz]),
                        vb=np.array([vx, vy, vz]), rad05=np.float64(rad05), ll=np.float64(ll), grid=grid,
                        n_cell=np.int64(n_cell), abc=np.array(abc, dtype=np.float32),
                        sigma=np.array(sig, dtype=np.float32))
    print("shape_sigma abc", abc, "sigma", sig)


GATHER_CASES = (  # (tag, radius in Mpc, centre offset in Mpc)
    ("r10", 0.010, (0.0, 0.0, 0.0)),
    ("r25off", 0.025, (0.004, -0.006, 0.002)),
    ("r60", 0.060, (-0.01, 0.0, 0.02)),
    ("tiny", 0.0004, (0.0, 0.0, 0.0)),
)
GATHER_SNAPSHOT = dict(n_levels=7, n_dm=20_000, n_st=30_000)
GATHER_FULL_MAX = 600          # cases with more gas particles than this are stored as sha256 digests


def gather_amr():
    """halo_gas.st_gas_dm_particles_inside (:223-277), the reference's own function with the real
    numba patch_to_particles and real scipy KD-trees, on synth.amr_snapshot(**GATHER_SNAPSHOT).
    The inputs are regenerated from the seed by the tests; only outputs are stored: in canonical
    order (oracle/gather.py::canonical_gather), as arrays for small cases and as sha256 digests
    of the arrays' bytes for large ones."""
    import importlib

    from scipy.spatial import KDTree

    from oracle import gather as OG
    from python_scripts import halo_gas
    importlib.reload(halo_gas)                      # undo the AMRgrid_to_particles stub of main()
    snap = synth.amr_snapshot(**GATHER_SNAPSHOT)
    dm_tree = KDTree(np.array(snap.masclet_dm_data[:3]).T)
    st_tree = KDTree(np.array(snap.masclet_st_data[:3]).T)
    d = {"n_cells": np.int64(snap.n_cells), "tags": np.array([c[0] for c in GATHER_CASES])}
    for tag, R, off in GATHER_CASES:
        cx, cy, cz = (snap.centre[k] + off[k] for k in range(3))
        out = halo_gas.st_gas_dm_particles_inside(snap.rete, snap.L, snap.ncoarse, snap.grid_data, snap.gas_data,
                                                  snap.masclet_dm_data, snap.masclet_st_data, st_tree, dm_tree,
                                                  cx, cy, cz, R, snap.rho_B)
        out = OG.canonical_gather(out)
        full = len(out[0]) <= GATHER_FULL_MAX
        for n, a in zip(OG.GATHER_NAMES, out):
            assert a.dtype in (np.float64, np.int64), (n, a.dtype)
            if full and n.startswith("gas"):
                d["%s_%s" % (tag, n)] = a
            d["%s_sha_%s" % (tag, n)] = OG.digest(a)
        d[tag + "_args"] = np.array([cx, cy, cz, R])
        d[tag + "_counts"] = np.array([len(out[0]), len(out[8]), len(out[12])])
        print("gather", tag, "gas", len(out[0]), "dm", len(out[8]), "stars", len(out[12]))
    np.savez_compressed(os.path.join(HERE, "gather_amr.npz"), **d)


STELLAR_AMR = dict(Rmax=0.012, r_members=0.006, factor_v=3.0)


def stellar_amr():
    """halo_properties.escape_velocity_unbinding_fortran (:282-361) END TO END on the synthetic AMR
    snapshot: the reference's own DM mask (:289-299), its AMR gather (halo_gas.AMRgrid_to_particles with
    the real numba patch_to_particles, :313-326) and its energy step, around the oracle kernel.
    Inputs are regenerated from the seed by the tests."""
    from python_scripts import halo_gas, halo_properties
    assert halo_gas.AMRgrid_to_particles.__module__ == "python_scripts.halo_gas"
    snap = synth.amr_snapshot(**GATHER_SNAPSHOT)
    st = snap.masclet_st_data
    cx0, cy0, cz0 = snap.centre
    r = np.sqrt((st[0] - cx0) ** 2 + (st[1] - cy0) ** 2 + (st[2] - cz0) ** 2)
    part_list = np.nonzero(r < STELLAR_AMR["r_members"])[0]
    cx, cy, cz, M = halo_properties.center_of_mass(part_list, st[0], st[1], st[2], st[6])
    vb = halo_properties.CM_velocity(M, part_list, st[3], st[4], st[5], st[6])
    CALLS.clear()
    bound = halo_properties.escape_velocity_unbinding_fortran(
        snap.rete, snap.L, snap.ncoarse, snap.grid_data, snap.gas_data, snap.masclet_dm_data, cx, cy, cz,
        vb[0], vb[1], vb[2], STELLAR_AMR["Rmax"], part_list, st[0], st[1], st[2], st[3], st[4], st[5], st[6],
        STELLAR_AMR["factor_v"], snap.rho_B)
    d = pack_calls()
    d.update(bound=np.asarray(bound), part_list=part_list, com=np.array([cx, cy, cz]), vb=np.array(vb),
             Rmax=np.float64(STELLAR_AMR["Rmax"]), factor_v=np.float64(STELLAR_AMR["factor_v"]))
    np.savez_compressed(os.path.join(HERE, "stellar_amr.npz"), **d)
    print("stellar_amr members", len(part_list), "bound", int(np.sum(bound)), "sources", int(d["call0_ntotal"]))


def call_sequences():
    """The kernel-call SEQUENCES of the reference's drivers with their inputs: every call that
    halo_gas.RPS, halo_gas.most_bound_particle and halo_properties.escape_velocity_unbinding_fortran make
    through `particle.particle.brute_force_binding_energy` on the small cases above, in order, with the
    float32 arrays that crossed the f2py boundary and the result the (oracle-backed) module returned.
    tests/test_gpu_unbind.py replays them through the drop-in package in EXACT mode, bit for bit.
    Runs the same cases as main() again (same seeds) and checks the outputs against the fixtures
    main() wrote, so the two files cannot drift apart."""
    global RECORD_INPUTS
    from python_scripts import halo_gas
    RECORD_INPUTS = True
    d = {}
    mass_dm_part = 8e7
    try:
        for tag, two_species, lim, seed in (("rps_one_dm", False, 5000, 11), ("rps_two_dm", True, 5000, 12),
                                            ("rps_sampled", True, 100, 13)):
            g = dict(np.load(os.path.join(HERE, tag + ".npz")))
            stars, gas, dm = small_case(seed, two_species)
            vb = g["vb"]
            for part, run in (("rps", lambda: halo_gas.RPS(
                    gas.x, gas.y, gas.z, gas.vx, gas.vy, gas.vz, gas.mass, gas.temp, dm.x, dm.y, dm.z, dm.mass,
                    stars.x, stars.y, stars.z, stars.mass, vb[0], vb[1], vb[2], lim, mass_dm_part,
                    2 if two_species else 1)),
                              ("mb", lambda: halo_gas.most_bound_particle(
                                  gas.x, gas.y, gas.z, gas.mass, dm.x, dm.y, dm.z, dm.mass, stars.x, stars.y, stars.z,
                                  stars.mass, np.arange(len(stars)) + 1000, lim, mass_dm_part))):
                CALLS.clear()
                CALL_INPUTS.clear()
                np.random.seed(4242)
                run()
                pre = "" if part == "rps" else "mb_"
                assert len(CALLS) == int(g[pre + "n_calls"]), (tag, part)
                d["%s_%s_n_calls" % (tag, part)] = np.int64(len(CALLS))
                for k, ((nt, ns, out), ins) in enumerate(zip(CALLS, CALL_INPUTS)):
                    assert np.array_equal(out.view(np.uint32), g["%scall%d_be" % (pre, k)].view(np.uint32)), (tag, part, k)
                    for name, a in zip(("tm", "tx", "ty", "tz", "sx", "sy", "sz"), ins):
                        assert a.dtype == np.float32, (tag, part, k, name, a.dtype)
                        d["%s_%s_call%d_%s" % (tag, part, k, name)] = a
                    d["%s_%s_call%d_be" % (tag, part, k)] = out
    finally:
        RECORD_INPUTS = False
        CALLS.clear()
        CALL_INPUTS.clear()
    np.savez_compressed(os.path.join(HERE, "call_sequences.npz"), **d)
    print("call_sequences:", {k: int(v) for k, v in d.items() if k.endswith("n_calls")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "calls":
        # only the call sequences, checked against the existing fixtures (which stay as they are)
        install_stubs()
        call_sequences()
        sys.exit(0)
    main()
    shape_and_sigma()
    gather_amr()
    stellar_amr()
    call_sequences()
