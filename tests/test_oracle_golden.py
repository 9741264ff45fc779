"""The oracle's Python restatements against fixtures written by the reference's own
Python drivers (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

RPS_CASES = ["rps_one_dm", "rps_two_dm", "rps_sampled"]


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


@pytest.mark.parametrize("name", RPS_CASES)
def test_rps_matches_reference_python(golden_dir, name):
    g = load(golden_dir, name)
    np.random.seed(int(g["np_seed"]))
    r = O.RPS(g["gas_x"], g["gas_y"], g["gas_z"], g["gas_vx"], g["gas_vy"], g["gas_vz"], g["gas_mass"],
              g["gas_temp"], g["dm_x"], g["dm_y"], g["dm_z"], g["dm_mass"], g["st_x"], g["st_y"],
              g["st_z"], g["st_mass"], *g["vb"], int(g["lim"]), float(g["mass_dm_part"]),
              int(g["num_dm_species"]))
    np.testing.assert_array_equal(np.array(r.as_tuple()), g["rps_out"])
    # class order and sizes of the kernel calls (halo_gas.py:306-450)
    n_calls = int(g["n_calls"])
    assert n_calls == (4 if int(g["num_dm_species"]) > 1 else 3)
    assert all(int(g["call%d_ntest" % k]) == len(g["gas_x"]) for k in range(n_calls))
    if int(g["lim"]) >= 5000:
        total = np.zeros(len(g["gas_x"]), np.float32)
        for k in range(n_calls):
            total += g["call%d_be" % k]
        np.testing.assert_array_equal(total, r.be32)
        assert int(g["call0_ntotal"]) == len(g["gas_x"])
        assert int(g["call%d_ntotal" % (n_calls - 1)]) == len(g["st_x"])


@pytest.mark.parametrize("name", RPS_CASES)
def test_most_bound_matches_reference_python(golden_dir, name):
    g = load(golden_dir, name)
    np.random.seed(int(g["np_seed"]))
    oripa = np.arange(len(g["st_x"])) + 1000
    out = O.most_bound_particle(g["gas_x"], g["gas_y"], g["gas_z"], g["gas_mass"], g["dm_x"], g["dm_y"],
                                g["dm_z"], g["dm_mass"], g["st_x"], g["st_y"], g["st_z"], g["st_mass"],
                                oripa, int(g["lim"]), float(g["mass_dm_part"]))
    np.testing.assert_array_equal(np.array(out, dtype=np.float64), g["mb_out"])


def test_stellar_onepass_matches_reference_python(golden_dir):
    g = load(golden_dir, "stellar_onepass")
    bound, be32, E = O.escape_velocity_unbinding(
        (g["gas_x"], g["gas_y"], g["gas_z"], g["gas_mass_seen"]),
        (g["st_x"], g["st_y"], g["st_z"], g["st_vx"], g["st_vy"], g["st_vz"], g["st_mass"]),
        (g["dm_x"], g["dm_y"], g["dm_z"], g["dm_mass"]), g["vb"], float(g["factor_v"]))
    np.testing.assert_array_equal(bound, g["bound"])
    assert 0 < bound.sum() < len(bound)
    assert int(g["n_calls"]) == 1
    np.testing.assert_array_equal(be32, g["call0_be"])
    assert int(g["call0_ntotal"]) == len(g["gas_x"]) + len(g["st_x"]) + len(g["dm_x"])
    # one pass of the fixed-point loop with the given bulk velocity is the same function
    r = O.unbind_halo(g["st_x"], g["st_y"], g["st_z"], g["st_vx"], g["st_vy"], g["st_vz"], g["st_mass"],
                      pre=[(g["gas_mass_seen"], g["gas_x"], g["gas_y"], g["gas_z"])],
                      post=[(g["dm_mass"], g["dm_x"], g["dm_y"], g["dm_z"])],
                      kappa=float(g["factor_v"]) ** 2, vb_fixed=g["vb"], max_iter=1)
    np.testing.assert_array_equal(r.mask, g["bound"])


def test_com_and_bulk_velocity_match_reference_numba(golden_dir):
    g = load(golden_dir, "rps_one_dm")
    idx = np.arange(len(g["st_x"]))
    cx, cy, cz, M = O.center_of_mass(idx, g["st_x"], g["st_y"], g["st_z"], g["st_mass"])
    vb = O.CM_velocity(M, idx, g["st_vx"], g["st_vy"], g["st_vz"], g["st_mass"])
    # numba parallel+fastmath reductions are order-nondeterministic: 1e-13, not bit-exact
    np.testing.assert_allclose([cx, cy, cz], g["com"], rtol=1e-13)
    np.testing.assert_allclose(vb, g["vb"], rtol=1e-12)
    np.testing.assert_allclose(M, g["M"], rtol=1e-14)


def test_call_sequences_fixture_is_consistent_with_the_driver_fixtures(golden_dir):
    """call_sequences.npz (kernel calls of the reference's drivers WITH their float32 inputs) against the oracle
    kernel and against the per-driver fixtures written by the same reference run."""
    g = np.load(os.path.join(golden_dir, "call_sequences.npz"))
    n = 0
    for tag in ("rps_one_dm", "rps_two_dm", "rps_sampled"):
        drv = np.load(os.path.join(golden_dir, tag + ".npz"))
        for part, pre in (("rps", ""), ("mb", "mb_")):
            assert int(g["%s_%s_n_calls" % (tag, part)]) == int(drv[pre + "n_calls"])
            for k in range(int(drv[pre + "n_calls"])):
                a = [g["%s_%s_call%d_%s" % (tag, part, k, nm)] for nm in ("tm", "tx", "ty", "tz", "sx", "sy", "sz")]
                be = g["%s_%s_call%d_be" % (tag, part, k)]
                assert all(x.dtype == np.float32 for x in a) and len(a[0]) == int(drv["%scall%d_ntotal" % (pre, k)])
                assert np.array_equal(be.view(np.uint32), drv["%scall%d_be" % (pre, k)].view(np.uint32))
                out = O.brute_force_binding_energy(1, len(a[0]), *a[:4], len(a[4]), *a[4:])
                assert np.array_equal(out.view(np.uint32), be.view(np.uint32))
                n += 1
    assert n == 22
