"""The oracle's restatement of the gather step (halo_gas.py:9-141, 216-277)
against the fixture written by the reference's own st_gas_dm_particles_inside (real numba
patch_to_particles, real scipy KD-trees): tests/golden/make_golden.py::gather_amr."""
import os

import numpy as np
import pytest

from oracle import gather as OG
from pyhalma_b200 import synth

SNAP_KW = dict(n_levels=7, n_dm=20_000, n_st=30_000)      # = make_golden.GATHER_SNAPSHOT


@pytest.fixture(scope="module")
def snap():
    return synth.amr_snapshot(**SNAP_KW)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "gather_amr.npz")))


def check_against_golden(out, g, tag):
    """Bit-exact: gas in the reference's order, DM / stars as sets (canonical order)."""
    out = OG.canonical_gather(out)
    assert [len(out[0]), len(out[8]), len(out[12])] == list(g[tag + "_counts"])
    for n, a in zip(OG.GATHER_NAMES, out):
        assert a.dtype == (np.int64 if n == "st_oripa" else np.float64), (n, a.dtype)
        assert np.array_equal(OG.digest(a), g["%s_sha_%s" % (tag, n)]), (tag, n)
        if "%s_%s" % (tag, n) in g:
            np.testing.assert_array_equal(a, g["%s_%s" % (tag, n)])


def test_gather_matches_reference(snap, golden):
    assert snap.n_cells == int(golden["n_cells"])
    assert len(golden["tags"]) == 4
    for tag in golden["tags"]:
        cx, cy, cz, R = golden[tag + "_args"]
        out = OG.st_gas_dm_particles_inside(snap.rete, snap.L, snap.ncoarse, snap.grid_data, snap.gas_data,
                                            snap.masclet_dm_data, snap.masclet_st_data, None, None, cx, cy, cz, R,
                                            snap.rho_B)
        check_against_golden(out, golden, tag)


def test_gather_properties(snap):
    cx, cy, cz = snap.centre
    R = 0.03
    out = OG.st_gas_dm_particles_inside(snap.rete, snap.L, snap.ncoarse, snap.grid_data, snap.gas_data,
                                        snap.masclet_dm_data, snap.masclet_st_data, None, None, cx, cy, cz, R,
                                        snap.rho_B)
    gx, gy, gz, gm = out[0], out[1], out[2], out[6]
    assert len(gx) > 1000
    assert np.all(np.sqrt((gx - cx) ** 2 + (gy - cy) ** 2 + (gz - cz) ** 2) < R)
    assert np.all(gm > 0)
    # cells are lattice points of their level: x - left edge is a half-integer number of cells
    res_min = (snap.L / snap.ncoarse) / 2 ** 7
    k = (gx + snap.L / 2) / res_min
    assert np.allclose(k * 2, np.round(k * 2), atol=1e-6)
    # a larger sphere contains the smaller one's particles (monotone selection)
    big = OG.st_gas_dm_particles_inside(snap.rete, snap.L, snap.ncoarse, snap.grid_data, snap.gas_data,
                                        snap.masclet_dm_data, snap.masclet_st_data, None, None, cx, cy, cz, 2 * R,
                                        snap.rho_B)
    assert set(out[16]).issubset(set(big[16])) and len(big[0]) > len(gx)
    key = lambda o: set(zip(o[0].tolist(), o[1].tolist(), o[2].tolist()))     # noqa: E731
    assert key(out).issubset(key(big))
    # refined / overlapped cells never appear: all-False flags -> no gas
    gd = list(snap.gas_data)
    gd[5] = [np.zeros_like(a) for a in gd[5]]
    none = OG.AMRgrid_to_particles(snap.L, snap.ncoarse, snap.grid_data, gd, R, cx, cy, cz, snap.rho_B)
    assert all(len(a) == 0 for a in none)


def test_helpers():
    assert list(OG.create_vector_levels([0, 2, 1])) == [0, 1, 1, 2]
    # a patch far from the box is dropped, patch 0 is always kept
    w = OG.which_patches_inside_box([0, 1, 0, 1, 0, 1], [4, 4, 4], [4, 4, 4], [4, 4, 4], [0., .5, 30.], [0., .5, 30.],
                                    [0., .5, 30.], [0, 2], 40.0, 128)
    assert w == [0, 1]


def stellar_amr_inputs(snap, g):
    st = snap.masclet_st_data
    return (snap.rete, snap.L, snap.ncoarse, snap.grid_data, snap.gas_data, snap.masclet_dm_data, *g["com"], *g["vb"],
            float(g["Rmax"]), g["part_list"], st[0], st[1], st[2], st[3], st[4], st[5], st[6], float(g["factor_v"]),
            snap.rho_B)


def test_stellar_unbinding_end_to_end_matches_reference(snap, golden_dir):
    """escape_velocity_unbinding_fortran run by the reference itself on the AMR snapshot (its own DM
    mask, AMR gather and energy step) against the oracle's gather + kernel + energy step."""
    from oracle import oracle as O
    g = dict(np.load(os.path.join(golden_dir, "stellar_amr.npz")))
    a = stellar_amr_inputs(snap, g)
    rete, L, ncoarse, grid_data, gas_data, dm_data, cx, cy, cz, vx, vy, vz, Rmax, part_list = a[:14]
    st = a[14:21]
    gx, gy, gz, _, _, _, gm, _ = OG.AMRgrid_to_particles(L, ncoarse, grid_data, gas_data, Rmax, cx, cy, cz, a[22])
    inside = OG.parallel_inside(gx, gy, gz, Rmax, cx, cy, cz)
    gas = (gx[inside], gy[inside], gz[inside], gm[inside] * rete ** 3)
    dins = np.sqrt((dm_data[0] - cx) ** 2 + (dm_data[1] - cy) ** 2 + (dm_data[2] - cz) ** 2) < Rmax
    dm = tuple(np.asarray(c)[dins] for c in dm_data[:4])
    stars = tuple(np.asarray(c)[part_list] for c in st)
    bound, be32, _ = O.escape_velocity_unbinding(gas, stars, dm, (vx, vy, vz), a[21])
    assert int(g["n_calls"]) == 1 and int(g["call0_ntotal"]) == len(gas[0]) + len(part_list) + len(dm[0])
    np.testing.assert_array_equal(be32, g["call0_be"])
    np.testing.assert_array_equal(bound, g["bound"])
    assert 0 < bound.sum() < len(bound)
