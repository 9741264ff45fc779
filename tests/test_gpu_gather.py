"""GPU parity of the gather step through the C-ABI: bit-exact against the
fixture written by the reference's own st_gas_dm_particles_inside and against the oracle."""
import os

import numpy as np
import pytest

from oracle import gather as OG
from pyhalma_b200 import gather, halo_gas, synth
from test_oracle_gather import SNAP_KW, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def snap():
    return synth.amr_snapshot(**SNAP_KW)


def call(s, cx, cy, cz, R, fn=halo_gas.st_gas_dm_particles_inside, **kw):
    return fn(s.rete, s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data, s.masclet_st_data, None, None,
              cx, cy, cz, R, s.rho_B, **kw)


def test_gather_golden_bit_exact(snap, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "gather_amr.npz")))
    for tag in g["tags"]:
        cx, cy, cz, R = g[tag + "_args"]
        out = call(snap, cx, cy, cz, R, mass_to_sun=1.0)
        assert len(out) == 17
        check_against_golden(out, g, tag)
        # DM and stars ascending in the input index
        assert np.all(np.diff(out[16]) > 0)


def test_gather_random_queries_against_oracle(snap):
    rng = np.random.default_rng(3)
    for _ in range(12):
        c = np.asarray(snap.centre) + rng.normal(0, 0.01, 3)
        R = float(10 ** rng.uniform(-3.3, -0.8))
        got = call(snap, *c, R, mass_to_sun=1.0)
        want = call(snap, *c, R, fn=OG.st_gas_dm_particles_inside)
        for n, a, b in zip(OG.GATHER_NAMES, got, want):
            assert a.dtype == b.dtype, n
            np.testing.assert_array_equal(a, b, err_msg=n)      # same order too: both ascending
    # mass_to_sun scales DM and star masses only (halo_gas.py:252,265)
    a = call(snap, *snap.centre, 0.02, mass_to_sun=1.0)
    b = call(snap, *snap.centre, 0.02, mass_to_sun=2.0)
    np.testing.assert_array_equal(b[11], 2.0 * a[11])
    np.testing.assert_array_equal(b[15], 2.0 * a[15])
    np.testing.assert_array_equal(b[6], a[6])
    gather.release_cached_snapshot()


def test_amrgrid_to_particles_box_only(snap):
    cx, cy, cz = snap.centre
    for R in (0.004, 0.03):
        got = halo_gas.AMRgrid_to_particles(snap.L, snap.ncoarse, snap.grid_data, snap.gas_data, R, cx, cy, cz, snap.rho_B)
        want = OG.AMRgrid_to_particles(snap.L, snap.ncoarse, snap.grid_data, snap.gas_data, R, cx, cy, cz, snap.rho_B)
        assert len(want[0]) > 0
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
    gather.release_cached_snapshot()


def test_snapshot_object_edge_cases(snap):
    with gather.Snapshot(snap.L, snap.ncoarse, snap.grid_data, snap.gas_data) as s:     # no particles uploaded
        assert s.n_cells == snap.n_cells
        out = s.gather(*snap.centre, 0.02, snap.rho_B, snap.rete)
        assert len(out[0]) > 0 and len(out[8]) == 0 and len(out[12]) == 0
        far = s.gather(100.0, 100.0, 100.0, 0.02, snap.rho_B, snap.rete)                # nothing there
        assert all(len(a) == 0 for a in far)
        zero = s.gather(*snap.centre, 0.0, snap.rho_B, snap.rete)
        assert all(len(a) == 0 for a in zero)
        s.upload_particles(1, [snap.centre[0]], [snap.centre[1]], [snap.centre[2]], [5.0], [42])
        one = s.gather(*snap.centre, 0.0, snap.rho_B, snap.rete)                        # distance 0 <= R = 0
        assert list(one[16]) == [42] and one[15][0] == 5.0
        with pytest.raises(ValueError):
            s.upload_particles(0, [0.0, 1.0], [0.0], [0.0], [1.0])
    # a grid whose fields disagree with the declared extents is refused
    bad = [list(a) for a in snap.gas_data]
    bad[0][1] = bad[0][1][:-1]
    with pytest.raises(ValueError):
        gather.Snapshot(snap.L, snap.ncoarse, snap.grid_data, bad)
    with pytest.raises(ImportError):
        call(snap, *snap.centre, 0.01)            # masclet_framework absent and no mass_to_sun given


def test_gather_feeds_rps_like_the_reference_pipeline(snap):
    """pyHALMA.py:1027-1054: gather, then RPS on the gathered arrays; the same through the oracle."""
    from oracle import oracle as O
    cx, cy, cz = snap.centre
    R = 0.012
    got = call(snap, cx, cy, cz, R, mass_to_sun=1.0)
    want = call(snap, cx, cy, cz, R, fn=OG.st_gas_dm_particles_inside)
    vb = synth.BULK_V
    a = halo_gas.RPS(*got[:16], *vb, 10 ** 9, 8e7, 1, mode="exact")
    b = O.RPS(*want[:16], *vb, 10 ** 9, 8e7, 1)
    np.testing.assert_array_equal(np.array(a), np.array(b.as_tuple()))
    gather.release_cached_snapshot()


@pytest.mark.parametrize("num_dm_species", [1, 2])
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_device_resident_potential_stage(snap, num_dm_species, mode):
    """pyHALMA.py:1023-1067 without host arrays: gather, RPS and most_bound_particle read the
    gathered particles in HBM; only scalars return.  Checked against the oracle's RPS /
    most_bound_particle on the oracle's gather."""
    from oracle import oracle as O
    from pyhalma_b200 import pipeline
    mass_dm_part = 8e7                       # synth DM particles weigh 1e7: light unless the limit is lowered
    s2 = synth.amr_snapshot(**SNAP_KW)
    dm_m = s2.masclet_dm_data[3].copy()
    dm_m[::3] *= 8.0                         # a heavy species: mass 8e7 >= 0.9 * 8e7 / 8
    s2.masclet_dm_data[3] = dm_m
    vb = synth.BULK_V
    with gather.Snapshot(s2.L, s2.ncoarse, s2.grid_data, s2.gas_data, s2.masclet_dm_data, s2.masclet_st_data) as dev:
        for R, off in ((0.012, 0.0), (0.03, 0.004)):
            cx, cy, cz = s2.centre[0] + off, s2.centre[1] - off, s2.centre[2]
            rps, mb = pipeline.halo_potential_stage(dev, cx, cy, cz, R, s2.rho_B, s2.rete, *vb, mass_dm_part,
                                                    num_dm_species, mode=mode)
            w = call(s2, cx, cy, cz, R, fn=OG.st_gas_dm_particles_inside)
            variant = "f32seq" if mode == "exact" else "f64acc"
            want_rps = O.RPS(*w[:16], *vb, 10 ** 9, mass_dm_part, num_dm_species, variant=variant).as_tuple()
            # ordered float64 sums on the device vs numpy's pairwise sums: 1e-13, not bitwise;
            # fast mode may move a cell whose energy is within 1e-6 of zero across the boundary
            tol = 1e-13 if mode == "exact" else 2e-3
            np.testing.assert_allclose(rps, want_rps, rtol=tol)
            want_mb = O.most_bound_particle(w[0], w[1], w[2], w[6], *w[8:12], *w[12:16], w[16], 10 ** 9, mass_dm_part,
                                            variant=variant)
            assert mb[3] == want_mb[3] and (mb[0], mb[1], mb[2]) == tuple(want_mb[:3])
        # nothing inside: RPS of an empty gas set is all zeros (halo_gas.py:483-490), no star -> None
        rps, mb = pipeline.halo_potential_stage(dev, 100.0, 100.0, 100.0, 0.01, s2.rho_B, s2.rete, *vb, mass_dm_part,
                                                num_dm_species, mode=mode)
        assert rps == (0.0, 0.0, 0.0, 0.0) and mb is None
        g = dev.gather_device(*s2.centre, 0.02, s2.rho_B, s2.rete, dm_heavy_min=0.9 * mass_dm_part / 8)
        full = dev.gather(*s2.centre, 0.02, s2.rho_B, s2.rete)
        assert g.n_dm + g.n_dm_light == len(full[8]) and g.n_dm == int(np.sum(full[11] >= 0.9 * mass_dm_part / 8))
        assert g.n_gas == len(full[0]) and g.n_st == len(full[12])


def test_stellar_unbinding_end_to_end_golden(snap, golden_dir):
    """halo_properties.escape_velocity_unbinding_fortran with the reference's signature: AMR gather
    on the GPU + EXACT kernel reproduce the mask the reference computed end to end."""
    from pyhalma_b200 import halo_properties
    from test_oracle_gather import stellar_amr_inputs
    g = dict(np.load(os.path.join(golden_dir, "stellar_amr.npz")))
    a = stellar_amr_inputs(snap, g)
    bound = halo_properties.escape_velocity_unbinding_fortran(*a, mode="exact", mass_to_sun=1.0)
    np.testing.assert_array_equal(bound.astype(bool), g["bound"])
    fast = halo_properties.escape_velocity_unbinding_fortran(*a, mode="fast", mass_to_sun=1.0).astype(bool)
    assert np.count_nonzero(fast != g["bound"]) <= 2          # only energies within 1e-6 of zero may flip
    # the DM masses are multiplied by masclet_framework.units.mass_to_sun (halo_properties.py:293): without that
    # package the factor has to be given, and it must reach the DM masses (code units / 7 with factor 7 = the same)
    with pytest.raises(ImportError):
        halo_properties.escape_velocity_unbinding_fortran(*a, mode="exact")
    a7 = list(a)
    dm = a7[5]
    a7[5] = (dm[0], dm[1], dm[2], np.asarray(dm[3]) / 7.0) + tuple(dm[4:])
    b7 = halo_properties.escape_velocity_unbinding_fortran(*a7, mode="exact", mass_to_sun=7.0)
    b1 = halo_properties.escape_velocity_unbinding_fortran(*a7, mode="exact", mass_to_sun=1.0)
    assert np.count_nonzero(b7.astype(bool) != g["bound"]) <= 2 and np.count_nonzero(b1 != b7) > 0
    # AMRgrid_to_particles (no particles) and st_gas_dm_particles_inside (with) alternate per halo in pyHALMA:
    # the resident copy with particles serves both, nothing is uploaded twice
    s_full = gather.snapshot_for(a[1], a[2], a[3], a[4], snap.masclet_dm_data, snap.masclet_st_data, 1.0, 0)
    assert gather.snapshot_for(a[1], a[2], a[3], a[4], None, None, 1.0, 0) is s_full
    assert gather.snapshot_for(a[1], a[2], a[3], a[4], snap.masclet_dm_data, snap.masclet_st_data, 1.0, 0) is s_full
    gather.release_cached_snapshot()


def test_grid_index_equals_brute_force(snap, monkeypatch):
    """Above 2^18 resident particles the ball query goes through a cell list; the result must be the
    brute-force pass's, bit for bit and in the same order.  Forced on here with the small snapshot,
    plus one case at the real threshold."""
    rng = np.random.default_rng(8)
    queries = [(np.asarray(snap.centre) + rng.normal(0, 0.02, 3), float(10 ** rng.uniform(-3.5, -0.5)))
               for _ in range(25)]
    queries += [(np.asarray(snap.centre), 0.0), (np.asarray(snap.centre), 1e3), (np.asarray(snap.centre) + 50.0, 0.01)]

    def run(min_particles, s_):
        monkeypatch.setenv("HALMA_GATHER_INDEX_MIN", str(min_particles))
        with gather.Snapshot(s_.L, s_.ncoarse, s_.grid_data, s_.gas_data, s_.masclet_dm_data, s_.masclet_st_data) as dev:
            out = [dev.gather(*c, R, s_.rho_B, s_.rete) for c, R in queries]
            g = dev.gather_device(*queries[0][0], 0.03, s_.rho_B, s_.rete, dm_heavy_min=1e7)
            out.append((g.n_dm, g.n_dm_light, g.n_st))
        return out

    brute, indexed = run(-1, snap), run(1, snap)
    assert any(len(o[8]) > 100 for o in brute[:-1]) and any(len(o[12]) > 1000 for o in brute[:-1])
    for a, b in zip(brute[:-1], indexed[:-1]):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    assert brute[-1] == indexed[-1]
    # coincident particles, a particle with NaN position and one far outlier must not break the grid
    s2 = synth.amr_snapshot(**SNAP_KW)
    dm = [np.array(a, dtype=np.float64) for a in s2.masclet_dm_data]
    dm[0][:50] = dm[0][50]; dm[1][:50] = dm[1][50]; dm[2][:50] = dm[2][50]
    dm[0][100] = np.nan
    dm[1][101] = 1e6
    s2.masclet_dm_data = dm
    b2, i2 = run(-1, s2), run(1, s2)
    for a, b in zip(b2[:-1], i2[:-1]):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    monkeypatch.delenv("HALMA_GATHER_INDEX_MIN")
    big = synth.amr_snapshot(n_levels=3, n_dm=400_000, n_st=300_000)         # default threshold: indexed
    with gather.Snapshot(big.L, big.ncoarse, big.grid_data, big.gas_data, big.masclet_dm_data, big.masclet_st_data) as dev:
        got = dev.gather(*big.centre, 0.02, big.rho_B, big.rete)
    want = OG.st_gas_dm_particles_inside(big.rete, big.L, big.ncoarse, big.grid_data, big.gas_data, big.masclet_dm_data,
                                         big.masclet_st_data, None, None, *big.centre, 0.02, big.rho_B)
    assert len(got[8]) > 1000 and len(got[12]) > 1000
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
