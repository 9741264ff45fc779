"""GPU parity of the potential kernel against the CPU oracle, through the C-ABI.

EXACT mode: bit-identical to the oracle's f32seq (the reference arithmetic).
FAST mode: within 1e-6 relative of the oracle's f64acc (tolerance from BASELINE.json's
north_star: "potentials ... must match within 1e-6 relative").
"""
import numpy as np
import pytest

from oracle import oracle as O
from pyhalma_b200 import halo_gas, particle, synth

pytestmark = pytest.mark.gpu
f32 = np.float32
FAST_RTOL = 1e-6


def gpu(m, x, y, z, tx, ty, tz, mode):
    return particle.brute_force_binding_energy(1, len(m), f32(m), f32(x), f32(y), f32(z), len(tx), f32(tx),
                                               f32(ty), f32(tz), mode=mode)


def ora(m, x, y, z, tx, ty, tz, variant):
    return O.brute_force_binding_energy(0, len(m), f32(m), f32(x), f32(y), f32(z), len(tx), f32(tx), f32(ty),
                                        f32(tz), variant=variant)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_known_answers(mode):
    out = gpu([26.0], [3.0], [4.0], [12.0], [0.0], [0.0], [0.0], mode)
    assert out.dtype == np.float32 and abs(out[0] - 2.0) <= (0 if mode == "exact" else 4e-7)
    m, x, y, z = [5.0, 7.0, 11.0], [1.0, 1.0, 2.0], [0.0, 3.0, 4.0], [0.0, 4.0, 12.0]
    out = gpu(m, x, y, z, x, y, z, mode)
    ref = ora(m, x, y, z, x, y, z, "f32seq")
    if mode == "exact":
        assert np.array_equal(bits(out), bits(ref))
    else:
        np.testing.assert_allclose(out, ref, rtol=FAST_RTOL)
    m, x, y, z = [1.0, 1.0, 4.0], [0.5, 0.5, 1.5], [0.25, 0.25, 2.25], [0.125, 0.125, 2.125]
    out = gpu(m, x, y, z, x, y, z, mode)
    assert np.all(np.isfinite(out))
    np.testing.assert_allclose(out, [4 / 3, 4 / 3, 2 / 3], rtol=FAST_RTOL)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_lattice_exclusion(mode):
    n = 8
    g = (np.arange(n) + 0.5) * 0.25
    X, Y, Z = (a.ravel() for a in np.meshgrid(g, g, g, indexing="ij"))
    m = np.ones(n ** 3)
    out = gpu(m, X, Y, Z, X, Y, Z, mode)
    if mode == "exact":
        assert np.array_equal(bits(out), bits(ora(m, X, Y, Z, X, Y, Z, "f32seq")))
    else:
        np.testing.assert_allclose(out, ora(m, X, Y, Z, X, Y, Z, "f64acc"), rtol=FAST_RTOL)


RAGGED = [(1, 1), (2, 1), (3, 5), (4, 31), (5, 32), (7, 33), (127, 127), (128, 128), (129, 129), (131, 257),
          (255, 1), (256, 640), (257, 1000), (383, 77), (1000, 513), (4099, 300)]


@pytest.mark.parametrize("n_src,n_tgt", RAGGED)
def test_ragged_sizes_exact_and_fast(n_src, n_tgt):
    rng = np.random.default_rng(n_src * 7919 + n_tgt)
    p = synth.plummer_stars(n_src, 2e-3, 1e6, rng)
    t = synth.plummer_stars(n_tgt, 2e-3, 1e6, rng)
    k = min(n_src, n_tgt) // 2
    t.x[:k], t.y[:k], t.z[:k] = p.x[:k], p.y[:k], p.z[:k]       # some targets are sources too
    args = (p.mass, p.x, p.y, p.z, t.x, t.y, t.z)
    assert np.array_equal(bits(gpu(*args, "exact")), bits(ora(*args, "f32seq")))
    ref = ora(*args, "f64acc")
    got = gpu(*args, "fast").astype(np.float64)
    ok = ref > 0
    assert np.all(got[~ok] == 0)
    assert np.max(np.abs(got[ok] / ref[ok] - 1)) < FAST_RTOL


def test_cfg1_stars_and_gas_exact_bit_parity():
    c = synth.config1()
    s, g, d = c.stars, c.gas, c.dm
    src = [np.concatenate((getattr(g, k), getattr(s, k), getattr(d, k))) for k in ("mass", "x", "y", "z")]
    out = gpu(*src, s.x, s.y, s.z, "exact")
    ref = ora(*src, s.x, s.y, s.z, "f32seq")
    assert np.array_equal(bits(out), bits(ref))
    out = gpu(g.mass, g.x, g.y, g.z, g.x, g.y, g.z, "exact")        # lattice gas: ~13 % pairs excluded
    ref = ora(g.mass, g.x, g.y, g.z, g.x, g.y, g.z, "f32seq")
    assert np.array_equal(bits(out), bits(ref))


def test_cfg1_fast_within_tolerance():
    c = synth.config1()
    s, g, d = c.stars, c.gas, c.dm
    src = [np.concatenate((getattr(g, k), getattr(s, k), getattr(d, k))) for k in ("mass", "x", "y", "z")]
    for tgt in (s, g):
        got = gpu(*src, tgt.x, tgt.y, tgt.z, "fast").astype(np.float64)
        ref = ora(*src, tgt.x, tgt.y, tgt.z, "f64acc")
        err = np.abs(got / ref - 1)
        assert err.max() < FAST_RTOL, err.max()
        # and the drift of the reference's own float32 sum against the same anchor, for the record
        drift = np.abs(ora(*src, tgt.x, tgt.y, tgt.z, "f32seq") / ref - 1).max()
        assert drift < 1e-4


def test_cfg2_full_size_fast_sampled_check_and_linearity():
    # BASELINE cfg2 sizes: 2e5 star targets against 7e5 sources; the oracle checks a sample
    c = synth.config2()
    s, g = c.stars, c.gas
    src = [np.concatenate((getattr(g, k), getattr(s, k))) for k in ("mass", "x", "y", "z")]
    got = gpu(*src, s.x, s.y, s.z, "fast")
    pick = np.random.default_rng(1).choice(len(s), 256, replace=False)
    ref = ora(*src, s.x[pick], s.y[pick], s.z[pick], "f64acc")
    assert np.max(np.abs(got[pick].astype(np.float64) / ref - 1)) < FAST_RTOL
    # linearity: scaling every mass by 4 scales Phi by exactly 4 (power of two); the same call shape, because
    # a call this large runs as a plan (the stars are a block of the sources) whose tiling depends on the shape
    got4 = gpu(src[0] * 4, src[1], src[2], src[3], s.x, s.y, s.z, "fast")
    assert np.array_equal(bits(got4), bits(got * f32(4)))
    # ... and on the direct path (a call below the plan threshold)
    few = gpu(src[0], src[1], src[2], src[3], s.x[:10_000], s.y[:10_000], s.z[:10_000], "fast")
    few4 = gpu(src[0] * 4, src[1], src[2], src[3], s.x[:10_000], s.y[:10_000], s.z[:10_000], "fast")
    assert np.array_equal(bits(few4), bits(few * f32(4)))
    np.testing.assert_allclose(few, got[:10_000], rtol=FAST_RTOL)
    # additivity over a source partition, in float64 tolerance
    a = gpu(g.mass, g.x, g.y, g.z, s.x[:20_000], s.y[:20_000], s.z[:20_000], "fast").astype(np.float64)
    b = gpu(s.mass, s.x, s.y, s.z, s.x[:20_000], s.y[:20_000], s.z[:20_000], "fast").astype(np.float64)
    np.testing.assert_allclose(a + b, got[:20_000], rtol=5e-7)


def test_empty_inputs():
    out = particle.brute_force_binding_energy(1, 0, f32([]), f32([]), f32([]), f32([]), 3, f32([1, 2, 3]),
                                              f32([1, 2, 3]), f32([1, 2, 3]))
    assert out.shape == (3,) and np.all(out == 0)
    out = particle.brute_force_binding_energy(1, 2, f32([1, 1]), f32([1, 2]), f32([1, 2]), f32([1, 2]), 0, f32([]),
                                              f32([]), f32([]))
    assert out.shape == (0,)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_wrapper_mirror_float64_inputs(mode):
    rng = np.random.default_rng(3)
    p = synth.plummer_stars(700, 2e-3, 1e6, rng)
    got = halo_gas.brute_force_binding_energy_fortran(p.mass, p.x, p.y, p.z, p.x[:99], p.y[:99], p.z[:99], mode=mode)
    ref = O.brute_force_binding_energy_fortran(p.mass, p.x, p.y, p.z, p.x[:99], p.y[:99], p.z[:99],
                                               variant="f32seq" if mode == "exact" else "f64acc")
    assert got.dtype == np.float32
    if mode == "exact":
        assert np.array_equal(bits(got), bits(ref))
    else:
        np.testing.assert_allclose(got, ref, rtol=FAST_RTOL)
    ser = halo_gas.serial_brute_force_binding_energy_fortran(p.mass, p.x, p.y, p.z, p.x[:99], p.y[:99], p.z[:99],
                                                             mode=mode)
    assert np.array_equal(bits(ser), bits(got))


def test_nan_inf_semantics_exact():
    out = gpu([1.0, 1.0], [np.nan, 5.0], [1.0, 0.0], [1.0, 7.0], [0.0], [0.0], [0.0], "exact")
    assert np.isnan(out[0])
    out = gpu([1.0, 1.0], [np.nan, 5.0], [0.0, 3.0], [1.0, 7.0], [0.0], [0.0], [0.0], "exact")
    assert out[0] == f32(1.0) / np.sqrt(f32(83.0))
    out = gpu([1.0], [np.inf], [1.0], [1.0], [0.0], [0.0], [0.0], "exact")
    assert out[0] == 0.0


def test_exact_branch_free_arithmetic_selftest():
    """EXACT mode's branch-free rn(m / rn(sqrt(r2))) against the IEEE library routines on the device:
    every float32 mantissa x exponent parity of r2, and 6e10 pseudo-random operand pairs over the
    safe window (biased to mantissas next to powers of two and to zero masses)."""
    import ctypes as C
    from pyhalma_b200 import _lib
    for seed in (1, 20240215, 987654321, 55):
        n = C.c_int64(-1)
        _lib.check(_lib.lib().halma_selftest_exact_arith(0, 15_000_000_000, seed, C.byref(n)))
        assert n.value == 0


def test_exact_outside_the_safe_window_falls_back_bit_exactly():
    """Masses and separations outside the window of the branch-free sequence (tiny, huge, subnormal
    r^2, zero and negative masses) take the IEEE path quad by quad: still bit-identical to the oracle."""
    rng = np.random.default_rng(77)
    n = 700
    x = 3.0 + rng.normal(0, 1e-3, n)
    y = -7.0 + rng.normal(0, 1e-3, n)
    z = 11.0 + rng.normal(0, 1e-3, n)
    m = rng.uniform(0.5e6, 2e6, n)
    m[::7] = 1e-30
    m[3::11] = 1e30
    m[5::13] = 0.0
    m[6::17] = -1e6
    m[9::19] = 1e-42          # subnormal mass
    # sources at subnormal and huge separations from the targets at the origin
    xs = np.concatenate((x, [1e-22, -3e-21, 1e19, 2e-19]))
    ys = np.concatenate((y, [2e-22, 1e-21, -1e19, 1e-19]))
    zs = np.concatenate((z, [-1e-22, 2e-21, 1e19, 3e-19]))
    ms = np.concatenate((m, [1.0, 2.0, 1e6, 5.0]))
    tx = np.concatenate((x[:300], [0.0, 0.0]))
    ty = np.concatenate((y[:300], [0.0, 0.0]))
    tz = np.concatenate((z[:300], [0.0, 0.0]))
    out = gpu(ms, xs, ys, zs, tx, ty, tz, "exact")
    ref = ora(ms, xs, ys, zs, tx, ty, tz, "f32seq")
    assert np.array_equal(bits(out), bits(ref))
    assert np.isinf(out[-1]) or out[-1] > 1e20          # the subnormal separations really were exercised


def test_large_calls_through_a_plan(monkeypatch):
    """Large FAST calls of halma_potential_f32 whose targets are a block of the sources run as a one-pass plan
    (from HALMA_POT_PLAN_MIN_PAIRS pairs on, default 1e10): predicate-free kernel + correction tickets + symmetric
    self-term.  Every shape the reference's callers have, against the direct (predicated) path and the float64
    oracle.  Calls whose targets are other particles (DM -> gas, stars -> gas) run as a plan too, with the targets
    as massless members that are not sources; targets that coincide with sources then go through the fallback."""
    rng = np.random.default_rng(77)
    st = synth.plummer_stars(50_000, 6 * synth.KPC, 1e6, rng)
    synth.add_coincident_pairs(st, 5, rng)                      # shared coordinates -> correction tickets
    gas = synth.lattice_gas(30_000, synth.CELL, rng, m_total=5e9)
    dm = synth.dm_cloud(4_000, 20 * synth.KPC, 1e7, rng)
    cat = lambda k, *ps: np.concatenate([getattr(p, k) for p in ps])       # noqa: E731
    all_src = [cat(k, gas, st, dm) for k in ("mass", "x", "y", "z")]
    cases = {
        # targets ARE the sources (gas-gas of RPS on the lattice; stars-stars of most_bound_particle)
        "self lattice": ([gas.mass, gas.x, gas.y, gas.z], gas),
        "self stars": ([st.mass, st.x, st.y, st.z], st),
        # targets are a block of the sources (escape_velocity_unbinding_fortran: concat(gas, stars, DM) -> stars)
        "block": (all_src, st),
        # block at the very start / end of the sources
        "block first": (all_src, gas),
        "block last": (all_src, dm),
        # targets are other particles (stars -> gas; lattice gas -> stars, whole planes of sources share a coordinate)
        "cross": ([st.mass, st.x, st.y, st.z], gas),
        "cross lattice sources": ([gas.mass, gas.x, gas.y, gas.z], st),
        # ... some of which coincide with sources (a sampled subset of the gas as sources, halo_gas.py:307-321)
        "cross coincident": ([a[len(gas) // 6:5 * len(gas) // 6] for a in (gas.mass, gas.x, gas.y, gas.z)], gas),
    }
    monkeypatch.setenv("HALMA_NP_MIN_PAIRS", "0")               # plans of any size on the predicate-free kernel
    for name, (src, tgt) in cases.items():
        monkeypatch.setenv("HALMA_POT_PLAN_MIN_PAIRS", "0")
        direct = gpu(*src, tgt.x, tgt.y, tgt.z, "fast")
        monkeypatch.setenv("HALMA_POT_PLAN_MIN_PAIRS", "1")
        plan = gpu(*src, tgt.x, tgt.y, tgt.z, "fast")
        assert plan.dtype == np.float32 and plan.shape == direct.shape, name
        np.testing.assert_allclose(plan, direct, rtol=FAST_RTOL, err_msg=name)
        pick = rng.choice(len(tgt), 300, replace=False)
        ref = ora(*src, tgt.x[pick], tgt.y[pick], tgt.z[pick], "f64acc")
        assert np.max(np.abs(plan[pick].astype(np.float64) / ref - 1)) < FAST_RTOL, name
        again = gpu(*src, tgt.x, tgt.y, tgt.z, "fast")
        assert np.array_equal(bits(again), bits(plan)), name           # bit-reproducible
        if name.startswith("cross"):
            # HALMA_POT_PLAN_CROSS=0 keeps cross calls on the direct path
            monkeypatch.setenv("HALMA_POT_PLAN_CROSS", "0")
            assert np.array_equal(bits(gpu(*src, tgt.x, tgt.y, tgt.z, "fast")), bits(direct)), name
            monkeypatch.delenv("HALMA_POT_PLAN_CROSS")
        # EXACT mode never takes the plan path
        if name == "block last":
            assert np.array_equal(bits(gpu(*src, tgt.x, tgt.y, tgt.z, "exact")), bits(ora(*src, tgt.x, tgt.y, tgt.z, "f32seq")))
    monkeypatch.delenv("HALMA_POT_PLAN_MIN_PAIRS")
