"""The reuse bookkeeping of the device loop (external sums cached, incremental passes), restated in
numpy (oracle/reuse_model.py), against the plain loop of the oracle: same member sets, pass counts and
potentials inside the FAST tolerances.  The GPU tests pin the CUDA implementation to this model through
the number of 1/r evaluations it reports."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import reuse_model as RM
from pyhalma_b200 import synth

FAST_RTOL = 1e-6
BAND = 1e-6


def small_case(extra=0, n_star=1500, n_gas=700, n_dm=200):
    c = synth.config1(n_star, n_gas, seed_extra=extra, n_dm=n_dm)
    return c.stars, c.gas, c.dm


@pytest.mark.parametrize("cache,incr", [(True, True), (True, False), (False, True), (False, False)])
def test_stellar_loop_model_matches_plain_oracle(cache, incr):
    s, g, d = small_case(5)
    args = (s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass)
    o = O.unbind_halo(*args, pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0, variant="f64acc")
    m = RM.unbind_halo(*args, ext=[g.pos_mass(), d.pos_mass()], kappa=9.0, cache_external=cache, incremental=incr)
    assert o.n_iter >= 3
    diff = m.mask != o.mask
    assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[diff] < BAND)
    if not diff.any():
        assert m.n_iter == o.n_iter and np.array_equal(m.idx, o.idx)
    seen = o.be32 > 0
    assert np.abs(m.be32[seen].astype(np.float64) / o.be32[seen] - 1).max() < FAST_RTOL
    n_ext = len(g) + len(d)
    plain = sum(n * (n + n_ext) for n in o.n_bound_history[:o.n_iter])
    if not diff.any():
        assert m.evaluations == RM.expected_evaluations(o.n_bound_history, n_ext, cache_external=cache,
                                                        incremental=incr)
    if not cache and not incr:
        assert m.evaluations == plain == o.pairs and set(m.passes) == {"full"}
    else:
        assert m.evaluations < plain
    if incr:
        assert "incr" in m.passes and m.passes[0] == "full"
        # survivors x removed instead of survivors^2: an incremental pass costs at most half a full one
        h = o.n_bound_history
        for k, kind in enumerate(m.passes):
            if kind == "incr":
                assert 2 * (h[k - 1] - h[k]) <= h[k]


def test_gas_lattice_model_with_fixed_bulk_velocity():
    # lattice gas: every pass subtracts the excluded (coordinate-sharing) pairs of the CURRENT sources
    s, g, d = small_case(6, n_star=900, n_gas=1800, n_dm=150)
    vb = O.CM_velocity(O.total_mass(np.arange(len(s)), s.mass), np.arange(len(s)), s.vx, s.vy, s.vz, s.mass)
    args = (g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass)
    o = O.unbind_halo(*args, post=[d.pos_mass(), s.pos_mass()], split_classes=True, kappa=2.0, vb_fixed=vb,
                      variant="f64acc")
    assert O.count_excluded(g.x, g.y, g.z, g.x, g.y, g.z) > len(g)         # the predicate really bites
    m = RM.unbind_halo(*args, ext=[d.pos_mass(), s.pos_mass()], kappa=2.0, vb_fixed=vb)
    diff = m.mask != o.mask
    assert np.all(O.energy_margin(o.energy, o.be32, 2.0)[diff] < BAND)
    if not diff.any():
        assert m.n_iter == o.n_iter
    seen = o.be32 > 0
    assert np.abs(m.be32[seen].astype(np.float64) / o.be32[seen] - 1).max() < FAST_RTOL


def tight_pair_case():
    """A bound member A whose potential is ~94 % due to a neighbour B at 7e-8 Mpc that leaves in the first pass."""
    rng = np.random.default_rng(51)
    p = synth.plummer_stars(3000, 2e-3, 1e6, rng, centre=(0.0, 0.0, 0.0), bulk_v=(0., 0., 0.), interloper_frac=0.1)
    o0 = O.unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, kappa=9.0, variant="f64acc")
    a = int(np.flatnonzero(o0.mask)[100])
    b = int(np.flatnonzero(~o0.mask & (o0.be32 > 0))[0])
    x, y, z = p.x.copy(), p.y.copy(), p.z.copy()
    vx, vy, vz = p.vx.copy(), p.vy.copy(), p.vz.copy()
    x[b], y[b], z[b] = x[a] + 3e-8, y[a] + 4e-8, z[a] - 5e-8
    vx[b], vy[b], vz[b] = 3e4, -2e4, 1e4
    return (x, y, z, vx, vy, vz, p.mass), a, b


def test_incremental_pass_needs_the_heavy_removal_guard():
    """Subtracting a removed neighbour that carried most of a member's potential amplifies the float32 rounding
    of that one term beyond the tolerance; the guard (csrc/potential.cu::kIncrHeavy, reuse_model.INCR_HEAVY)
    redoes such a pass with the predicated kernel."""
    args, a, b = tight_pair_case()
    o = O.unbind_halo(*args, kappa=9.0, variant="f64acc")
    assert o.mask[a] and not o.mask[b] and o.n_iter >= 3
    seen = o.be32 > 0
    err = lambda m: np.abs(m.be32[seen].astype(np.float64) / o.be32[seen] - 1).max()      # noqa: E731
    unguarded = RM.unbind_halo(*args, kappa=9.0, cache_external=False, heavy_guard=False)
    assert unguarded.passes[:2] == ["full", "incr"] and err(unguarded) > FAST_RTOL        # what the guard is for
    guarded = RM.unbind_halo(*args, kappa=9.0, cache_external=False)
    assert guarded.passes == ["full", "fallback", "full"]          # a pass that fell back is followed by a full one
    assert err(guarded) < FAST_RTOL and np.array_equal(guarded.mask, o.mask) and guarded.n_iter == o.n_iter
    assert guarded.evaluations > RM.expected_evaluations(o.n_bound_history, 0, cache_external=False, incremental=True)
