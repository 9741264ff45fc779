"""Work the device loop does not repeat (halma_unbind_config.cache_external / .incremental).

The loop of SURVEY.md §3.4 re-evaluates the whole potential every pass.  Two things in it cannot
have changed from one pass to the next: the sum over the fixed external sources, and -- when a
pass removed only a few members -- most of the member x member sum.  With the options on, the
per-pair terms are the same ones, only fewer of them are evaluated; results must agree with the
plain loop and with the float64-accumulating oracle inside the FAST tolerances (1e-6 on
potentials, masks identical outside the 1e-6 energy band), pass for pass.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from oracle import reuse_model as RM
from pyhalma_b200 import synth
from pyhalma_b200.unbind import UnbindPlan, unbind_catalogue, unbind_halo

pytestmark = pytest.mark.gpu
FAST_RTOL = 1e-6
BAND = 1e-6
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [dict(cache_external=True, incremental=False), dict(cache_external=False, incremental=True),
            dict(cache_external=True, incremental=True)]


def check_against(r, ref, o, kappa, n_ext_pairs_saved):
    """r: run with reuse; ref: the plain loop on the GPU; o: oracle f64acc."""
    assert np.count_nonzero(r.mask != ref.mask) <= 2
    if np.array_equal(r.mask, ref.mask):
        assert r.n_iter == ref.n_iter and r.pairs == ref.pairs    # interactions are counted the same way
    if np.array_equal(r.mask, o.mask):
        assert r.n_iter == o.n_iter
    np.testing.assert_allclose(r.be32, ref.be32, rtol=FAST_RTOL)
    both = r.mask & o.mask
    assert np.abs(r.be32[both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL
    diff = r.mask != o.mask
    assert np.all(O.energy_margin(o.energy, o.be32, kappa)[diff] < BAND)
    np.testing.assert_allclose(r.mass, ref.mass, rtol=1e-4)
    np.testing.assert_allclose(r.vb, ref.vb, rtol=1e-4, atol=1e-6)
    assert r.stats.evaluations <= ref.stats.evaluations - n_ext_pairs_saved


@pytest.mark.parametrize("symmetric", [True, False])
def test_stellar_halo_with_externals(symmetric):
    # 4e4 stars against 9e4 gas cells + 1e3 DM: 5.2e9 pairs per pass -> predicate-free path
    c = synth.config1(40000, 90000, seed_extra=21, n_dm=1000)
    s, g, d = c.stars, c.gas, c.dm
    args = (s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass)
    kw = dict(pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0)
    o = O.unbind_halo(*args, variant="f64acc", **kw)
    assert o.n_iter >= 3
    ref = unbind_halo(*args, mode="fast", symmetric=symmetric, cache_external=False, incremental=False, **kw)
    assert ref.stats.evaluations == ref.stats.pairs or symmetric
    n_ext = len(g) + len(d)
    for v in VARIANTS:
        r = unbind_halo(*args, mode="fast", symmetric=symmetric, **v, **kw)
        # cached externals are evaluated once: at least (passes - 1) x (bound members) x n_ext fewer evaluations
        check_against(r, ref, o, 9.0, (o.n_iter - 1) * int(o.mask.sum()) * n_ext if v["cache_external"] else 1)
        if np.array_equal(r.mask, o.mask):
            # exactly the passes the numpy model of the bookkeeping takes (oracle/reuse_model.py)
            assert r.stats.evaluations == RM.expected_evaluations(o.n_bound_history, n_ext, symmetric=symmetric, **v)
        again = unbind_halo(*args, mode="fast", symmetric=symmetric, **v, **kw)
        assert np.array_equal(again.be32.view(np.uint32), r.be32.view(np.uint32))      # bit-reproducible
        assert np.array_equal(again.mask, r.mask) and np.array_equal(again.energy, r.energy)
        assert again.stats.evaluations == r.stats.evaluations


def test_gas_layout_lattice_with_fixed_bulk_velocity():
    # gas cells on a lattice (many excluded pairs -> correction tickets every pass), fixed bulk velocity,
    # members first, two external classes
    c = synth.config1(30000, 60000, seed_extra=22, n_dm=800)
    s, g, d = c.stars, c.gas, c.dm
    vb = O.CM_velocity(O.total_mass(np.arange(len(s)), s.mass), np.arange(len(s)), s.vx, s.vy, s.vz, s.mass)
    args = (g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass)
    kw = dict(post=[d.pos_mass(), s.pos_mass()], split_classes=True, kappa=2.0, vb_fixed=vb)
    o = O.unbind_halo(*args, variant="f64acc", **kw)
    assert o.n_iter >= 2
    ref = unbind_halo(*args, mode="fast", cache_external=False, incremental=False, **kw)
    for v in VARIANTS:
        r = unbind_halo(*args, mode="fast", **v, **kw)
        check_against(r, ref, o, 2.0, 1 if v["cache_external"] else 0)
        if np.array_equal(r.mask, o.mask):
            assert r.stats.evaluations == RM.expected_evaluations(o.n_bound_history, len(d) + len(s), symmetric=True, **v)
        # potentials of the particles removed in earlier passes are the ones of the pass that removed them
        gone = ~o.mask & (o.be32 > 0)
        if gone.any():
            assert np.abs(r.be32[gone].astype(np.float64) / o.be32[gone] - 1).max() < FAST_RTOL


def test_catalogue_mixed_haloes_and_graph_driver(monkeypatch):
    # a ragged catalogue with per-halo external groups: haloes converge in different passes, some take
    # incremental passes while others take full ones, some have no externals at all
    monkeypatch.setenv("HALMA_NP_MIN_PAIRS", "0")
    rng = np.random.default_rng(31)
    sizes = [700, 0, 129, 4000, 1, 2600, 128, 900, 5000, 33]
    gsz = [[300, 10, 0, 2000, 5, 0, 64, 0, 1500, 3], [0, 4, 77, 500, 0, 0, 1, 250, 0, 0]]
    mem = [[] for _ in range(7)]
    grp = [[[] for _ in range(4)] for _ in gsz]
    for h, n in enumerate(sizes):
        centre = rng.uniform(-5, 5, 3)
        p = synth.plummer_stars(n, 1.5e-3 * max(n, 1) ** (1 / 3) / 10, 1e6, rng, centre=centre,
                                bulk_v=rng.normal(0, 100, 3), interloper_frac=0.12)
        for col, a in zip(mem, (p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)):
            col.append(a)
        for k in range(2):
            q = synth.dm_cloud(gsz[k][h], 4e-3, 2e6, rng, centre=centre)
            for col, a in zip(grp[k], (q.mass, q.x, q.y, q.z)):
                col.append(a)
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    mem = [np.concatenate(col) for col in mem]
    groups = []
    for k in range(2):
        eo = np.concatenate(([0], np.cumsum(gsz[k]))).astype(np.int64)
        groups.append((eo,) + tuple(np.concatenate(col) for col in grp[k]))
    kw = dict(groups=groups, n_pre=1, kappa=4.0, mode="fast")
    ref = unbind_catalogue(off, *mem, cache_external=False, incremental=False, **kw)
    for v in VARIANTS:
        res = unbind_catalogue(off, *mem, **v, **kw)
        assert res.stats.pairs == ref.stats.pairs and res.stats.evaluations < ref.stats.evaluations
        for h in range(len(sizes)):
            a, b = off[h], off[h + 1]
            pre = [tuple(arr[groups[0][0][h]:groups[0][0][h + 1]] for arr in groups[0][1:])]
            post = [tuple(arr[groups[1][0][h]:groups[1][0][h + 1]] for arr in groups[1][1:])]
            o = O.unbind_halo(*[col[a:b] for col in mem], pre=pre, post=post, kappa=4.0, variant="f64acc")
            diff = res.halo_mask(h) != o.mask
            assert np.all(O.energy_margin(o.energy, o.be32, 4.0)[diff] < BAND), h
            if not diff.any():
                assert res.halos[h].n_iter == o.n_iter, h
                both = o.mask
                if both.any():
                    assert np.abs(res.be32[a:b][both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL, h
                assert np.array_equal(res.members(h), o.idx), h
        np.testing.assert_allclose(res.be32, ref.be32, rtol=FAST_RTOL)
        # the same loop as ONE CUDA-graph launch (device-side WHILE node): identical results
        eoffs = [g_[0] for g_ in groups]
        with UnbindPlan(off, eoffs, mode="fast", n_pre=1, kappa=4.0, use_graph=True, **v) as plan:
            plan.upload_members(*mem)
            for k, g_ in enumerate(groups):
                plan.upload_group(k, *g_[1:])
            st = plan.run()
            gr = plan.download()
            st2 = plan.run()                     # plans are re-runnable: kept sums are rebuilt from scratch
            gr2 = plan.download()
        assert st.passes == res.stats.passes and st.evaluations == res.stats.evaluations == st2.evaluations
        assert np.array_equal(gr.mask, res.mask) and np.array_equal(gr.be32.view(np.uint32), res.be32.view(np.uint32))
        assert np.array_equal(gr2.mask, gr.mask) and np.array_equal(gr2.be32.view(np.uint32), gr.be32.view(np.uint32))


def test_duplicates_fall_back_and_recover(monkeypatch):
    # Exact duplicates (zero separation) make the predicate-free sums non-finite: the pass falls back to the
    # predicated kernel, its sums are not kept, and the next pass is a full one.  One duplicate pair sits
    # between a member and an EXTERNAL source (spoils the cached external sums of that halo for good), one
    # between two members.
    monkeypatch.setenv("HALMA_NP_MIN_PAIRS", "0")
    rng = np.random.default_rng(41)
    p = synth.plummer_stars(3000, 2e-3, 1e6, rng, interloper_frac=0.1)
    d = synth.dm_cloud(900, 5e-3, 3e6, rng)
    kw = dict(kappa=9.0, mode="fast")
    for which in ("member-ext", "member-member", "both"):
        x, y, z = p.x.copy(), p.y.copy(), p.z.copy()
        dx, dy, dz = d.x.copy(), d.y.copy(), d.z.copy()
        if which in ("member-ext", "both"):
            dx[700], dy[700], dz[700] = x[100], y[100], z[100]
        if which in ("member-member", "both"):
            x[2900], y[2900], z[2900] = x[5], y[5], z[5]
        args = (x, y, z, p.vx, p.vy, p.vz, p.mass)
        post = [(d.mass, dx, dy, dz)]
        o = O.unbind_halo(*args, post=post, kappa=9.0, variant="f64acc")
        ref = unbind_halo(*args, post=post, cache_external=False, incremental=False, **kw)
        for v in VARIANTS:
            r = unbind_halo(*args, post=post, **v, **kw)
            assert np.all(np.isfinite(r.be32))
            if np.array_equal(r.mask, ref.mask):
                assert r.n_iter == ref.n_iter
            np.testing.assert_allclose(r.be32, ref.be32, rtol=FAST_RTOL)
            diff = r.mask != o.mask
            assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[diff] < BAND)


def test_incremental_pass_that_removes_most_of_a_members_potential(monkeypatch):
    # A tight pair: B sits 1e-7 Mpc from A and provides ~90 % of A's potential; B is an interloper and leaves in
    # the first pass, A stays.  Subtracting B's term from A's kept potential would amplify the float32 rounding
    # of that term ten times, so the incremental ticket hands the halo to the predicated kernel in that pass
    # (csrc/potential.cu, kIncrHeavy): results as exact as the plain loop's, one pass counted as a full one.
    monkeypatch.setenv("HALMA_NP_MIN_PAIRS", "0")
    rng = np.random.default_rng(51)
    p = synth.plummer_stars(3000, 2e-3, 1e6, rng, centre=(0.0, 0.0, 0.0), bulk_v=(0., 0., 0.), interloper_frac=0.1)
    o0 = O.unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, kappa=9.0, variant="f64acc")
    a = int(np.flatnonzero(o0.mask)[100])                       # a member that stays bound
    b = int(np.flatnonzero(~o0.mask & (o0.be32 > 0))[0])        # one that leaves
    x, y, z = p.x.copy(), p.y.copy(), p.z.copy()
    vx, vy, vz = p.vx.copy(), p.vy.copy(), p.vz.copy()
    x[b], y[b], z[b] = x[a] + 3e-8, y[a] + 4e-8, z[a] - 5e-8
    vx[b], vy[b], vz[b] = 3e4, -2e4, 1e4                        # far beyond any escape velocity
    args = (x, y, z, vx, vy, vz, p.mass)
    o = O.unbind_halo(*args, kappa=9.0, variant="f64acc")
    assert o.mask[a] and not o.mask[b] and o.n_iter >= 3
    first = O.unbind_halo(*args, kappa=9.0, variant="f64acc", max_iter=1)
    assert first.be32[a] > 5 * o.be32[a]                        # B carried most of A's potential
    ref = unbind_halo(*args, kappa=9.0, mode="fast", cache_external=False, incremental=False)
    r = unbind_halo(*args, kappa=9.0, mode="fast", cache_external=False, incremental=True)
    assert np.array_equal(r.mask, ref.mask) and r.n_iter == ref.n_iter
    np.testing.assert_allclose(r.be32, ref.be32, rtol=FAST_RTOL)
    both = r.mask & o.mask
    assert np.abs(r.be32[both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL
    assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[r.mask != o.mask] < BAND)
    if np.array_equal(r.mask, o.mask):
        # the second pass was NOT taken incrementally (it is counted as a full, predicated one) and the pass after
        # a fallback is a full one: exactly the passes of the numpy model of the bookkeeping
        m = RM.unbind_halo(*args, kappa=9.0, cache_external=False, incremental=True)
        assert m.passes == ["full", "fallback", "full"]
        assert r.stats.evaluations == m.evaluations
        assert r.stats.evaluations > RM.expected_evaluations(o.n_bound_history, 0, cache_external=False, incremental=True)


def test_reuse_forced_on_for_the_whole_battery():
    """The ragged / degenerate / duplicate / NaN / external-group / golden / full-size / graph battery of
    test_gpu_unbind.py with both options on and every plan forced onto the predicate-free kernel."""
    if os.environ.get("HALMA_CACHE_EXT") == "1" and os.environ.get("HALMA_NP_MIN_PAIRS") == "0":
        pytest.skip("already running with reuse forced on")
    small = ("ragged or degenerate or duplicates or zero_mass or external_groups or golden_fast or fused_large "
             "or fast_mode_against or lattice or rps_mass_sums or idempotence or graph_loop or fixed_bulk")
    # with the symmetric tickets (throughput shape forced) including the full-size cases; without them (the
    # kernel shape picked per plan) on the small ones
    for extra, sel in ((dict(HALMA_SYMMETRIC="1", HALMA_FAST_VARIANT="0"), small + " or cfg3_full or cfg2_full"),
                       (dict(HALMA_SYMMETRIC="0"), small)):
        env = dict(os.environ, HALMA_NP_MIN_PAIRS="0", HALMA_CACHE_EXT="1", HALMA_INCREMENTAL="1", **extra)
        out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_unbind.py"), "-q",
                              "-x", "-m", "gpu", "-k", sel], capture_output=True, text=True, timeout=1500, env=env,
                             cwd=ROOT)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
        assert " passed" in out.stdout and "failed" not in out.stdout
