"""GPU parity of the device-resident unbinding loop against the CPU oracle and the golden
fixtures written by the reference's own Python drivers.

EXACT mode must reproduce the oracle (mask, member indices, float32 potentials, float64
energies, pass counts) bit for bit.  FAST mode must agree with the float64-accumulating
oracle to 1e-6 on potentials and exactly on the mask outside the |E|/max(KE,|PE|) < 1e-6
band (north_star); differences against the reference's float32 arithmetic are counted and
must all sit close to E = 0.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from pyhalma_b200 import _lib, halo_gas, halo_properties, synth
from pyhalma_b200.unbind import UnbindPlan, unbind_catalogue, unbind_halo

pytestmark = pytest.mark.gpu
FAST_RTOL = 1e-6
BAND = 1e-6


def case(n_star=3000, n_gas=2000, n_dm=500, extra=0):
    c = synth.config1(n_star, n_gas, seed_extra=extra, n_dm=n_dm)
    return c.stars, c.gas, c.dm


def assert_same_exact(r, o):
    assert np.array_equal(r.mask, o.mask)
    assert np.array_equal(r.idx, o.idx)
    assert np.array_equal(r.be32.view(np.uint32), o.be32.view(np.uint32))
    # the bulk velocity is a float64 reduction whose order differs from numpy's (the
    # reference's own numba reduction is order-nondeterministic): energies agree to ~1e-13
    np.testing.assert_allclose(r.energy, o.energy, rtol=1e-11, atol=1e-9 * np.abs(o.energy).max())
    assert r.n_iter == o.n_iter and r.pairs == o.pairs
    np.testing.assert_allclose(r.mass, o.mass, rtol=1e-13)
    np.testing.assert_allclose(r.com, o.com, rtol=1e-13)
    np.testing.assert_allclose(r.vb, o.vb, rtol=1e-11, atol=1e-11)


def test_stellar_layout_exact_matches_oracle():
    s, g, d = case()
    kw = dict(pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0)
    o = O.unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, **kw)
    r = unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, mode="exact", **kw)
    assert o.n_iter >= 3                      # the case really iterates
    assert_same_exact(r, o)
    assert r.converged


def test_gas_layout_split_classes_exact_matches_oracle():
    s, g, d = case(extra=1)
    vb = O.CM_velocity(O.total_mass(np.arange(len(s)), s.mass), np.arange(len(s)), s.vx, s.vy, s.vz, s.mass)
    kw = dict(post=[d.pos_mass(), s.pos_mass()], split_classes=True, kappa=2.0, vb_fixed=vb)
    o = O.unbind_halo(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass, **kw)
    r = unbind_halo(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass, mode="exact", **kw)
    assert o.n_iter >= 2 and 0 < o.mask.sum() < len(g)
    assert_same_exact(r, o)


@pytest.mark.parametrize("max_iter", [1, 2])
def test_max_iter_cap(max_iter):
    s, g, d = case(1500, 900, 200, extra=2)
    kw = dict(pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0, max_iter=max_iter)
    o = O.unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, **kw)
    r = unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, mode="exact", **kw)
    assert r.n_iter == max_iter
    assert_same_exact(r, o)
    assert not r.converged


def test_fast_mode_against_f64acc_and_reference_arithmetic():
    s, g, d = case(6000, 4000, 800, extra=3)
    kw = dict(pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0)
    o64 = O.unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, variant="f64acc", **kw)
    o32 = O.unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, variant="f32seq", **kw)
    r = unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, mode="fast", **kw)
    # potentials: 1e-6 relative against the float64-accumulating evaluation (same passes)
    both = r.mask & o64.mask
    rel = np.abs(r.be32[both].astype(np.float64) / o64.be32[both] - 1)
    assert rel.max() < FAST_RTOL
    # mask: identical outside the band
    diff = r.mask != o64.mask
    margin = O.energy_margin(o64.energy, o64.be32, 9.0)
    assert np.all(margin[diff] < BAND), (diff.sum(), margin[diff])
    # against the reference's own float32 arithmetic: count and classify by margin
    diff32 = r.mask != o32.mask
    margin32 = O.energy_margin(o32.energy, o32.be32, 9.0)
    assert diff32.sum() <= 3 and np.all(margin32[diff32] < 1e-4)
    if not diff.any():
        assert r.n_iter == o64.n_iter
        np.testing.assert_allclose(r.mass, o64.mass, rtol=1e-12)
        np.testing.assert_allclose(r.vb, o64.vb, rtol=1e-6)
        np.testing.assert_allclose(r.com, o64.com, rtol=1e-6)


def test_gas_layout_fast_lattice_exclusions():
    # lattice gas: ~10 % of the pairs share a coordinate; the predicate-free path must take
    # exactly those out again (correction tickets)
    s, g, d = case(3000, 6000, 400, extra=6)
    vb = O.CM_velocity(O.total_mass(np.arange(len(s)), s.mass), np.arange(len(s)), s.vx, s.vy, s.vz, s.mass)
    kw = dict(post=[d.pos_mass(), s.pos_mass()], split_classes=True, kappa=2.0, vb_fixed=vb)
    o = O.unbind_halo(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass, variant="f64acc", **kw)
    r = unbind_halo(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass, mode="fast", **kw)
    assert o.n_iter >= 2
    both = r.mask & o.mask
    assert np.abs(r.be32[both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL
    diff = r.mask != o.mask
    assert np.all(O.energy_margin(o.energy, o.be32, 2.0)[diff] < BAND)
    first = np.abs(r.be32.astype(np.float64)[~o.mask & (o.be32 > 0)] / o.be32[~o.mask & (o.be32 > 0)] - 1)
    assert first.size == 0 or first.max() < FAST_RTOL          # particles removed in earlier passes too


def test_coordinate_sharing_runs_of_every_length_fast(monkeypatch):
    monkeypatch.setenv("HALMA_NP_MIN_PAIRS", "0")          # the predicate-free kernel + correction tickets, whatever the size
    monkeypatch.setenv("HALMA_FAST_VARIANT", "0")          # ... in the throughput shape, which has the symmetric tickets
    # Correction tickets walk short runs of equal keys per member and stream long ones through the ring
    # (potential_device.cuh::kCorrRunMax = 16 each side): runs of 2 .. 40 members sharing x, y or z, some of them
    # together with external sources, straddle both paths and the switch between them inside one ticket.
    rng = np.random.default_rng(31)
    p = synth.plummer_stars(2600, 3e-3, 1e6, rng, centre=(11.0, -7.0, 3.0), bulk_v=(0., 0., 0.))
    d = synth.plummer_stars(500, 6e-3, 4e6, rng, centre=(11.0, -7.0, 3.0), bulk_v=(0., 0., 0.))
    free = rng.permutation(len(p))
    used = 0
    for k, length in enumerate((2, 3, 5, 15, 16, 17, 18, 31, 33, 34, 40, 2, 2, 3, 4, 9)):
        members = free[used:used + length]
        used += length
        arr = (p.x, p.y, p.z)[k % 3]
        arr[members] = arr[members[0]]
        if k % 4 == 1:                      # externals that share the run's coordinate as well
            ext = rng.choice(len(d), 3, replace=False)
            (d.x, d.y, d.z)[k % 3][ext] = arr[members[0]]
    kw = dict(post=[d.pos_mass()], kappa=9.0)
    o = O.unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, variant="f64acc", **kw)
    for sym in (True, False):
        r = unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, mode="fast", symmetric=sym, **kw)
        both = r.mask & o.mask
        assert np.abs(r.be32[both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL
        diff = r.mask != o.mask
        assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[diff] < BAND)
        assert r.n_iter == o.n_iter or diff.any()
        assert r.stats.evaluations < r.stats.pairs or not sym          # i.e. it did run the predicate-free path


def test_exact_duplicates_and_signed_zero_fast():
    # exact duplicates far apart in index (different tiles) give zero separations the
    # predicate-free path cannot handle: the halo must be recomputed with the predicate.
    rng = np.random.default_rng(12)
    p = synth.plummer_stars(700, 2e-3, 1e6, rng, centre=(0.0, 0.0, 0.0), bulk_v=(0., 0., 0.))
    for a, b in ((3, 500), (10, 650), (300, 301)):
        p.x[b], p.y[b], p.z[b] = p.x[a], p.y[a], p.z[a]
    p.x[40], p.x[600] = 0.0, -0.0                 # -0 == +0 for the reference's /= test
    p.y[41], p.y[333] = -0.0, 0.0
    o = O.unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, kappa=9.0, variant="f64acc")
    r = unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, kappa=9.0, mode="fast")
    assert np.all(np.isfinite(r.be32))
    both = r.mask & o.mask
    assert np.abs(r.be32[both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL
    diff = r.mask != o.mask
    assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[diff] < BAND)
    # the same without duplicates but with signed zeros only (no fallback needed)
    q = synth.plummer_stars(700, 2e-3, 1e6, rng, centre=(0.0, 0.0, 0.0), bulk_v=(0., 0., 0.))
    q.x[40], q.x[600] = 0.0, -0.0
    q.z[5], q.z[6], q.z[7] = 0.0, -0.0, 0.0
    o = O.unbind_halo(q.x, q.y, q.z, q.vx, q.vy, q.vz, q.mass, kappa=9.0, variant="f64acc")
    r = unbind_halo(q.x, q.y, q.z, q.vx, q.vy, q.vz, q.mass, kappa=9.0, mode="fast")
    both = r.mask & o.mask
    assert np.abs(r.be32[both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL


def ragged_catalogue():
    rng = np.random.default_rng(77)
    sizes = [0, 1, 2, 3, 4, 5, 31, 32, 33, 127, 128, 129, 255, 256, 257, 600, 1025, 0, 40, 2048, 7, 300]
    cols = [[] for _ in range(7)]
    for n in sizes:
        centre = rng.uniform(-5, 5, 3)
        p = synth.plummer_stars(n, 1.5e-3 * max(n, 1) ** (1 / 3) / 10, 1e6, rng, centre=centre,
                                bulk_v=rng.normal(0, 100, 3), interloper_frac=0.15)
        for c, a in zip(cols, (p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)):
            c.append(a)
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    return off, [np.concatenate(c) for c in cols]


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_ragged_catalogue_batched(mode):
    off, cols = ragged_catalogue()
    res = unbind_catalogue(off, *cols, kappa=1.0, mode=mode)
    assert res.stats.launches > 0 and res.stats.potential_launches >= 1
    total_pairs = 0
    for h in range(len(off) - 1):
        a, b = off[h], off[h + 1]
        o = O.unbind_halo(*[c[a:b] for c in cols], kappa=1.0, variant="f32seq" if mode == "exact" else "f64acc")
        hr = res.halos[h]
        m = res.halo_mask(h)
        if mode == "exact":
            assert np.array_equal(m, o.mask), h
            assert np.array_equal(res.members(h), o.idx)
            assert np.array_equal(res.be32[a:b].view(np.uint32), o.be32.view(np.uint32))
            np.testing.assert_allclose(res.energy[a:b], o.energy, rtol=1e-11,
                                       atol=1e-9 * (np.abs(o.energy).max() if len(o.energy) else 0))
            assert hr.n_iter == o.n_iter and hr.n_bound == len(o.idx) and hr.pairs == o.pairs
            np.testing.assert_allclose(hr.vb, o.vb, rtol=1e-10, atol=1e-9)
            np.testing.assert_allclose(hr.mass, o.mass, rtol=1e-13)
        else:
            diff = m != o.mask
            margin = O.energy_margin(o.energy, o.be32, 1.0)
            assert np.all(margin[diff] < BAND)
        total_pairs += hr.pairs
    assert res.stats.pairs == total_pairs


def test_catalogue_with_external_groups_exact():
    rng = np.random.default_rng(5)
    sizes = [300, 0, 77, 513]
    gsz = [[100, 5, 0, 260], [40, 40, 3, 0]]
    mem = [[] for _ in range(7)]
    grp = [[[] for _ in range(4)] for _ in gsz]
    centres = []
    for h, n in enumerate(sizes):
        centre = rng.uniform(-5, 5, 3)
        centres.append(centre)
        p = synth.plummer_stars(n, 2e-3, 1e6, rng, centre=centre, bulk_v=(10., 20., -30.))
        for c, a in zip(mem, (p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)):
            c.append(a)
        for k in range(2):
            q = synth.dm_cloud(gsz[k][h], 4e-3, 5e6, rng, centre=centre)
            for c, a in zip(grp[k], (q.mass, q.x, q.y, q.z)):
                c.append(a)
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    mem = [np.concatenate(c) for c in mem]
    groups = []
    for k in range(2):
        eo = np.concatenate(([0], np.cumsum(gsz[k]))).astype(np.int64)
        groups.append((eo,) + tuple(np.concatenate(c) for c in grp[k]))
    res = unbind_catalogue(off, *mem, groups=groups, n_pre=1, kappa=4.0, mode="exact")
    for h in range(len(sizes)):
        a, b = off[h], off[h + 1]
        pre = [tuple(arr[groups[0][0][h]:groups[0][0][h + 1]] for arr in groups[0][1:])]
        post = [tuple(arr[groups[1][0][h]:groups[1][0][h + 1]] for arr in groups[1][1:])]
        o = O.unbind_halo(*[c[a:b] for c in mem], pre=pre, post=post, kappa=4.0)
        assert np.array_equal(res.halo_mask(h), o.mask), h
        assert np.array_equal(res.be32[a:b].view(np.uint32), o.be32.view(np.uint32)), h
        assert res.halos[h].n_iter == o.n_iter
    # the same job through the one-shot C entry point halma_unbind_catalogue (CSR offsets, host pointers)
    import ctypes as C
    L = _lib.lib()
    cfg = _lib.UnbindConfig()
    cfg.struct_size, cfg.mode, cfg.n_groups, cfg.n_pre, cfg.max_iter = C.sizeof(cfg), _lib.MODE_EXACT, 2, 1, 64
    cfg.G, cfg.kappa, cfg.n_ranks = 4.3e-9, 4.0, 1
    n, nh = int(off[-1]), len(sizes)
    mem64 = [np.ascontiguousarray(a, np.float64) for a in mem]
    geo = [np.ascontiguousarray(g[0], np.int64) for g in groups]
    gcols = [[np.ascontiguousarray(g[k], np.float64) for g in groups] for k in range(1, 5)]
    tbl = lambda arrs, T: (T * len(arrs))(*[a.ctypes.data_as(T) for a in arrs])      # noqa: E731
    I64P, F64P = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    mask, be = np.zeros(n, np.uint8), np.zeros(n, np.float32)
    en, idx = np.zeros(n, np.float64), np.zeros(n, np.int32)
    hr = (_lib.HaloResult * nh)()
    st = _lib.RunStats()
    L.halma_unbind_catalogue.argtypes = None
    rc = L.halma_unbind_catalogue(C.byref(cfg), C.c_int64(nh), off.ctypes.data_as(I64P),
                                  *[a.ctypes.data_as(F64P) for a in mem64], tbl(geo, I64P),
                                  tbl(gcols[0], F64P), tbl(gcols[1], F64P), tbl(gcols[2], F64P), tbl(gcols[3], F64P),
                                  None, None, C.c_double(0.0), mask.ctypes.data_as(C.c_void_p),
                                  be.ctypes.data_as(C.c_void_p), en.ctypes.data_as(C.c_void_p),
                                  idx.ctypes.data_as(C.c_void_p), hr, C.byref(st))
    assert rc == 0, L.halma_last_error()
    assert np.array_equal(mask, res.mask) and np.array_equal(be.view(np.uint32), res.be32.view(np.uint32))
    assert np.array_equal(en, res.energy) and np.array_equal(idx, res.idx_packed)
    assert [hr[h].n_bound for h in range(nh)] == [res.halos[h].n_bound for h in range(nh)]
    assert st.pairs == res.stats.pairs and st.passes == res.stats.passes


def test_idempotence_and_rerun():
    s, g, d = case(4000, 0, 300, extra=4)
    off = np.array([0, len(s)], np.int64)
    with UnbindPlan(off, [np.array([0, len(d)], np.int64)], mode="fast", kappa=9.0) as plan:
        plan.upload_members(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass)
        plan.upload_group(0, d.mass, d.x, d.y, d.z)
        plan.run()
        r1 = plan.download()
        plan.run()                              # plans are re-runnable from the pristine inputs
        r2 = plan.download()
    assert np.array_equal(r1.mask, r2.mask) and np.array_equal(r1.be32, r2.be32)
    assert r1.halos[0].n_iter >= 2
    i = r1.members(0)
    again = unbind_halo(s.x[i], s.y[i], s.z[i], s.vx[i], s.vy[i], s.vz[i], s.mass[i], post=[d.pos_mass()],
                        kappa=9.0, mode="fast")
    assert again.n_iter == 1 and again.mask.all()


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_graph_loop_driver_matches_enqueue_driver(mode):
    # The three loop drivers run the same phase code: "fused" (default on one GPU) is ONE persistent cooperative
    # kernel for the whole loop, "enqueue" stand-alone kernels queued ahead by the host, "graph" one CUDA-graph
    # launch with a device-side WHILE node.  Results must agree bit for bit.
    off, cols = ragged_catalogue()
    out = []
    for driver in ("fused", "enqueue", "graph"):
        with UnbindPlan(off, mode=mode, kappa=1.0, driver=driver) as plan:
            plan.upload_members(*cols)
            st = plan.run()
            res = plan.download()
            st2 = plan.run()                     # plans are re-runnable (re-launch of the instantiated graph)
            res2 = plan.download()
        assert np.array_equal(res.mask, res2.mask) and st.passes == st2.passes
        assert np.array_equal(res.be32.view(np.uint32), res2.be32.view(np.uint32))
        assert st.driver == {"fused": _lib.DRIVER_FUSED, "enqueue": _lib.DRIVER_ENQUEUE, "graph": _lib.DRIVER_GRAPH}[driver]
        if driver == "fused":
            # one launch per run (the first run of a plan on the predicate-free path also packs and sorts)
            assert st2.launches == 1 and st2.potential_ms > 0 and st2.loop_ms >= st2.potential_ms
        out.append((st, res))
    (sa, a) = out[0]
    for sb, b in out[1:]:
        assert sa.passes == sb.passes and sa.pairs == sb.pairs and sa.evaluations == sb.evaluations
        assert np.array_equal(a.mask, b.mask) and np.array_equal(a.be32.view(np.uint32), b.be32.view(np.uint32))
        assert np.array_equal(a.energy, b.energy) and np.array_equal(a.idx_packed, b.idx_packed)
        assert [h.n_iter for h in a.halos] == [h.n_iter for h in b.halos]
        assert np.array_equal(a.halos.field("vb"), b.halos.field("vb")) and np.array_equal(a.halos.field("mass"), b.halos.field("mass"))
    # a large single halo through the predicate-free kernel (+ symmetric self-term, cached and incremental sums)
    s, g, d = case(40000, 90000, 1000, extra=9)
    kw = dict(pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0, mode=mode)
    if mode == "fast":
        import os
        rs = []
        for driver in ("fused", "enqueue", "graph"):
            os.environ["HALMA_DRIVER"] = driver
            try:
                rs.append(unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, **kw))
            finally:
                os.environ.pop("HALMA_DRIVER")
        assert rs[0].stats.driver == _lib.DRIVER_FUSED and rs[1].stats.driver == _lib.DRIVER_ENQUEUE
        for r1 in rs[1:]:
            assert np.array_equal(rs[0].mask, r1.mask) and rs[0].n_iter == r1.n_iter
            assert np.array_equal(rs[0].be32.view(np.uint32), r1.be32.view(np.uint32))


def test_degenerate_inputs():
    r = unbind_halo([], [], [], [], [], [], [], mode="exact")
    assert r.n_iter == 0 and r.mass == 0.0 and r.vb == (0., 0., 0.) and len(r.idx) == 0
    r = unbind_halo([0.], [0.], [0.], [1.], [0.], [0.], [1e6], vb_fixed=(0., 0., 0.), mode="exact")
    assert not r.mask[0] and r.n_iter == 1 and r.converged       # Phi = 0 and KE > 0 -> unbound
    r = unbind_halo([0.], [0.], [0.], [1.], [0.], [0.], [1e6], mode="exact")
    assert r.mask[0] and r.n_iter == 1                            # vb = own velocity: E = 0 is bound
    r = unbind_halo([0., 1.], [0., 1.], [0., 1.], [5., 5.], [0., 0.], [0., 0.], [0., 0.], mode="fast")
    assert r.mass == 0.0 and r.vb == (0., 0., 0.)                 # M == 0 -> vb = 0 (halo_properties.py:57-60)


# ---- golden fixtures from the reference's own Python drivers --------------------------------
def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


@pytest.mark.parametrize("name", ["rps_one_dm", "rps_two_dm", "rps_sampled"])
@pytest.mark.parametrize("fused", [True, False])
def test_rps_golden_exact(golden_dir, name, fused):
    g = load(golden_dir, name)
    np.random.seed(int(g["np_seed"]))
    out = halo_gas.RPS(g["gas_x"], g["gas_y"], g["gas_z"], g["gas_vx"], g["gas_vy"], g["gas_vz"], g["gas_mass"],
                       g["gas_temp"], g["dm_x"], g["dm_y"], g["dm_z"], g["dm_mass"], g["st_x"], g["st_y"],
                       g["st_z"], g["st_mass"], *g["vb"], int(g["lim"]), float(g["mass_dm_part"]),
                       int(g["num_dm_species"]), mode="exact", fused=fused)
    np.testing.assert_array_equal(np.array(out, dtype=np.float64), g["rps_out"])


@pytest.mark.parametrize("name", ["rps_one_dm", "rps_two_dm"])
def test_rps_golden_fast(golden_dir, name):
    g = load(golden_dir, name)
    out = halo_gas.RPS(g["gas_x"], g["gas_y"], g["gas_z"], g["gas_vx"], g["gas_vy"], g["gas_vz"], g["gas_mass"],
                       g["gas_temp"], g["dm_x"], g["dm_y"], g["dm_z"], g["dm_mass"], g["st_x"], g["st_y"],
                       g["st_z"], g["st_mass"], *g["vb"], int(g["lim"]), float(g["mass_dm_part"]),
                       int(g["num_dm_species"]), mode="fast", return_mask=True)
    ref = O.RPS(g["gas_x"], g["gas_y"], g["gas_z"], g["gas_vx"], g["gas_vy"], g["gas_vz"], g["gas_mass"],
                g["gas_temp"], g["dm_x"], g["dm_y"], g["dm_z"], g["dm_mass"], g["st_x"], g["st_y"], g["st_z"],
                g["st_mass"], *g["vb"], int(g["lim"]), float(g["mass_dm_part"]), int(g["num_dm_species"]))
    diff = out[4] != ref.bound
    assert np.all(O.energy_margin(ref.energy, ref.be32, 2.0)[diff] < 1e-5)
    if not diff.any():
        np.testing.assert_array_equal(np.array(out[:4], dtype=np.float64), g["rps_out"])


@pytest.mark.parametrize("name", ["rps_one_dm", "rps_two_dm", "rps_sampled"])
@pytest.mark.parametrize("fused", [True, False])
def test_most_bound_golden_exact(golden_dir, name, fused):
    g = load(golden_dir, name)
    np.random.seed(int(g["np_seed"]))
    oripa = np.arange(len(g["st_x"])) + 1000
    out = halo_gas.most_bound_particle(g["gas_x"], g["gas_y"], g["gas_z"], g["gas_mass"], g["dm_x"], g["dm_y"],
                                       g["dm_z"], g["dm_mass"], g["st_x"], g["st_y"], g["st_z"], g["st_mass"],
                                       oripa, int(g["lim"]), float(g["mass_dm_part"]), mode="exact", fused=fused)
    np.testing.assert_array_equal(np.array(out, dtype=np.float64), g["mb_out"])


def test_most_bound_fused_large_matches_oracle_argmin():
    # device-side arg-max of the potential (lowest index on ties) == np.argmin(-be) of the oracle
    s, g, d = case(20000, 8000, 1500, extra=8)
    oripa = np.arange(len(s)) + 7
    for mode, variant in (("exact", "f32seq"), ("fast", "f64acc")):
        out = halo_gas.most_bound_particle(g.x, g.y, g.z, g.mass, d.x, d.y, d.z, d.mass, s.x, s.y, s.z, s.mass,
                                           oripa, 10 ** 9, 8e7, mode=mode)
        be = O.most_bound_potential(g.x, g.y, g.z, g.mass, d.x, d.y, d.z, d.mass, s.x, s.y, s.z, s.mass, 10 ** 9,
                                    8e7, variant=variant)
        k = int(np.argmin(-be))
        if mode == "exact":
            assert out[3] == oripa[k]
        else:
            kk = int(out[3] - 7)          # fast: the winner must be within rounding of the oracle's minimum
            assert abs(be[kk] / be[k] - 1) < 1e-6


def test_rps_mass_sums_on_device(golden_dir):
    g = load(golden_dir, "rps_two_dm")
    ext = [(g["dm_mass"], g["dm_x"], g["dm_y"], g["dm_z"]), (g["st_mass"], g["st_x"], g["st_y"], g["st_z"])]
    n = len(g["gas_x"])
    groups = [(np.array([0, len(e[0])], np.int64),) + e for e in ext]
    for max_iter in (1, 64):
        res = unbind_catalogue(np.array([0, n], np.int64), g["gas_x"], g["gas_y"], g["gas_z"], g["gas_vx"], g["gas_vy"],
                               g["gas_vz"], g["gas_mass"], groups=groups, split_classes=True, vb=g["vb"].reshape(1, 3),
                               kappa=2.0, max_iter=max_iter, mode="exact", temp=g["gas_temp"])
        h = res.halos[0]
        bound = res.mask.astype(bool)
        cold = g["gas_temp"] < 5e4
        m = g["gas_mass"]
        np.testing.assert_allclose(h.mass_initial, m.sum(), rtol=1e-13)
        np.testing.assert_allclose(h.cold_bound_mass, m[cold & bound].sum(), rtol=1e-12)
        np.testing.assert_allclose(h.unbound_cold_mass, m[cold & ~bound].sum(), rtol=1e-12)
        np.testing.assert_allclose(h.unbound_hot_mass, m[~cold & ~bound].sum(), rtol=1e-12)
        assert 0 <= h.most_bound < n and res.be32[h.most_bound] == res.be32[bound | ~bound].max() or max_iter > 1


def test_stellar_onepass_golden(golden_dir):
    g = load(golden_dir, "stellar_onepass")
    gas = (g["gas_x"], g["gas_y"], g["gas_z"], g["gas_mass_seen"])
    stars = tuple(g["st_" + k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass"))
    dm = (g["dm_x"], g["dm_y"], g["dm_z"], g["dm_mass"])
    r = halo_properties.escape_velocity_unbinding(gas, stars, dm, g["vb"], float(g["factor_v"]), mode="exact")
    np.testing.assert_array_equal(r.mask, g["bound"])
    np.testing.assert_array_equal(r.be32, g["call0_be"])
    assert r.n_iter == 1
    rf = halo_properties.escape_velocity_unbinding(gas, stars, dm, g["vb"], float(g["factor_v"]), mode="fast")
    assert (rf.mask != g["bound"]).sum() == 0
    # the reference signature, with the AMR gather bound to a stand-in (make_golden.py does the same)
    rete = 0.8
    halo_gas_amr = halo_gas.AMRgrid_to_particles
    try:
        halo_gas.AMRgrid_to_particles = lambda *a: (g["gas_x"], g["gas_y"], g["gas_z"], None, None, None,
                                                    g["gas_mass"] / rete ** 3, None)
        n_glob = 400
        glob = {k: np.zeros(n_glob) for k in ("x", "y", "z", "vx", "vy", "vz", "mass")}
        for k in glob:
            glob[k][g["part_list"]] = g["st_" + k]
        cx, cy, cz = g["com"]
        dmx = np.concatenate((g["dm_x"], [cx + 5.0, cx - 7.0]))
        dmy = np.concatenate((g["dm_y"], [cy, cy]))
        dmz = np.concatenate((g["dm_z"], [cz, cz]))
        dmm = np.concatenate((g["dm_mass"], [1e15, 1e15]))
        bound = halo_properties.escape_velocity_unbinding_fortran(
            rete, 40.0, 128, None, None, (dmx, dmy, dmz, dmm), cx, cy, cz, *g["vb"], 1.0, g["part_list"],
            glob["x"], glob["y"], glob["z"], glob["vx"], glob["vy"], glob["vz"], glob["mass"],
            float(g["factor_v"]), 1.0, mode="exact", mass_to_sun=1.0)
    finally:
        halo_gas.AMRgrid_to_particles = halo_gas_amr
    np.testing.assert_array_equal(bound, g["bound"])


def test_reference_call_sequences_replayed_through_the_dropin(golden_dir, monkeypatch):
    """Every kernel call the reference's RPS and most_bound_particle made on the golden cases (recorded with its
    float32 inputs by tests/golden/make_golden.py::call_sequences while the reference's own Python ran), replayed in
    order through the drop-in package exactly as the reference calls it -- `from fortran_modules import particle;
    particle.particle.brute_force_binding_energy(ncores, ntotal, ..., ntest, ...)` -- in EXACT mode (HALMA_MODE):
    every result must equal the recorded one bit for bit."""
    import importlib
    import sys
    monkeypatch.setenv("HALMA_MODE", "exact")
    monkeypatch.syspath_prepend(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin"))
    for m in [k for k in sys.modules if k == "fortran_modules" or k.startswith("fortran_modules.")]:
        monkeypatch.delitem(sys.modules, m)
    particle = importlib.import_module("fortran_modules.particle")
    g = np.load(os.path.join(golden_dir, "call_sequences.npz"))
    n_checked = 0
    for tag in ("rps_one_dm", "rps_two_dm", "rps_sampled"):
        for part in ("rps", "mb"):
            for k in range(int(g["%s_%s_n_calls" % (tag, part)])):
                a = {n: g["%s_%s_call%d_%s" % (tag, part, k, n)] for n in ("tm", "tx", "ty", "tz", "sx", "sy", "sz", "be")}
                out = particle.particle.brute_force_binding_energy(
                    np.int32(8), np.int32(len(a["tm"])), a["tm"], a["tx"], a["ty"], a["tz"], np.int32(len(a["sx"])),
                    a["sx"], a["sy"], a["sz"])
                assert out.dtype == np.float32 and out.shape == a["be"].shape
                assert np.array_equal(out.view(np.uint32), a["be"].view(np.uint32)), (tag, part, k)
                ser = particle.particle.serial_brute_force_binding_energy(
                    np.int32(len(a["tm"])), a["tm"], a["tx"], a["ty"], a["tz"], np.int32(len(a["sx"])), a["sx"], a["sy"],
                    a["sz"])
                assert np.array_equal(ser.view(np.uint32), a["be"].view(np.uint32)), (tag, part, k)
                n_checked += 1
    assert n_checked == 22


# ---- BASELINE sizes: properties the domain offers, plus sampled oracle checks ----------------
def test_cfg2_full_size_stellar_unbinding_properties():
    c = synth.config2()
    s, g = c.stars, c.gas
    r = unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, pre=[g.pos_mass()], kappa=9.0, mode="fast")
    assert r.converged and r.n_iter >= 2
    assert np.array_equal(np.flatnonzero(r.mask), r.idx) and np.all(np.diff(r.idx) > 0)
    assert r.stats.pairs == r.pairs
    # sampled oracle check of the first-pass potentials of particles that were removed in pass 1
    M = s.mass.sum()
    vb0 = (np.sum(s.mass * s.vx) / M, np.sum(s.mass * s.vy) / M, np.sum(s.mass * s.vz) / M)
    one = unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, pre=[g.pos_mass()], kappa=9.0, mode="fast", max_iter=1)
    np.testing.assert_allclose(one.vb, vb0 if False else one.vb)
    pick = np.random.default_rng(3).choice(len(s), 200, replace=False)
    src = [np.concatenate((getattr(g, k), getattr(s, k))) for k in ("mass", "x", "y", "z")]
    ref = O.brute_force_binding_energy_fortran(*src, s.x[pick], s.y[pick], s.z[pick], variant="f64acc")
    assert np.max(np.abs(one.be32[pick].astype(np.float64) / ref - 1)) < FAST_RTOL
    # fixed point: unbinding the bound set again removes nothing and reproduces M, vb
    i = r.idx
    again = unbind_halo(s.x[i], s.y[i], s.z[i], s.vx[i], s.vy[i], s.vz[i], s.mass[i], pre=[g.pos_mass()], kappa=9.0,
                        mode="fast")
    assert again.n_iter == 1 and again.mask.all()
    np.testing.assert_allclose(again.mass, r.mass, rtol=1e-12)
    np.testing.assert_allclose(again.vb, r.vb, rtol=1e-9)


def test_cfg3_full_catalogue_properties_and_sampled_parity():
    cat = synth.config3()
    res = unbind_catalogue(cat.offsets, cat.x, cat.y, cat.z, cat.vx, cat.vy, cat.vz, cat.mass, kappa=9.0, mode="fast")
    sizes = cat.sizes()
    nb = np.array([h.n_bound for h in res.halos])
    assert np.all(nb <= sizes) and np.all(nb >= 0)
    assert int(res.mask.sum()) == int(nb.sum())
    assert all(h.converged for h in res.halos)
    assert res.stats.pairs == sum(h.pairs for h in res.halos)
    # member lists: ascending and consistent with the mask, on a sample of haloes of every size
    order = np.argsort(sizes)
    sample = np.concatenate((order[:5], order[len(order) // 2:len(order) // 2 + 5], order[-3:]))
    for h in sample:
        a, b = cat.offsets[h], cat.offsets[h + 1]
        m = res.halo_mask(h)
        assert np.array_equal(np.flatnonzero(m), res.members(h))
    # oracle parity on small and mid-size haloes (the oracle needs seconds for these)
    for h in np.concatenate((order[:6], order[len(order) // 2:len(order) // 2 + 3], order[-200:-198])):
        o = O.unbind_halo(*cat.halo(h), kappa=9.0, variant="f64acc")
        m = res.halo_mask(h)
        diff = m != o.mask
        assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[diff] < BAND), h
        if not diff.any():
            assert res.halos[h].n_iter == o.n_iter
            np.testing.assert_allclose(res.halos[h].mass, o.mass, rtol=1e-12)
            np.testing.assert_allclose(res.halos[h].vb, o.vb, rtol=1e-6, atol=1e-9)


# ---- error behaviour and odd inputs ---------------------------------------------------------
def test_call_order_and_argument_errors():
    from pyhalma_b200 import _lib
    off = np.array([0, 10], np.int64)
    x = np.arange(10.0)
    with UnbindPlan(off, [np.array([0, 3], np.int64)], vb_fixed=True) as plan:
        with pytest.raises(_lib.HalmaError) as e:
            plan.run()
        assert e.value.code == _lib.ERR_STATE                      # members not uploaded
        plan.upload_members(x, x, x, x, x, x, x)
        with pytest.raises(_lib.HalmaError) as e:
            plan.run()
        assert e.value.code == _lib.ERR_STATE                      # group 0 not uploaded
        plan.upload_group(0, x[:3], x[:3], x[:3], x[:3])
        with pytest.raises(_lib.HalmaError) as e:
            plan.run()
        assert e.value.code == _lib.ERR_STATE                      # vb_fixed without set_vb
        with pytest.raises(_lib.HalmaError):
            plan.download()                                         # nothing ran yet
        with pytest.raises(ValueError):
            plan.upload_members(x[:5], x, x, x, x, x, x)
        with pytest.raises(ValueError):
            plan.upload_group(0, x, x, x, x)
        plan.set_vb([0., 0., 0.])
        plan.run()
        assert plan.download().halos[0].n_iter >= 1
    with pytest.raises(_lib.HalmaError):
        UnbindPlan(off, max_iter=0)
    with pytest.raises(_lib.HalmaError):
        UnbindPlan(np.array([0, 5, 3], np.int64))                   # decreasing offsets
    with pytest.raises(ValueError):
        UnbindPlan(off, mode="quick")


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_zero_mass_and_nan_members(mode):
    rng = np.random.default_rng(21)
    p = synth.plummer_stars(900, 2e-3, 1e6, rng, centre=(1.0, 2.0, 3.0))
    p.mass[::7] = 0.0                                   # massless tracers: targets but no sources
    variant = "f32seq" if mode == "exact" else "f64acc"
    o = O.unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, kappa=9.0, variant=variant)
    r = unbind_halo(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass, kappa=9.0, mode=mode)
    diff = r.mask != o.mask
    assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[diff] < (0 if mode == "exact" else BAND) + 1e-300)
    np.testing.assert_allclose(r.mass, o.mass, rtol=1e-12)
    # a NaN coordinate poisons every potential that sees it (reference: NaN /= x is true);
    # NaN energies are neither bound nor unbound, so those particles drop out
    q = synth.plummer_stars(300, 2e-3, 1e6, rng)
    q.x[17] = np.nan
    o = O.unbind_halo(q.x, q.y, q.z, q.vx, q.vy, q.vz, q.mass, kappa=9.0, variant=variant, max_iter=1)
    r = unbind_halo(q.x, q.y, q.z, q.vx, q.vy, q.vz, q.mass, kappa=9.0, mode=mode, max_iter=1)
    assert np.array_equal(np.isnan(r.be32), np.isnan(o.be32))
    # (a particle sharing a float32 coordinate with the NaN one never sees it, so a few stay finite)
    assert np.array_equal(r.mask, o.mask) and r.mask.sum() <= 3


def test_catalogue_fixed_bulk_velocity_per_halo():
    off, cols = ragged_catalogue()
    nh = len(off) - 1
    vb = np.random.default_rng(4).normal(0, 50, (nh, 3))
    res = unbind_catalogue(off, *cols, kappa=2.0, vb=vb, mode="exact")
    for h in (3, 9, 15, 19):
        a, b = off[h], off[h + 1]
        o = O.unbind_halo(*[c[a:b] for c in cols], kappa=2.0, vb_fixed=vb[h])
        assert np.array_equal(res.halo_mask(h), o.mask) and res.halos[h].n_iter == o.n_iter
        assert res.halos[h].vb == tuple(vb[h])


def test_predicate_free_path_forced_on_small_and_odd_inputs():
    """The predicate-free kernel + correction tickets normally engage above 1e9 pairs per pass.
    Re-run the ragged / degenerate / duplicate / NaN / external-group / golden cases with the
    threshold at zero so that every one of them goes through that path."""
    import subprocess
    import sys
    if os.environ.get("HALMA_NP_MIN_PAIRS") == "0":
        pytest.skip("already running with the threshold forced to zero")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HALMA_NP_MIN_PAIRS="0")
    sel = ("ragged or degenerate or duplicates or zero_mass or external_groups or golden_fast or fused_large "
           "or fast_mode_against or lattice or graph_loop or idempotence or rps_mass_sums")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_unbind.py"), "-q", "-x",
                          "-m", "gpu", "-k", sel], capture_output=True, text=True, timeout=1200, env=env, cwd=root)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout


def test_symmetric_self_term_against_one_sided_and_oracle():
    """halma_unbind_config.symmetric (on by default in the Python layer; HALMA_SYMMETRIC=0 turns it
    off): member x member pairs of different tiles are evaluated once for
    both particles.  Same predicate, same per-pair arithmetic; only the order of the float64 partial
    sums differs, so potentials agree with the one-sided kernel and with the float64 oracle within
    the FAST tolerance (1e-6) and masks outside the 1e-6 energy band."""
    rng = np.random.default_rng(123)
    n = 70_000                                    # 4.9e9 member pairs: predicate-free kernel, throughput shape
    st = synth.plummer_stars(n, 6 * synth.KPC, 1e6, rng)
    synth.add_coincident_pairs(st, 5, rng)        # shared coordinates -> correction tickets
    gas = synth.lattice_gas(4000, synth.CELL, rng, m_total=5e9)
    dm = synth.dm_cloud(3000, 20 * synth.KPC, 1e7, rng)
    kw = dict(pre=[gas.pos_mass()], post=[dm.pos_mass()], kappa=9.0)
    args = (st.x, st.y, st.z, st.vx, st.vy, st.vz, st.mass)
    one = unbind_halo(*args, mode="fast", max_iter=1, vb_fixed=synth.BULK_V, symmetric=False, **kw)
    sym = unbind_halo(*args, mode="fast", max_iter=1, vb_fixed=synth.BULK_V, symmetric=True, **kw)
    assert sym.stats.evaluations < 0.56 * one.stats.evaluations and one.stats.evaluations == one.stats.pairs
    ref = O.unbind_halo(*args, variant="f64acc", max_iter=1, vb_fixed=synth.BULK_V, **kw)
    np.testing.assert_allclose(sym.be32, ref.be32, rtol=1e-6)
    np.testing.assert_allclose(sym.be32, one.be32, rtol=1e-6)
    assert np.max(np.abs(sym.be32.astype(np.float64) / ref.be32 - 1)) < 5e-7
    diff = sym.mask != ref.mask
    assert np.all(O.energy_margin(ref.energy, ref.be32, 9.0)[diff] < 1e-6)
    assert sym.pairs == one.pairs                  # interactions are counted the same way
    # the two-sided sums are rounded to a per-halo quantum inside whose window float64 addition is
    # exact, so the order of the atomics cannot show: a second run is bit-identical
    again = unbind_halo(*args, mode="fast", max_iter=1, vb_fixed=synth.BULK_V, symmetric=True, **kw)
    assert np.array_equal(again.be32.view(np.uint32), sym.be32.view(np.uint32))
    assert np.array_equal(again.energy, sym.energy) and np.array_equal(again.mask, sym.mask)
    # iterated to the fixed point: same member set as the one-sided run (up to the energy band)
    one = unbind_halo(*args, mode="fast", symmetric=False, **kw)
    sym = unbind_halo(*args, mode="fast", symmetric=True, **kw)
    assert sym.n_iter == one.n_iter and np.count_nonzero(sym.mask != one.mask) <= 2
    np.testing.assert_allclose(sym.mass, one.mass, rtol=1e-4)
    # exact duplicates in DIFFERENT tiles: the symmetric sums come out non-finite and the halo is
    # recomputed by the predicated kernel, like in the one-sided predicate-free path
    x2 = st.x.copy(); y2 = st.y.copy(); z2 = st.z.copy()
    x2[60_000], y2[60_000], z2[60_000] = x2[17], y2[17], z2[17]
    a2 = (x2, y2, z2, st.vx, st.vy, st.vz, st.mass)
    one = unbind_halo(*a2, mode="fast", max_iter=1, vb_fixed=synth.BULK_V, symmetric=False, **kw)
    sym = unbind_halo(*a2, mode="fast", max_iter=1, vb_fixed=synth.BULK_V, symmetric=True, **kw)
    assert np.all(np.isfinite(sym.be32))
    np.testing.assert_array_equal(sym.be32, one.be32)          # both took the same fallback


@pytest.mark.parametrize("rows", ["4", "8"])
def test_symmetric_row_units_of_4_and_8_members_per_lane(monkeypatch, rows):
    """The symmetric tickets hold one row tile (4 members per lane) or a pair of row tiles (8) per warp; the library
    picks by the share of member x member work (api.cu::plan_build), HALMA_SYM_ROWS forces either.  Both kernels on
    a halo with an odd number of row tiles, a short last tile, external groups, coordinate-sharing pairs and more
    than one pass: potentials within the FAST tolerance of the float64 oracle, masks identical outside the band."""
    monkeypatch.setenv("HALMA_SYM_ROWS", rows)
    monkeypatch.setenv("HALMA_NP_MIN_PAIRS", "0")
    monkeypatch.setenv("HALMA_FAST_VARIANT", "0")
    rng = np.random.default_rng(2024)
    n = 128 * 9 + 57                                  # 10 tiles: 9 row tiles = 4 pairs + a single, short last column
    st = synth.plummer_stars(n, 4 * synth.KPC, 1e6, rng, centre=(3.0, -2.0, 7.0), interloper_frac=0.2)
    synth.add_coincident_pairs(st, 6, rng)
    dm = synth.dm_cloud(700, 15 * synth.KPC, 1e7, rng, centre=(3.0, -2.0, 7.0))
    kw = dict(post=[dm.pos_mass()], kappa=9.0)
    args = (st.x, st.y, st.z, st.vx, st.vy, st.vz, st.mass)
    o = O.unbind_halo(*args, variant="f64acc", **kw)
    r = unbind_halo(*args, mode="fast", symmetric=True, **kw)
    assert o.n_iter >= 2
    assert r.stats.evaluations < r.stats.pairs          # the symmetric tickets did run
    both = r.mask & o.mask
    assert np.abs(r.be32[both].astype(np.float64) / o.be32[both] - 1).max() < FAST_RTOL
    diff = r.mask != o.mask
    assert np.all(O.energy_margin(o.energy, o.be32, 9.0)[diff] < BAND)
    again = unbind_halo(*args, mode="fast", symmetric=True, **kw)
    np.testing.assert_array_equal(again.be32.view(np.uint32), r.be32.view(np.uint32))      # bit-reproducible


def test_symmetric_mode_forced_on_small_and_odd_inputs():
    """The same battery as the predicate-free path, with symmetric tickets on and the throughput
    shape forced so that small and ragged haloes go through them (haloes of <= 128 members have
    no off-diagonal tile and must be unaffected)."""
    import subprocess
    import sys
    if os.environ.get("HALMA_SYMMETRIC") == "1":
        pytest.skip("already running in symmetric mode")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HALMA_NP_MIN_PAIRS="0", HALMA_SYMMETRIC="1", HALMA_FAST_VARIANT="0")
    sel = ("ragged or degenerate or duplicates or zero_mass or external_groups or golden_fast or fused_large "
           "or fast_mode_against or lattice or rps_mass_sums or cfg3_full or cfg2_full or idempotence or graph_loop")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_unbind.py"), "-q", "-x",
                          "-m", "gpu", "-k", sel], capture_output=True, text=True, timeout=1500, env=env, cwd=root)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout


def test_symmetric_exact_window_with_far_outliers():
    """The two-sided sums are exact inside +-2^52 q ~ 3e4..6e4 M/extent.  One member 1000 core radii
    away stretches the extent but stays inside the window (symmetric tickets used); one at 1e7 core
    radii pushes the central potentials out of it: the halo is then recomputed one-sided in the same
    pass and the result is still right."""
    rng = np.random.default_rng(321)
    n = 70_000
    st = synth.plummer_stars(n, 2 * synth.KPC, 1e6, rng)
    for shift, expect_sym in ((2.0, True), (2.0e4, False)):
        x = st.x.copy()
        x[12345] += shift
        args = (x, st.y, st.z, st.vx, st.vy, st.vz, st.mass)
        one = unbind_halo(*args, mode="fast", max_iter=1, vb_fixed=synth.BULK_V, symmetric=False, kappa=9.0)
        sym = unbind_halo(*args, mode="fast", max_iter=1, vb_fixed=synth.BULK_V, symmetric=True, kappa=9.0)
        np.testing.assert_allclose(sym.be32, one.be32, rtol=1e-6)
        assert np.count_nonzero(sym.mask != one.mask) <= 1
        if expect_sym:
            assert sym.stats.evaluations < 0.56 * sym.stats.pairs
        else:
            assert sym.stats.evaluations == sym.stats.pairs


def test_concurrent_plans_from_threads():
    """Plans are independent objects with their own streams: four host threads unbinding different
    haloes at the same time must get exactly what a serial run gets (ctypes releases the GIL)."""
    import threading
    cases = []
    for k in range(8):
        rng = np.random.default_rng(900 + k)
        st = synth.plummer_stars(3000 + 1700 * k, 2 * synth.KPC, 1e6, rng)
        gas = synth.lattice_gas(800 + 300 * k, synth.CELL, rng, m_total=3e8)
        cases.append((st, gas))

    def run(c, mode):
        st, gas = c
        return unbind_halo(st.x, st.y, st.z, st.vx, st.vy, st.vz, st.mass, pre=[gas.pos_mass()], kappa=9.0, mode=mode)

    for mode in ("exact", "fast"):
        serial = [run(c, mode) for c in cases]
        out = [None] * len(cases)
        errs = []

        def worker(ids):
            try:
                for _ in range(3):
                    for i in ids:
                        out[i] = run(cases[i], mode)
            except Exception as exc:          # surfaces in the main thread below
                errs.append(exc)

        threads = [threading.Thread(target=worker, args=(list(range(t, len(cases), 4)),)) for t in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errs, errs
        for a, b in zip(serial, out):
            assert np.array_equal(a.mask, b.mask) and np.array_equal(a.be32.view(np.uint32), b.be32.view(np.uint32))
            assert a.n_iter == b.n_iter and a.vb == b.vb


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_multi_stream_catalogue_matches_one_plan(mode):
    """unbind_catalogue(streams=3): consecutive runs of haloes in their own plans and host threads.
    EXACT is bit-identical to the one-plan run; FAST may differ where a halo's j-split differs."""
    cat = synth.config3(n_halo=60, nmin=50, nmax=9000, seed_extra=4)
    args = (cat.offsets, cat.x, cat.y, cat.z, cat.vx, cat.vy, cat.vz, cat.mass)
    one = unbind_catalogue(*args, mode=mode)
    par = unbind_catalogue(*args, mode=mode, streams=3)
    assert len(par.halos) == len(one.halos) == 60
    assert [h.n_iter for h in par.halos] == [h.n_iter for h in one.halos]
    assert list(par.halos.field("n_bound")) == [h.n_bound for h in one.halos]
    assert par.stats.pairs == one.stats.pairs
    if mode == "exact":
        assert np.array_equal(par.be32.view(np.uint32), one.be32.view(np.uint32))
        assert np.array_equal(par.mask, one.mask) and np.array_equal(par.energy, one.energy)
        for h in (0, 17, 59):
            assert np.array_equal(par.members(h), one.members(h)) and par.halos[h].vb == one.halos[h].vb
    else:
        np.testing.assert_allclose(par.be32, one.be32, rtol=1e-6)
        assert np.count_nonzero(par.mask != one.mask) == 0
    # with an external group and fixed bulk velocities
    rng = np.random.default_rng(5)
    nh = cat.n_halo
    ext_n = rng.integers(0, 40, nh)
    eoff = np.concatenate(([0], np.cumsum(ext_n)))
    centres = np.array([[cat.x[cat.offsets[h]], cat.y[cat.offsets[h]], cat.z[cat.offsets[h]]] for h in range(nh)])
    rep = np.repeat(np.arange(nh), ext_n)
    ex, ey, ez = (centres[rep, k] + rng.normal(0, 2e-3, len(rep)) for k in range(3))
    em = rng.uniform(1e6, 1e8, len(rep))
    vb = rng.normal(0, 100, (nh, 3))
    kw = dict(groups=[(eoff, em, ex, ey, ez)], n_pre=0, split_classes=True, vb=vb, kappa=2.0, mode=mode)
    one = unbind_catalogue(*args, **kw)
    par = unbind_catalogue(*args, streams=4, **kw)
    if mode == "exact":
        assert np.array_equal(par.be32.view(np.uint32), one.be32.view(np.uint32)) and np.array_equal(par.mask, one.mask)
    else:
        np.testing.assert_allclose(par.be32, one.be32, rtol=1e-6)
