"""Host-side multi-GPU logic on CPU: LPT cost sharding and the world_size-2 gather (gloo)."""
import os
import socket

import numpy as np
import pytest

from pyhalma_b200 import sharding, synth


def test_lpt_partition_properties():
    rng = np.random.default_rng(0)
    sizes = synth.powerlaw_sizes(10_000, 100, 100_000, 1.9, rng)
    off = np.concatenate(([0], np.cumsum(sizes)))
    costs = sharding.halo_costs(off)
    assert costs.shape == (10_000,) and np.all(costs == sizes.astype(float) ** 2)
    for n in (1, 2, 4, 8):
        parts = sharding.lpt_partition(costs, n)
        allids = np.concatenate(parts)
        assert len(allids) == 10_000 and len(np.unique(allids)) == 10_000
        assert all(np.all(np.diff(p) > 0) for p in parts)
        # near-linear scaling needs a balanced split (largest halo < total/8)
        assert sharding.partition_imbalance(costs, parts) < 1.02, n
    assert sharding.lpt_partition(costs, 4)[2].tolist() == sharding.lpt_partition(costs.copy(), 4)[2].tolist()


def test_lpt_edge_cases():
    assert [p.tolist() for p in sharding.lpt_partition([], 3)] == [[], [], []]
    parts = sharding.lpt_partition([5.0], 4)
    assert sum(len(p) for p in parts) == 1
    parts = sharding.lpt_partition([1, 1, 1, 1], 2)
    assert sorted(len(p) for p in parts) == [2, 2]
    costs = sharding.halo_costs([0, 10, 10, 14], [[0, 5, 5, 6]])
    assert costs.tolist() == [10 * 15, 0, 4 * 5]


def test_take_haloes_and_split_owner():
    off = np.array([0, 3, 3, 7, 9])
    x = np.arange(9.0)
    o2, (x2,) = sharding.take_haloes(off, [x], [2, 0])
    assert o2.tolist() == [0, 4, 7] and x2.tolist() == [3, 4, 5, 6, 0, 1, 2]
    o3, (x3,) = sharding.take_haloes(off, [x], [])
    assert o3.tolist() == [0] and len(x3) == 0
    own = sharding.split_owner(1000, 128, 3)
    assert own[0] == 0 and own[127] == 0 and own[128] == 1 and own[256] == 2 and own[384] == 0
    counts = np.bincount(sharding.split_owner(2_000_000, 128, 8), minlength=8)
    assert counts.max() - counts.min() <= 128


def _worker(rank, world, port, n_halo, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)                 # same catalogue on every rank
        sizes = synth.powerlaw_sizes(n_halo, 10, 5000, 1.9, rng)
        off = np.concatenate(([0], np.cumsum(sizes)))
        parts = sharding.lpt_partition(sharding.halo_costs(off), world)
        mine = parts[rank]
        sub_off, _ = sharding.take_haloes(off, [], mine)
        # stand-in for the per-rank GPU run: rows = (size, size^2, rank)
        n = np.diff(sub_off).astype(float)
        rows = np.stack([n, n * n, np.full(len(n), float(rank))], axis=1) if len(n) else np.zeros((0, 3))
        full = sharding.gather_catalogue(mine, rows, n_halo)
        if rank == 0:
            ok = (full is not None and np.array_equal(full[:, 0], sizes.astype(float))
                  and set(np.unique(full[:, 2])) == set(range(world)))
            q.put(bool(ok))
        else:
            q.put(full is None)
    finally:
        dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 200, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def _split_worker(rank, world, port, q):
    """Emulates the split-mode protocol of csrc/api.cu with the CPU oracle: every rank holds
    the whole halo, evaluates only the target groups it owns, and the potentials are summed
    over ranks (all-reduce).  Each element has exactly one non-zero contributor, so the sum
    must equal the single-process result bit for bit."""
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = synth.config1(1500, 700, n_dm=100)
        s, g = c.stars, c.gas
        src = [np.concatenate((getattr(g, k), getattr(s, k))) for k in ("mass", "x", "y", "z")]
        own = sharding.split_owner(len(s), 128, world) == rank
        phi = np.zeros(len(s), np.float64)
        if own.any():
            phi[own] = O.brute_force_binding_energy_fortran(*src, s.x[own], s.y[own], s.z[own], variant="f64acc")
        t = torch.from_numpy(phi)
        dist.all_reduce(t)
        full = O.brute_force_binding_energy_fortran(*src, s.x, s.y, s.z, variant="f64acc")
        q.put(bool(np.array_equal(t.numpy(), full)) and bool(own.sum() > 0))
    finally:
        dist.destroy_process_group()


def test_split_protocol_world_size_2_gloo():
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def _split_incremental_worker(rank, world, port, q):
    """Emulates the split-mode LOOP with incremental passes (csrc/api.cu::enqueue_pass) with the
    CPU oracle: every rank keeps the whole halo and the kept potentials (kernels 2-3 run replicated), evaluates
    only the target groups it owns -- against all current sources in a full pass, against the members the last
    pass removed in an incremental one -- and ONE all-reduce per pass completes the sums.  No other exchange: the
    full / incremental decision depends on replicated state only, so every rank must take the same one, and the
    result must be the single-process unbinding."""
    import hashlib
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = synth.config1(1500, 700, n_dm=100)
        s, g, d = c.stars, c.gas, c.dm
        ext = [np.concatenate((getattr(g, k), getattr(d, k))) for k in ("mass", "x", "y", "z")]
        N = len(s)
        idx = np.arange(N)
        keep = np.zeros(N)
        be_out = np.zeros(N, np.float32)
        removed = np.zeros(0, np.int64)
        incr, kinds, it = False, [], 0
        while len(idx) and it < 64:
            own = sharding.split_owner(len(idx), 128, world) == rank        # groups of the CURRENT member list
            if incr:
                src = [getattr(s, k)[removed] for k in ("mass", "x", "y", "z")]
            else:
                src = [np.concatenate((getattr(s, k)[idx], e)) for k, e in zip(("mass", "x", "y", "z"), ext)]
            part = np.zeros(len(idx))
            if own.any():
                t = idx[own]
                part[own] = O.brute_force_binding_energy_fortran(*src, s.x[t], s.y[t], s.z[t], variant="f64acc")
            red = torch.from_numpy(part)
            dist.all_reduce(red)                                            # the one collective of the pass
            phi = keep[idx] - red.numpy() if incr else red.numpy()
            keep[idx] = phi
            kinds.append("incr" if incr else "full")
            be = phi.astype(np.float32)
            M = O.total_mass(idx, s.mass)
            vb = O.CM_velocity(M, idx, s.vx, s.vy, s.vz, s.mass)
            E = O.energy_step(be, s.vx[idx], s.vy[idx], s.vz[idx], vb[0], vb[1], vb[2], 9.0)
            be_out[idx] = be
            it += 1
            bound = E <= 0.0
            removed, new = idx[~bound], idx[bound]
            incr = len(removed) > 0 and 2 * len(removed) <= len(new)
            done = len(new) == len(idx)
            idx = new
            if done:
                break
        ref = O.unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0,
                            variant="f64acc")
        mask = np.zeros(N, bool)
        mask[idx] = True
        seen = ref.be32 > 0
        ok = (np.array_equal(mask, ref.mask) and it == ref.n_iter and "incr" in kinds and kinds[0] == "full"
              and float(np.abs(be_out[seen].astype(np.float64) / ref.be32[seen] - 1).max()) < 1e-6)
        # every rank holds the same result, bit for bit
        digest = hashlib.sha256(mask.tobytes() + be_out.tobytes() + repr(kinds).encode()).digest()
        mine = torch.frombuffer(bytearray(digest), dtype=torch.uint8).clone()
        others = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(others, mine)
        q.put(bool(ok) and all(bool(torch.equal(o, mine)) for o in others))
    finally:
        dist.destroy_process_group()


def test_split_loop_with_incremental_passes_world_size_2_gloo():
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_split_incremental_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def test_cost_cuts_for_the_multi_stream_catalogue_path():
    from pyhalma_b200.unbind import cost_cuts
    sizes = np.array([10, 10, 1000, 10, 10, 10, 500, 500, 10, 10])
    off = np.concatenate(([0], np.cumsum(sizes)))
    cuts = cost_cuts(off, [], 3)
    assert cuts[0] == 0 and cuts[-1] == len(sizes) and all(b > a for a, b in zip(cuts, cuts[1:])) and len(cuts) <= 4
    cost = sizes.astype(float) ** 2
    parts = [cost[a:b].sum() for a, b in zip(cuts, cuts[1:])]
    assert max(parts) <= 1.0 * cost.sum()          # the giant halo bounds any partition of consecutive runs
    assert cost_cuts(off, [], 1) == [0, len(sizes)]
    # externals count: a halo with many external sources is as expensive as a bigger one
    ext = np.concatenate(([0], np.cumsum(np.where(np.arange(10) == 0, 10_000_000, 0))))
    assert cost_cuts(off, [ext], 2)[1] == 1
    # equal haloes split evenly
    off = np.arange(0, 1201, 100)
    assert cost_cuts(off, [], 4) == [0, 3, 6, 9, 12]
    # tapered: the first and the last run carry half the cost of the others (short upload head / download tail)
    assert cost_cuts(off, [], 4, taper=True) == [0, 2, 6, 10, 12] and cost_cuts(off, [], 2, taper=True) == cost_cuts(off, [], 2)


def test_auto_parts_of_the_one_shot_catalogue_call():
    from pyhalma_b200.unbind import auto_parts
    assert auto_parts(0) == 1
    assert auto_parts(1_110_367) == 1          # an eighth of cfg3: 62 MB, one plan
    assert auto_parts(4_500_000) == 2
    assert auto_parts(9_062_762) == 3          # the cfg3 catalogue on one GPU: 507 MB
    assert auto_parts(10 ** 9) == 3            # never more: the parts' fixed costs add up
    assert auto_parts(1_000_000, 6_000_000) >= auto_parts(1_000_000)
