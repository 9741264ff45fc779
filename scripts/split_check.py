"""Split mode (one halo shared by all ranks: target groups dealt round-robin, sources replicated, one grouped NCCL
all-reduce per pass) against (a) the single-GPU run of the same halo with the same, default options -- must be
bit-identical -- and (b) the CPU oracle (f32seq bitwise in EXACT mode; f64acc within 1e-6 / masks identical outside
the 1e-6 energy band in FAST mode).  The single-GPU run uses the persistent loop kernel, the split run the
enqueue-ahead driver: the two share their phase code.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/split_check.py [--giant 2000000]
"""
import os
import sys

# torchrun exports OMP_NUM_THREADS=1; the oracle loops (rank 0 only) want the host's cores
os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0))) if int(os.environ.get("RANK", "0")) == 0 else "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from oracle import oracle as O
from pyhalma_b200 import synth
from pyhalma_b200.unbind import Communicator, UnbindPlan, nccl_unique_id, unbind_halo

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid0 = [nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid0, src=0)
shared = Communicator(uid0[0], rank, world, local)
FAST_RTOL, BAND = 1e-6, 1e-6
ok = True


def all_true(flag: bool) -> bool:
    t = torch.tensor([1 if flag else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def run_case(name, mode, members, pre, post, *, split_classes=False, kappa=9.0, vb=None, own_comm=False, oracle=True):
    """members: 7 float64 arrays; pre / post: lists of Particles."""
    global ok
    groups = list(pre) + list(post)
    off = np.array([0, len(members[0])], np.int64)
    eoff = [np.array([0, len(g)], np.int64) for g in groups]
    with UnbindPlan(off, eoff, mode=mode, n_pre=len(pre), split_classes=split_classes, vb_fixed=vb is not None,
                    kappa=kappa, device=local, rank=rank, n_ranks=world) as plan:
        if own_comm:
            uid = [nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            plan.join(uid[0])
        else:
            plan.use_comm(shared)
        plan.upload_members(*members)
        for k, g in enumerate(groups):
            plan.upload_group(k, g.mass, g.x, g.y, g.z)
        if vb is not None:
            plan.set_vb(vb)
        plan.run()
        st = plan.run()                      # plans are re-runnable in split mode too
        res = plan.download()
    kw = dict(pre=[g.pos_mass() for g in pre], post=[g.pos_mass() for g in post], split_classes=split_classes,
              kappa=kappa, vb_fixed=vb)
    single = unbind_halo(*members, mode=mode, device=local, **kw)
    same = (np.array_equal(res.mask.astype(bool), single.mask)
            and np.array_equal(res.be32.view(np.uint32), single.be32.view(np.uint32))
            and np.array_equal(res.energy, single.energy) and res.halos[0].n_iter == single.n_iter
            and res.halos[0].vb == single.vb and res.halos[0].pairs == single.pairs
            and np.array_equal(res.members(0), single.idx) and st.evaluations == single.stats.evaluations)
    same = all_true(same)
    par = "-"
    if oracle and rank != 0:
        par = str(all_true(True))          # the ranks hold identical results (checked above): rank 0 compares
        ok = ok and par == "True"
    elif oracle:
        o = O.unbind_halo(*members, variant="f32seq" if mode == "exact" else "f64acc", **kw)
        mask = res.mask.astype(bool)
        if mode == "exact":
            good = np.array_equal(mask, o.mask) and np.array_equal(res.be32.view(np.uint32), o.be32.view(np.uint32))
        else:
            diff = mask != o.mask
            both = mask & o.mask
            rel = np.abs(res.be32[both].astype(np.float64) / o.be32[both] - 1).max() if both.any() else 0.0
            good = bool(np.all(O.energy_margin(o.energy, o.be32, kappa)[diff] < BAND)) and rel < FAST_RTOL
            if not diff.any():
                good = good and res.halos[0].n_iter == o.n_iter
        par = str(all_true(bool(good)))
        ok = ok and par == "True"
    if rank == 0:
        print("split %s %s: n=%d passes=%d bound=%d identical_on_all_ranks=%s oracle_parity=%s  (split %.2f ms of which "
              "collectives %.2f ms, %.1f MB per pass; single %.2f ms)" % (
                  name, mode, len(members[0]), res.halos[0].n_iter, res.halos[0].n_bound, same, par, st.total_ms,
                  st.comm_ms, st.comm_bytes / max(st.passes, 1) / 1e6, single.stats.total_ms), flush=True)
    ok = ok and same


def star_members(s):
    return (s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass)


# stellar layout with gas before and DM after the members: EXACT, FAST below and above the predicate-free threshold
# (1e9 pairs), the larger one with symmetric tickets, cached external sums and incremental passes in split mode
for k, (mode, n_star, n_gas) in enumerate((("exact", 6000, 3000), ("fast", 30000, 20000), ("fast", 90000, 40000))):
    c = synth.config1(n_star, n_gas, n_dm=500)
    run_case("stellar", mode, star_members(c.stars), [c.gas], [c.dm], own_comm=k < 2)
# gas layout on a lattice (dense coordinate sharing): correction tickets are distributed over
# the ranks and their planes all-reduced; fixed bulk velocity; externals cached
c = synth.config1(30000, 60000, n_dm=800)
s, g, d = c.stars, c.gas, c.dm
M = s.mass.sum()
vb = np.array([np.sum(s.mass * s.vx), np.sum(s.mass * s.vy), np.sum(s.mass * s.vz)]) / M
run_case("gas-lattice", "fast", (g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass), [], [d, s], split_classes=True, kappa=2.0, vb=vb)
# BASELINE configs[3]: the cluster-scale stellar halo (no oracle loop at this size: first-pass potentials of a sample)
if "--giant" in sys.argv:
    n = int(sys.argv[sys.argv.index("--giant") + 1])
    p = synth.config4(n)
    run_case("giant", "fast", star_members(p), [], [], oracle=False)
shared.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
