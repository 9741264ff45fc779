"""Split mode (one halo shared by all ranks, one NCCL all-reduce per pass) against the
single-GPU run of the same halo: results must be bit-identical (with the same settings: the
external-sum cache is a single-GPU feature, so the reference run has it off; incremental passes
run in both).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/split_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from pyhalma_b200 import synth
from pyhalma_b200.unbind import Communicator, UnbindPlan, nccl_unique_id, unbind_halo

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
uid0 = [nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid0, src=0)
shared = Communicator(uid0[0], rank, world, local)          # reused by the last case
case_no = 0
for mode, n_star, n_gas in (("exact", 6000, 3000), ("fast", 30000, 20000), ("fast", 90000, 40000)):
    c = synth.config1(n_star, n_gas, n_dm=500)
    s, g, d = c.stars, c.gas, c.dm
    uid = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    off = np.array([0, len(s)], np.int64)
    eoff = [np.array([0, len(g)], np.int64), np.array([0, len(d)], np.int64)]
    with UnbindPlan(off, eoff, mode=mode, n_pre=1, kappa=9.0, device=local, rank=rank, n_ranks=world) as plan:
        if case_no == 2:
            plan.use_comm(shared)
        else:
            plan.join(uid[0])
        case_no += 1
        plan.upload_members(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass)
        plan.upload_group(0, g.mass, g.x, g.y, g.z)
        plan.upload_group(1, d.mass, d.x, d.y, d.z)
        st = plan.run()
        res = plan.download()
    single = unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, pre=[g.pos_mass()], post=[d.pos_mass()],
                         kappa=9.0, mode=mode, device=local, cache_external=False)
    same = (np.array_equal(res.mask.astype(bool), single.mask)
            and np.array_equal(res.be32.view(np.uint32), single.be32.view(np.uint32))
            and np.array_equal(res.energy, single.energy) and res.halos[0].n_iter == single.n_iter
            and res.halos[0].vb == single.vb and res.halos[0].pairs == single.pairs)
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("split %s: n=%d passes=%d bound=%d identical_on_all_ranks=%s  (split %.2f ms, single %.2f ms)" % (
            mode, len(s), res.halos[0].n_iter, res.halos[0].n_bound, bool(flag.item()), st.total_ms,
            single.stats.total_ms), flush=True)
    ok = ok and bool(flag.item())
# gas layout on a lattice (dense coordinate sharing): correction tickets are distributed over
# the ranks and their planes all-reduced
c = synth.config1(30000, 60000, n_dm=800)
s, g, d = c.stars, c.gas, c.dm
M = s.mass.sum()
vb = np.array([np.sum(s.mass * s.vx), np.sum(s.mass * s.vy), np.sum(s.mass * s.vz)]) / M
off = np.array([0, len(g)], np.int64)
eoff = [np.array([0, len(d)], np.int64), np.array([0, len(s)], np.int64)]
with UnbindPlan(off, eoff, mode="fast", split_classes=True, vb_fixed=True, kappa=2.0, device=local, rank=rank,
                n_ranks=world) as plan:
    plan.use_comm(shared)
    plan.upload_members(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass)
    plan.upload_group(0, d.mass, d.x, d.y, d.z)
    plan.upload_group(1, s.mass, s.x, s.y, s.z)
    plan.set_vb(vb)
    st = plan.run()
    res = plan.download()
single = unbind_halo(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass, post=[d.pos_mass(), s.pos_mass()], split_classes=True,
                     kappa=2.0, vb_fixed=vb, mode="fast", device=local, cache_external=False)
same = (np.array_equal(res.mask.astype(bool), single.mask) and np.array_equal(res.be32.view(np.uint32), single.be32.view(np.uint32))
        and np.array_equal(res.energy, single.energy) and res.halos[0].n_iter == single.n_iter)
flag = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("split gas-lattice fast: n=%d passes=%d bound=%d identical_on_all_ranks=%s" % (
        len(g), res.halos[0].n_iter, res.halos[0].n_bound, bool(flag.item())), flush=True)
ok = ok and bool(flag.item())
shared.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
