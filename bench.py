#!/usr/bin/env python
"""Benchmark of the unbinding hot path (BASELINE.json metric: unbinding Ginteractions/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

A step is one full pass of the hot path over one batch of synthetic input.  Default
workload = BASELINE.json configs[1] ("cfg2"): one NFW galaxy halo with 2e5 stars and 5e5
gas cells per GPU; a step = iterative stellar unbinding (sources gas + stars) followed by
iterative gas unbinding (sources gas + stars, fixed stellar bulk velocity).  With N > 1
every rank owns its own halo (independent units, no data-path collective: weak scaling).

One interaction = one (target, source) pair visited in one potential pass; excluded pairs
count.

  value     whole-job Ginteractions/s, inputs resident in HBM, CUDA-event time of the runs
  e2e       the same through the public one-shot API with pinned HOST buffers: H2D of every
            input and D2H of mask + potentials + energies + member lists inside the timing
  roofline  the potential kernel alone against the MUFU.RSQ issue roofline
            (16 rsqrt / clk / SM nominal; the rate and the SM clock are measured), in 1/r
            EVALUATIONS per second.  With --symmetric 1 (default) a member pair in different tiles
            is evaluated once and serves both particles, so there are fewer evaluations than
            interactions; `one_sided` repeats the measurement with --symmetric 0, where the two
            are the same number
  cpu_baseline  the C/OpenMP oracle (a port of the Fortran kernel, oracle/) on the host cores

--impl reference times that CPU port alone, with all host threads, on a bounded sample of
the same workload (the reference's Fortran cannot be compiled in this image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unbinding Ginteractions/s"
UNIT = "Ginteractions/s"


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def make_workload(name: str, rank: int, world: int = 1):
    """Returns (jobs, desc, scaling).  A job describes one plan: CSR offsets, member arrays,
    external groups (ext_offsets, mass, x, y, z), layout keywords, and whether the halo is
    shared by all ranks (split mode)."""
    from pyhalma_b200 import sharding, synth

    def single(n):
        return np.array([0, n], np.int64)

    def star_gas_jobs(case):
        s, g, d = case.stars, case.gas, case.dm
        M = float(np.sum(s.mass))
        vb = np.array([np.sum(s.mass * s.vx), np.sum(s.mass * s.vy), np.sum(s.mass * s.vz)]) / M
        grp = lambda p: (single(len(p)), p.mass, p.x, p.y, p.z)  # noqa: E731
        jobs = [dict(kind="stellar", offsets=single(len(s)), members=(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass),
                     groups=[grp(g)] + ([grp(d)] if len(d) else []), split=False,
                     kw=dict(n_pre=1, split_classes=False, vb=None, kappa=case.factor_v ** 2)),
                dict(kind="gas", offsets=single(len(g)), members=(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass),
                     groups=([grp(d)] if len(d) else []) + [grp(s)], split=False,
                     kw=dict(n_pre=0, split_classes=True, vb=vb, kappa=2.0))]
        return jobs

    if name in ("cfg2", "cfg1"):
        case = synth.config2(seed_extra=rank) if name == "cfg2" else synth.config1(seed_extra=rank)
        what = ("cfg2: NFW galaxy halo, 2e5 stars + 5e5 gas cells per GPU" if name == "cfg2"
                else "cfg1: Plummer halo, 1e4 stars + 1e4 gas cells per GPU")
        desc = {"workload": what + ", iterative stellar + gas unbinding", "n_star": len(case.stars),
                "n_gas": len(case.gas), "parallelism": "1 halo per GPU, no collective"}
        return star_gas_jobs(case), desc, "weak"
    if name == "cfg3":
        cat = synth.config3()                       # the same catalogue on every rank
        parts = sharding.lpt_partition(sharding.halo_costs(cat.offsets), world)
        off, cols = sharding.take_haloes(cat.offsets, [cat.x, cat.y, cat.z, cat.vx, cat.vy, cat.vz, cat.mass],
                                         parts[rank])
        desc = {"workload": "cfg3: catalogue of 1e4 Plummer haloes, N = 1e2..1e5 (dN/dN ~ N^-1.9), batched launch, "
                            "iterative stellar unbinding", "n_halo": cat.n_halo, "sum_n": int(cat.offsets[-1]),
                "sum_n2": cat.meta["sum_n2"], "halos_this_rank": int(len(parts[rank])),
                "lpt_imbalance": sharding.partition_imbalance(sharding.halo_costs(cat.offsets), parts),
                "parallelism": "haloes LPT-sharded by N^2 over %d GPU(s), no collective" % world}
        jobs = [dict(kind="catalogue", offsets=off, members=tuple(cols), groups=[], split=False,
                     kw=dict(n_pre=0, split_classes=False, vb=None, kappa=9.0))]
        return jobs, desc, "strong"
    if name == "cfg4":
        n = int(os.environ.get("HALMA_CFG4_N", "2000000"))
        p = synth.config4(n)                        # the same halo on every rank
        desc = {"workload": "cfg4: one cluster-scale stellar halo, N = %d, targets split over the GPUs, one NCCL "
                            "all-reduce of the potentials per pass" % n, "n_star": n,
                "parallelism": "target groups round-robin over %d GPU(s), sources replicated" % world}
        jobs = [dict(kind="giant", offsets=single(n), members=(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass), groups=[],
                     split=world > 1, kw=dict(n_pre=0, split_classes=False, vb=None, kappa=9.0))]
        return jobs, desc, "strong"
    if name == "cfg5":
        scale = float(os.environ.get("HALMA_CFG5_SCALE", "1.0"))
        case = synth.config5(int(1e7 * scale), int(2e6 * scale), int(5e5 * scale))    # the same on every rank
        s, g, d = case.stars, case.gas, case.dm
        M = float(np.sum(s.mass))
        vb = np.array([np.sum(s.mass * s.vx), np.sum(s.mass * s.vy), np.sum(s.mass * s.vz)]) / M
        desc = {"workload": "cfg5: cluster gas unbinding, %d lattice gas cells against themselves + %d stars + %d DM "
                            "particles, targets split over the GPUs, all-reduce of potentials and corrections per pass"
                            % (len(g), len(s), len(d)), "n_gas": len(g), "n_star": len(s), "n_dm": len(d),
                "parallelism": "target groups round-robin over %d GPU(s), sources replicated" % world}
        jobs = [dict(kind="cluster-gas", offsets=single(len(g)), members=(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass),
                     groups=[(single(len(d)), d.mass, d.x, d.y, d.z), (single(len(s)), s.mass, s.x, s.y, s.z)],
                     split=world > 1, kw=dict(n_pre=0, split_classes=True, vb=vb, kappa=2.0))]
        return jobs, desc, "strong"
    raise SystemExit("unknown workload %r" % name)


_COMM = {}


def shared_comm(rank, world, device):
    """One NCCL communicator per process for all split-mode plans."""
    if "c" not in _COMM:
        import torch.distributed as dist
        from pyhalma_b200.unbind import Communicator, nccl_unique_id
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        _COMM["c"] = Communicator(uid[0], rank, world, device)
    return _COMM["c"]


SYMMETRIC = True      # set from --symmetric
REUSE = None          # set from --reuse: external-sum cache + incremental passes (None: the library default)
E2E_STREAMS = 3       # parts of a catalogue in the one-shot (e2e) call, --e2e-streams


def make_plan(job, mode, device, rank=0, world=1, upload=True, symmetric=None, reuse="default"):
    from pyhalma_b200.unbind import UnbindPlan
    kw = job["kw"]
    split = job["split"] and world > 1
    plan = UnbindPlan(job["offsets"], [g[0] for g in job["groups"]], mode=mode, n_pre=kw["n_pre"],
                      split_classes=kw["split_classes"], vb_fixed=kw["vb"] is not None, max_iter=64,
                      kappa=kw["kappa"], device=device, rank=rank if split else 0, n_ranks=world if split else 1,
                      symmetric=SYMMETRIC if symmetric is None else symmetric,
                      cache_external=REUSE if reuse == "default" else reuse,
                      incremental=REUSE if reuse == "default" else reuse)
    if split:
        plan.use_comm(shared_comm(rank, world, device))
    if upload:
        upload_job(plan, job)
    return plan


def reuse_requested() -> bool:
    from pyhalma_b200 import unbind as _unbind
    if REUSE is not None:
        return bool(REUSE)
    return _unbind._reuse_default("HALMA_CACHE_EXT") or _unbind._reuse_default("HALMA_INCREMENTAL")


def upload_job(plan, job):
    plan.upload_members(*job["members"])
    for k, g in enumerate(job["groups"]):
        plan.upload_group(k, *g[1:])
    if job["kw"]["vb"] is not None:
        plan.set_vb(job["kw"]["vb"])


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU port timing (oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_sample(jobs, n_targets: int):
    """Times the oracle's f32seq kernel on a contiguous slice of `n_targets` targets of every
    job (of its largest halo) against that halo's full source set (first pass of the loop).
    Cost is strictly ntest x ntotal, so the rate carries over to the whole workload."""
    from oracle import oracle as O
    f32 = np.float32
    pairs, secs = 0, 0.0
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for job in jobs:
        off = job["offsets"]
        h = int(np.argmax(np.diff(off)))
        a, b = int(off[h]), int(off[h + 1])
        x, y, z, _, _, _, m = [c[a:b] for c in job["members"]]
        ext = [tuple(arr[int(g[0][h]):int(g[0][h + 1])] for arr in g[1:]) for g in job["groups"]]
        src = [f32(np.concatenate([m] + [e[0] for e in ext])), f32(np.concatenate([x] + [e[1] for e in ext])),
               f32(np.concatenate([y] + [e[2] for e in ext])), f32(np.concatenate([z] + [e[3] for e in ext]))]
        nt = min(n_targets, len(x))
        tgt = [f32(c[:nt]) for c in (x, y, z)]
        t0 = time.perf_counter()
        O.brute_force_binding_energy(threads, len(src[0]), *src, nt, *tgt)
        secs += time.perf_counter() - t0
        pairs += nt * len(src[0])
    return pairs, secs, threads


def run_reference(args, rank, world):
    """--impl reference: the CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    os.environ.setdefault("OMP_WAIT_POLICY", "active")       # run.sh:3
    jobs, desc, scaling = make_workload(args.workload, 0, 1)
    n_t = args.ref_targets
    for _ in range(args.warmup):
        cpu_sample(jobs, max(256, n_t // 8))
    pairs, secs = 0, 0.0
    for _ in range(args.steps):
        p, s, threads = cpu_sample(jobs, n_t)
        pairs += p
        secs += s
    val = pairs / secs / 1e9
    sample = "per step: first %d targets of each job's largest halo x all its sources, float32 in-order sum" % n_t
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": desc,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C/OpenMP port of particle_subroutines.f90:466-514 (oracle/); the Fortran reference cannot "
                    "be compiled in this image (no Fortran compiler)"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def pinned_copy(arr):
    """float64 numpy array backed by pinned host memory."""
    from pyhalma_b200 import _lib
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    p = C.c_void_p()
    _lib.check(_lib.lib().halma_host_alloc(C.byref(p), max(arr.nbytes, 8)))
    buf = (C.c_double * max(arr.size, 1)).from_address(p.value)
    out = np.frombuffer(buf, dtype=np.float64, count=arr.size)
    out[:] = arr
    return out


def e2e_step(jobs_pinned, mode, device, rank=0, world=1):
    """One step through the public plan API from host buffers: create, H2D, run, D2H, destroy."""
    pairs = h2d = d2h = 0
    for job in jobs_pinned:
        if job["kind"] == "catalogue" and E2E_STREAMS > 1:
            # the one-shot catalogue call with its parts overlapped on separate streams / host threads
            from pyhalma_b200.unbind import unbind_catalogue
            kw = job["kw"]
            st = unbind_catalogue(job["offsets"], *job["members"], groups=job["groups"], n_pre=kw["n_pre"],
                                  split_classes=kw["split_classes"], vb=kw["vb"], kappa=kw["kappa"], max_iter=64,
                                  mode=mode, device=device, symmetric=SYMMETRIC, streams=E2E_STREAMS,
                                  cache_external=REUSE, incremental=REUSE).stats
        else:
            plan = make_plan(job, mode, device, rank, world, upload=True)
            try:
                st = plan.run()
                plan.download()
            finally:
                plan.close()
        n = len(job["members"][0])
        pairs += st.pairs // world if (job["split"] and world > 1) else st.pairs
        h2d += 7 * 8 * n + sum(4 * 8 * len(g[1]) for g in job["groups"])
        d2h += n * (1 + 4 + 8 + 4) + 80 * (len(job["offsets"]) - 1)
    return pairs, h2d, d2h


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from pyhalma_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.require_device(local_rank)
    mode = args.mode
    jobs, desc, scaling = make_workload(args.workload, rank, world)
    plans = [make_plan(j, mode, local_rank, rank, world) for j in jobs]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.fill_(1)                 # evict L2 between steps (inputs are far smaller than L2)
        torch.cuda.synchronize()
        out = []
        for p in plans:
            out.append(p.run())
        return out

    for _ in range(args.warmup):
        step()
    mb = _lib.microbench(local_rank) if rank == 0 else None
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    dev_ms = pot_ms = 0.0
    pairs = evals = launches = pot_launches = passes = 0
    busy_windows = []
    for _ in range(args.steps):
        tb = time.perf_counter()
        for st, job in zip(step(), jobs):
            dev_ms += st.total_ms
            pot_ms += st.potential_ms
            # in split mode every rank's counter covers the whole halo: count it once
            pairs += st.pairs // world if (job["split"] and world > 1) else st.pairs
            evals += st.evaluations // world if (job["split"] and world > 1) else st.evaluations
            launches += st.launches
            pot_launches += st.potential_launches
            passes += st.passes
        busy_windows.append((tb, time.perf_counter()))
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    wall_ms = (t1 - t0) * 1e3

    # e2e: public one-shot API from pinned host buffers
    jobs_pinned = []
    for j in jobs:
        jobs_pinned.append(dict(j, members=tuple(pinned_copy(a) for a in j["members"]),
                                groups=[(g[0],) + tuple(pinned_copy(a) for a in g[1:]) for g in j["groups"]]))
    e2e_step(jobs_pinned, mode, local_rank, rank, world)
    barrier()
    te0 = time.perf_counter()
    e2e_pairs = 0
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        p_, h2d, d2h = e2e_step(jobs_pinned, mode, local_rank, rank, world)
        e2e_pairs += p_
    barrier()
    e2e_s = time.perf_counter() - te0

    # the same workload with one-sided sums only (symmetric = 0): evaluations == interactions
    one_sided = None
    if SYMMETRIC or reuse_requested():
        for p in plans:
            p.close()
        plans = [make_plan(j, mode, local_rank, rank, world, symmetric=False, reuse=False) for j in jobs]
        step()
        barrier()
        o_ms = o_pot = 0.0
        o_pairs = 0
        o_steps = max(1, min(args.steps, 2))
        for _ in range(o_steps):
            for st, job in zip(step(), jobs):
                o_ms += st.total_ms
                o_pot += st.potential_ms
                o_pairs += st.pairs // world if (job["split"] and world > 1) else st.pairs
        barrier()
        one_sided = (o_ms, o_pot, o_pairs, o_steps)

    # reduce over ranks: max time, sum of work
    t = torch.tensor([dev_ms, wall_ms, e2e_s, pot_ms], dtype=torch.float64, device="cuda")
    w = torch.tensor([pairs, e2e_pairs, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, e2e_s_max, pot_ms_max = t.tolist()
    pairs_all, e2e_pairs_all, launches_all = w.tolist()

    reuse_on = reuse_requested() and mode == "fast"
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        value = pairs_all / (dev_ms_max * 1e-3) / 1e9
        # roofline of the dominant kernel (rank 0's launches)
        # rank 0's share of the work (already divided in split mode) over rank 0's kernel time
        pot_rate = evals / (pot_ms * 1e-3) / 1e9
        sm_mhz = clocks.get("sm_mhz") or mb["sm_clock_mhz"]
        peak = mb["rsq_per_clk_sm"] * mb["sm_count"] * sm_mhz * 1e6 / 1e9
        n_src = [len(j["members"][0]) + sum(len(g[1]) for g in j["groups"]) for j in jobs]
        n_tgt = [len(j["members"][0]) for j in jobs]
        # predicate-free path: sources are read once by the main tickets and once per axis-sorted copy
        alg_bytes_first_pass = sum(16 * s * 4 + 20 * t_ + 24 * t_ for s, t_ in zip(n_src, n_tgt))
        roofline = {
            "bound": "mufu", "kernel": "k_potential_fast" if mode == "fast" else "k_potential_exact",
            "achieved": pot_rate, "peak": peak, "unit": "G 1/r evaluations/s (one MUFU.RSQ each)", "frac": pot_rate / peak,
            "interactions_per_evaluation": pairs / max(evals, 1),
            "achieved_interactions": pairs / (pot_ms * 1e-3) / 1e9,
            "peak_how": "measured: MUFU.RSQ/clk/SM from halma_microbench (%.2f) x %d SMs x median SM clock sampled "
                        "during the timed region (%.0f MHz)" % (mb["rsq_per_clk_sm"], mb["sm_count"], sm_mhz),
            "nominal_peak": 16 * mb["sm_count"] * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e9,
            "avg_launch_ms": pot_ms / max(pot_launches, 1), "launches": pot_launches,
            "share_of_step": pot_ms / dev_ms,
            "hbm": {"algorithmic_bytes_first_pass": alg_bytes_first_pass,
                    "achieved_GBs": alg_bytes_first_pass * passes / max(len(jobs), 1) / (pot_ms * 1e-3) / 1e9
                    if pot_ms else None,
                    "peak_GBs": peaks.get("hbm_gbs"), "note": "far below HBM peak: the kernel is issue/MUFU bound"},
            "traffic": None,
            "microbench": mb,
        }
        try:
            # DRAM bytes per launch of the same kernel at this workload's first pass, from the
            # committed `ncu --set full` capture (profiles/ncu_traffic_r01.json)
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")))
            if tr.get("workload") == args.workload:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_note"] = tr["note"]
        except Exception:
            pass
        if world == 1:
            cp, cs, threads = cpu_sample(jobs, args.cpu_targets)
            cpu = {"value": cp / cs / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "first pass, first %d targets of each job x all sources (%.1f s of CPU work); "
                             "oracle/ f32seq, OpenMP" % (args.cpu_targets, cs)}
        else:
            cpu = None          # reported at N = 1 only
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(desc, mode=mode, l2="flushed between steps (256 MiB write)",
                           passes_per_step=passes / args.steps, symmetric_self_term=bool(SYMMETRIC),
                           reuse=reuse_on),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "interactions_per_step": pairs_all / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_pairs_all / e2e_s_max / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "how": "public plan API per job (catalogue jobs: unbind_catalogue with %d overlapped parts), "
                           "everything inside the timing: plan create + H2D from pinned host + device loop + D2H of "
                           "mask, potentials, energies, member lists + destroy" % E2E_STREAMS},
            "gpu_launches": int(launches_all),
            "one_sided": None if one_sided is None else {
                "value": one_sided[2] / (one_sided[0] * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": one_sided[0] / one_sided[3],
                "kernel": one_sided[2] / (one_sided[1] * 1e-3) / 1e9, "roofline_frac": one_sided[2] / (one_sided[1] * 1e-3) / 1e9 / peak,
                "scope": "rank 0" if world > 1 else "whole job", "steps": one_sided[3],
                "note": "same workload with halma_unbind_config.symmetric = cache_external = incremental = 0: every "
                        "interaction the reference's loop visits is evaluated, in every pass"},
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    for p in plans:
        p.close()
    if "c" in _COMM:
        _COMM.pop("c").close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--e2e-streams", type=int, default=3,
                    help="catalogue workloads: parts of the one-shot call that overlap upload, sort and download")
    ap.add_argument("--symmetric", type=int, default=1, choices=[0, 1],
                    help="evaluate member x member pairs once for both particles (FAST mode, not in split mode)")
    ap.add_argument("--reuse", type=int, default=None, choices=[0, 1],
                    help="do not repeat work between passes: external sums cached, incremental passes (FAST mode; "
                         "default: the library's, HALMA_CACHE_EXT / HALMA_INCREMENTAL)")
    ap.add_argument("--cpu-targets", type=int, default=90000,
                    help="targets per job in the cpu_baseline sample (cfg2: 1.3e11 pairs, ~14 s on 16 cores)")
    ap.add_argument("--ref-targets", type=int, default=20000, help="targets per job per step of --impl reference")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    global SYMMETRIC, E2E_STREAMS, REUSE
    REUSE = None if args.reuse is None else bool(args.reuse)
    SYMMETRIC = bool(args.symmetric) and args.mode == "fast"
    E2E_STREAMS = max(1, args.e2e_streams)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
