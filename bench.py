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
            (16 interactions / clk / SM nominal; the rate and the SM clock are measured)
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
def make_workload(name: str, rank: int):
    """Returns a list of jobs; a job = dict(kind, members, groups, kw) describing one plan."""
    from pyhalma_b200 import synth

    def star_gas_jobs(case):
        s, g, d = case.stars, case.gas, case.dm
        M = float(np.sum(s.mass))
        vb = np.array([np.sum(s.mass * s.vx), np.sum(s.mass * s.vy), np.sum(s.mass * s.vz)]) / M
        jobs = [dict(kind="stellar", members=(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass),
                     groups=[(g.mass, g.x, g.y, g.z)] + ([(d.mass, d.x, d.y, d.z)] if len(d) else []),
                     kw=dict(n_pre=1, split_classes=False, vb=None, kappa=case.factor_v ** 2)),
                dict(kind="gas", members=(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass),
                     groups=([(d.mass, d.x, d.y, d.z)] if len(d) else []) + [(s.mass, s.x, s.y, s.z)],
                     kw=dict(n_pre=0, split_classes=True, vb=vb, kappa=2.0))]
        return jobs

    if name == "cfg2":
        case = synth.config2(seed_extra=rank)
        desc = {"workload": "cfg2: NFW galaxy halo, 2e5 stars + 5e5 gas cells per GPU, iterative stellar + gas "
                            "unbinding", "n_star": len(case.stars), "n_gas": len(case.gas)}
        return star_gas_jobs(case), desc
    if name == "cfg1":
        case = synth.config1(seed_extra=rank)
        desc = {"workload": "cfg1: Plummer halo, 1e4 stars + 1e4 gas cells per GPU, iterative stellar + gas "
                            "unbinding", "n_star": len(case.stars), "n_gas": len(case.gas)}
        return star_gas_jobs(case), desc
    raise SystemExit("unknown workload %r" % name)


def job_offsets(job):
    n = len(job["members"][0])
    return np.array([0, n], np.int64), [np.array([0, len(g[0])], np.int64) for g in job["groups"]]


def make_plan(job, mode, device):
    from pyhalma_b200.unbind import UnbindPlan
    off, eoff = job_offsets(job)
    kw = job["kw"]
    plan = UnbindPlan(off, eoff, mode=mode, n_pre=kw["n_pre"], split_classes=kw["split_classes"],
                      vb_fixed=kw["vb"] is not None, max_iter=64, kappa=kw["kappa"], device=device)
    plan.upload_members(*job["members"])
    for k, g in enumerate(job["groups"]):
        plan.upload_group(k, *g)
    if kw["vb"] is not None:
        plan.set_vb(kw["vb"])
    return plan


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
    job against that job's full source set (first pass of the loop).  Cost is strictly
    ntest x ntotal, so the rate carries over to the whole workload."""
    from oracle import oracle as O
    f32 = np.float32
    pairs, secs = 0, 0.0
    threads = O.max_threads()
    for job in jobs:
        x, y, z, _, _, _, m = job["members"]
        src = [np.concatenate([m] + [g[0] for g in job["groups"]]),
               np.concatenate([x] + [g[1] for g in job["groups"]]),
               np.concatenate([y] + [g[2] for g in job["groups"]]),
               np.concatenate([z] + [g[3] for g in job["groups"]])]
        src = [f32(a) for a in src]
        nt = min(n_targets, len(x))
        tgt = [f32(a[:nt]) for a in (x, y, z)]
        best = None
        for _ in range(1):
            t0 = time.perf_counter()
            O.brute_force_binding_energy(threads, len(src[0]), *src, nt, *tgt)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        pairs += nt * len(src[0])
        secs += best
    return pairs, secs, threads


def run_reference(args, rank, world):
    """--impl reference: the CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    os.environ.setdefault("OMP_WAIT_POLICY", "active")       # run.sh:3
    jobs, desc = make_workload(args.workload, 0)
    n_t = args.ref_targets
    for _ in range(args.warmup):
        cpu_sample(jobs, max(256, n_t // 8))
    pairs, secs = 0, 0.0
    for _ in range(args.steps):
        p, s, threads = cpu_sample(jobs, n_t)
        pairs += p
        secs += s
    val = pairs / secs / 1e9
    sample = "per step: first %d targets of each job (stars, gas) x all %s sources, float32 in-order sum" % (
        n_t, "+".join(str(len(j["members"][0]) + sum(len(g[0]) for g in j["groups"])) for j in jobs))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
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


def e2e_step(jobs_pinned, mode, device):
    """One step through the public one-shot API: create, H2D, run, D2H, destroy."""
    from pyhalma_b200.unbind import unbind_catalogue
    pairs = h2d = d2h = 0
    for job in jobs_pinned:
        off, eoff = job_offsets(job)
        groups = [(eo,) + tuple(g) for eo, g in zip(eoff, job["groups"])]
        kw = job["kw"]
        res = unbind_catalogue(off, *job["members"], groups=groups, n_pre=kw["n_pre"],
                               split_classes=kw["split_classes"],
                               vb=None if kw["vb"] is None else np.asarray(kw["vb"]).reshape(1, 3),
                               kappa=kw["kappa"], mode=mode, device=device)
        n = len(job["members"][0])
        pairs += res.stats.pairs
        h2d += 7 * 8 * n + sum(4 * 8 * len(g[0]) for g in job["groups"])
        d2h += n * (1 + 4 + 8 + 4) + 80
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
    jobs, desc = make_workload(args.workload, rank)
    plans = [make_plan(j, mode, local_rank) for j in jobs]
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
    pairs = launches = pot_launches = passes = 0
    busy_windows = []
    for _ in range(args.steps):
        tb = time.perf_counter()
        for st in step():
            dev_ms += st.total_ms
            pot_ms += st.potential_ms
            pairs += st.pairs
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
        jobs_pinned.append(dict(kind=j["kind"], members=tuple(pinned_copy(a) for a in j["members"]),
                                groups=[tuple(pinned_copy(a) for a in g) for g in j["groups"]], kw=j["kw"]))
    e2e_step(jobs_pinned, mode, local_rank)
    barrier()
    te0 = time.perf_counter()
    e2e_pairs = 0
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        p_, h2d, d2h = e2e_step(jobs_pinned, mode, local_rank)
        e2e_pairs += p_
    barrier()
    e2e_s = time.perf_counter() - te0

    # reduce over ranks: max time, sum of work
    t = torch.tensor([dev_ms, wall_ms, e2e_s, pot_ms], dtype=torch.float64, device="cuda")
    w = torch.tensor([pairs, e2e_pairs, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, e2e_s_max, pot_ms_max = t.tolist()
    pairs_all, e2e_pairs_all, launches_all = w.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        value = pairs_all / (dev_ms_max * 1e-3) / 1e9
        # roofline of the dominant kernel (rank 0's launches)
        pot_rate = pairs / (pot_ms * 1e-3) / 1e9
        sm_mhz = clocks.get("sm_mhz") or mb["sm_clock_mhz"]
        peak = mb["rsq_per_clk_sm"] * mb["sm_count"] * sm_mhz * 1e6 / 1e9
        n_src = [len(j["members"][0]) + sum(len(g[0]) for g in j["groups"]) for j in jobs]
        n_tgt = [len(j["members"][0]) for j in jobs]
        alg_bytes_first_pass = sum(16 * s + 12 * t_ + 8 * t_ for s, t_ in zip(n_src, n_tgt))
        roofline = {
            "bound": "mufu", "kernel": "k_potential_fast" if mode == "fast" else "k_potential_exact",
            "achieved": pot_rate, "peak": peak, "unit": UNIT, "frac": pot_rate / peak,
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
        cp, cs, threads = cpu_sample(jobs, args.cpu_targets)
        cpu = {"value": cp / cs / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first pass, first %d targets of each job x all sources (%.1f s of CPU work); "
                         "oracle/ f32seq, OpenMP" % (args.cpu_targets, cs)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(desc, mode=mode, l2="flushed between steps (256 MiB write)",
                           passes_per_step=passes / args.steps, parallelism="1 halo per GPU, no collective"),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "interactions_per_step": pairs_all / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_pairs_all / e2e_s_max / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "how": "pyhalma_b200.unbind_catalogue one-shot per job: plan create + H2D from pinned host + "
                           "device loop + D2H of mask, potentials, energies, member lists"},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    for p in plans:
        p.close()
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
    ap.add_argument("--cpu-targets", type=int, default=12000, help="targets per job in the cpu_baseline sample")
    ap.add_argument("--ref-targets", type=int, default=2500, help="targets per job per step of --impl reference")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
