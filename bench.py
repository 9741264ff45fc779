#!/usr/bin/env python
"""Benchmark of the unbinding hot path (BASELINE.json metric: unbinding Ginteractions/s and halo-catalogue
wall time at 1/2/4/8 B200 vs host OpenMP).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg3]

Default workload at EVERY N = BASELINE.json configs[2] ("cfg3"): the catalogue of 1e4 haloes, N = 1e2..1e5,
cost-sharded (LPT on N^2) over the N GPUs with no data-path collective -- strong scaling of a fixed job, the
configuration the metric's "halo-catalogue wall time at 1/2/4/8" is quoted on.  A step is `reps` complete,
back-to-back unbindings of the catalogue (each from the pristine resident inputs, L2 flushed in between), so
that the timed region spans seconds; `catalogue_wall_ms` = ms_per_step / reps.

One interaction = one (target, source) pair visited in one potential pass; excluded pairs count
(BASELINE.md §3).

  value     whole-job Ginteractions/s, inputs resident in HBM, CUDA-event time of the runs, max over ranks
  e2e       the same through the public one-shot API with pinned HOST buffers: plan creation, H2D of every
            input, the device loop and D2H of mask + potentials + energies + member lists inside the timing
  roofline  the potential evaluation alone against the MUFU.RSQ issue roofline (16 rsqrt / clk / SM nominal;
            the rate and the SM clock are measured), in 1/r EVALUATIONS per second (the symmetric self-term
            and the reuse between passes make evaluations < interactions)
  cpu_baseline  the C/OpenMP oracle (a port of the Fortran kernel, oracle/) on the host cores (N = 1 only)
  parity    the cached full-size parity report (profiles/parity_r02.json, scripts/parity_report.py)
Sub-objects (default workload only): `cfg2` (N = 1: configs[1], the round-1 headline), `single_halo_1e6`
(N = 1: one potential pass over one 1e6-star halo, one-sided and symmetric, north_star's >= 0.60 target),
`split_cfg4` (N > 1: ONE 2e6-star halo with its targets split over the GPUs, NCCL all-reduce per pass,
compared bit for bit with the single-GPU run), `f2py_level` (N = 1: RPS's sequence of f2py-level calls).

--impl reference times the CPU port alone, with all host threads, on a bounded sample of the same workload
(the reference's Fortran cannot be compiled in this image).
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
def single(n):
    return np.array([0, n], np.int64)


def star_gas_jobs(case):
    s, g, d = case.stars, case.gas, case.dm
    M = float(np.sum(s.mass))
    vb = np.array([np.sum(s.mass * s.vx), np.sum(s.mass * s.vy), np.sum(s.mass * s.vz)]) / M
    grp = lambda p: (single(len(p)), p.mass, p.x, p.y, p.z)  # noqa: E731
    return [dict(kind="stellar", offsets=single(len(s)), members=(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass),
                 groups=[grp(g)] + ([grp(d)] if len(d) else []), split=False,
                 kw=dict(n_pre=1, split_classes=False, vb=None, kappa=case.factor_v ** 2)),
            dict(kind="gas", offsets=single(len(g)), members=(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass),
                 groups=([grp(d)] if len(d) else []) + [grp(s)], split=False,
                 kw=dict(n_pre=0, split_classes=True, vb=vb, kappa=2.0))]


def make_workload(name: str, rank: int, world: int = 1):
    """Returns (jobs, desc, scaling).  A job describes one plan: CSR offsets, member arrays,
    external groups (ext_offsets, mass, x, y, z), layout keywords, and whether the halo is
    shared by all ranks (split mode)."""
    from pyhalma_b200 import sharding, synth

    if name in ("cfg2", "cfg1"):
        case = synth.config2(seed_extra=rank) if name == "cfg2" else synth.config1(seed_extra=rank)
        what = ("cfg2: NFW galaxy halo, 2e5 stars + 5e5 gas cells per GPU" if name == "cfg2"
                else "cfg1: Plummer halo, 1e4 stars + 1e4 gas cells per GPU")
        desc = {"workload": what + ", iterative stellar + gas unbinding", "n_star": len(case.stars),
                "n_gas": len(case.gas), "parallelism": "1 halo per GPU, no collective"}
        return star_gas_jobs(case), desc, "weak"
    if name == "cfg3":
        cat = synth.config3()                       # the same catalogue on every rank
        costs = sharding.halo_costs(cat.offsets)
        parts = sharding.lpt_partition(costs, world)
        off, cols = sharding.take_haloes(cat.offsets, [cat.x, cat.y, cat.z, cat.vx, cat.vy, cat.vz, cat.mass],
                                         parts[rank])
        desc = {"workload": "cfg3: catalogue of 1e4 Plummer haloes, N = 1e2..1e5 (dN/dN ~ N^-1.9), batched launch, "
                            "iterative stellar unbinding, haloes cost-sharded over the GPUs", "n_halo": cat.n_halo,
                "sum_n": int(cat.offsets[-1]), "sum_n2": cat.meta["sum_n2"], "halos_this_rank": int(len(parts[rank])),
                "lpt_imbalance": sharding.partition_imbalance(costs, parts),
                "parallelism": "haloes LPT-sharded by N^2 over %d GPU(s), no collective" % world}
        jobs = [dict(kind="catalogue", offsets=off, members=tuple(cols), groups=[], split=False,
                     kw=dict(n_pre=0, split_classes=False, vb=None, kappa=9.0))]
        return jobs, desc, "strong"
    if name == "cfg4":
        n = int(os.environ.get("HALMA_CFG4_N", "2000000"))
        p = synth.config4(n)                        # the same halo on every rank
        desc = {"workload": "cfg4: one cluster-scale stellar halo, N = %d, targets split over the GPUs, NCCL "
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
E2E_STREAMS = 0       # parts of a catalogue in the one-shot (e2e) call, --e2e-streams (0: the library's choice)
DRIVER = None         # set from --driver (None: the library default = the persistent loop kernel on one GPU)


def make_plan(job, mode, device, rank=0, world=1, upload=True, symmetric=None, reuse="default", max_iter=64):
    from pyhalma_b200.unbind import UnbindPlan
    kw = job["kw"]
    split = job["split"] and world > 1
    plan = UnbindPlan(job["offsets"], [g[0] for g in job["groups"]], mode=mode, n_pre=kw["n_pre"],
                      split_classes=kw["split_classes"], vb_fixed=kw["vb"] is not None, max_iter=max_iter,
                      kappa=kw["kappa"], device=device, rank=rank if split else 0, n_ranks=world if split else 1,
                      symmetric=SYMMETRIC if symmetric is None else symmetric,
                      cache_external=REUSE if reuse == "default" else reuse,
                      incremental=REUSE if reuse == "default" else reuse, driver=DRIVER)
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
# clocks: NVML sampled every 10 ms in a thread (nvidia-smi -lms as the fallback)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"))
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period: float = 0.01):
        self.index, self.period = index, period
        self.rows = []              # (t, sm_mhz, reasons bitmask)
        self.sm_max = None
        self.how = None
        self._stop = threading.Event()
        self._thread = None
        self.proc = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def pump():
                while not self._stop.is_set():
                    try:
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                          int(reasons(h))))
                    except Exception:
                        pass
                    time.sleep(self.period)

            self._thread = threading.Thread(target=pump, daemon=True)
            self._thread.start()
            self.how = "nvml every %d ms" % round(self.period * 1e3)
            return
        except Exception:
            self._thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump_smi, daemon=True).start()
            self.how = "nvidia-smi -lms 20"
        except Exception:
            self.proc = None

    def _pump_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                mask = 0
                for (bit, _), v in zip(self.REASONS, f[3:7]):
                    if v.lower().startswith("active"):
                        mask |= bit
                self.sm_max = float(f[1])
                self.rows.append((time.perf_counter(), float(f[0]), mask))
            except Exception:
                continue

    def stop(self, windows) -> dict:
        """windows: list of (t0, t1) perf_counter intervals during which the GPU was under load."""
        if self._thread is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock source"]}
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        self._stop.set()
        sm, mask = [], 0
        for t, mhz, m in self.rows:
            if any(a <= t <= b for a, b in windows):
                sm.append(mhz)
                mask |= m
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                "reasons": sorted(n for bit, n in self.REASONS if mask & bit), "how": self.how}


# ----------------------------------------------------------------------------------------------
# CPU port timing (oracle/)
# ----------------------------------------------------------------------------------------------
def host_threads() -> int:
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_sample(jobs, budget_pairs: float):
    """Times the oracle's f32seq kernel (first pass of the loop: every member of a halo against all its
    sources) on a bounded sample of the workload: per job, haloes in descending size at a stride chosen so
    that the sample holds about budget_pairs / len(jobs) pairs (the largest halo is always in; a halo that
    alone exceeds the share contributes a contiguous slice of its targets).  Cost is strictly
    ntest x ntotal, so the rate carries over to the whole workload.  Returns (pairs, seconds, threads, n_calls)."""
    from oracle import oracle as O
    f32 = np.float32
    pairs, secs, calls = 0, 0.0, 0
    threads = host_threads()
    share = budget_pairs / max(len(jobs), 1)
    for job in jobs:
        off = job["offsets"]
        n = np.diff(off).astype(np.float64)
        next_ = sum((np.diff(g[0]).astype(np.float64) for g in job["groups"]), np.zeros(len(n)))
        cost = n * (n + next_)
        order = np.argsort(-cost, kind="stable")
        stride = max(1, int(round(cost.sum() / share)))
        left = share
        for h in order[::stride]:
            if left <= 0 or cost[h] <= 0:
                break
            a, b = int(off[h]), int(off[h + 1])
            x, y, z, _, _, _, m = [c[a:b] for c in job["members"]]
            ext = [tuple(arr[int(g[0][h]):int(g[0][h + 1])] for arr in g[1:]) for g in job["groups"]]
            src = [f32(np.concatenate([m] + [e[0] for e in ext])), f32(np.concatenate([x] + [e[1] for e in ext])),
                   f32(np.concatenate([y] + [e[2] for e in ext])), f32(np.concatenate([z] + [e[3] for e in ext]))]
            nt = int(min(len(x), max(1, left // len(src[0]))))
            tgt = [f32(c[:nt]) for c in (x, y, z)]
            t0 = time.perf_counter()
            O.brute_force_binding_energy(threads, len(src[0]), *src, nt, *tgt)
            secs += time.perf_counter() - t0
            pairs += nt * len(src[0])
            left -= nt * len(src[0])
            calls += 1
    return pairs, secs, threads, calls


def run_reference(args, rank, world):
    """--impl reference: the CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    os.environ.setdefault("OMP_WAIT_POLICY", "active")       # run.sh:3
    # the whole job: at N > 1 the GPU arm shards this same catalogue, the CPU arm has one host
    jobs, desc, scaling = make_workload(args.workload, 0, 1)
    desc["parallelism"] = "host OpenMP, %d threads" % host_threads()
    budget = args.ref_pairs
    for _ in range(args.warmup):
        cpu_sample(jobs, budget / 8)
    pairs, secs, calls = 0, 0.0, 0
    for _ in range(args.steps):
        p, s, threads, c = cpu_sample(jobs, budget)
        pairs += p
        secs += s
        calls += c
    val = pairs / secs / 1e9
    sample = ("per step: first pass (all sources of the halo) over a strided sample of the haloes in descending size, "
              "about %.2g pairs in %d oracle calls, float32 in-order sum" % (pairs / args.steps, calls // max(args.steps, 1)))
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


def pin_jobs(jobs):
    return [dict(j, members=tuple(pinned_copy(a) for a in j["members"]),
                 groups=[(g[0],) + tuple(pinned_copy(a) for a in g[1:]) for g in j["groups"]]) for j in jobs]


def e2e_step(jobs_pinned, mode, device, rank=0, world=1):
    """One step through the public plan API from host buffers: create, H2D, run, D2H, destroy."""
    pairs = h2d = d2h = 0
    for job in jobs_pinned:
        if job["kind"] == "catalogue" and E2E_STREAMS != 1:
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


def e2e_parts_note(jobs):
    if E2E_STREAMS:
        return str(E2E_STREAMS)
    from pyhalma_b200.unbind import auto_parts
    return "/".join(str(auto_parts(len(j["members"][0]), sum(len(g[1]) for g in j["groups"]))) for j in jobs
                    if j["kind"] == "catalogue") + " (auto)"


class Timed:
    """Device-resident measurement of a list of plans: sums of the runs' CUDA-event times and counters."""

    def __init__(self):
        self.dev_ms = self.pot_ms = self.comm_ms = 0.0
        self.pairs = self.evals = self.launches = self.pot_launches = self.passes = self.comm_bytes = self.runs = 0
        self.windows = []
        self.phase_ms = [0.0] * 5

    def add(self, st, job, world):
        shared = job["split"] and world > 1      # in split mode every rank's counter covers the whole halo
        self.dev_ms += st.total_ms
        self.pot_ms += st.potential_ms
        self.comm_ms += st.comm_ms
        self.comm_bytes += st.comm_bytes
        self.pairs += st.pairs // world if shared else st.pairs
        self.evals += st.evaluations // world if shared else st.evaluations
        self.launches += st.launches
        self.pot_launches += st.potential_launches
        self.passes += st.passes
        self.runs += 1
        self.phase_ms = [a + b for a, b in zip(self.phase_ms, st.phase_ms)] if st.phase_ms else self.phase_ms
        DRIVER_RAN["d"] = st.driver


def run_steps(plans, jobs, world, steps, reps, flush_fn, out: Timed):
    for _ in range(steps):
        for _ in range(reps):
            flush_fn()
            tb = time.perf_counter()
            for p, job in zip(plans, jobs):
                out.add(p.run(), job, world)
            out.windows.append((tb, time.perf_counter()))


def roofline_entry(t: Timed, mb, sm_mhz, mode, kernel):
    peak = mb["rsq_per_clk_sm"] * mb["sm_count"] * sm_mhz * 1e6 / 1e9
    rate = t.evals / (t.pot_ms * 1e-3) / 1e9 if t.pot_ms else 0.0
    return peak, {"bound": "mufu", "kernel": kernel, "achieved": rate, "peak": peak,
                  "unit": "G 1/r evaluations/s (one MUFU.RSQ each)", "frac": rate / peak,
                  "interactions_per_evaluation": t.pairs / max(t.evals, 1),
                  "achieved_interactions": t.pairs / (t.pot_ms * 1e-3) / 1e9 if t.pot_ms else 0.0,
                  "whole_run_frac": t.evals / (t.dev_ms * 1e-3) / 1e9 / peak,
                  "peak_how": "measured: MUFU.RSQ/clk/SM from halma_microbench (%.2f) x %d SMs x median SM clock sampled "
                              "during the timed region (%.0f MHz)" % (mb["rsq_per_clk_sm"], mb["sm_count"], sm_mhz),
                  "avg_pass_ms": t.pot_ms / max(t.pot_launches, 1), "potential_passes": t.pot_launches,
                  "share_of_step": t.pot_ms / t.dev_ms if t.dev_ms else None,
                  "phase_ms_per_run": dict(zip(("prologue", "potential", "energy_compaction", "tables", "epilogue"),
                                               [v / max(t.runs, 1) for v in t.phase_ms]))}


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from pyhalma_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.require_device(local_rank)
    mode = args.mode
    default_line = args.workload == "cfg3" and not args.no_sub
    jobs, desc, scaling = make_workload(args.workload, rank, world)
    plans = [make_plan(j, mode, local_rank, rank, world) for j in jobs]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        flush.fill_(1)                 # evict L2 between runs (256 MiB write)
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up: W complete steps of one repetition; sizes `reps` so that the timed region spans >= --min-seconds
    wt = Timed()
    run_steps(plans, jobs, world, args.warmup, 1, flush_l2, wt)
    run_ms = allmax(wt.dev_ms / max(wt.runs // max(len(jobs), 1), 1))
    reps = args.reps
    if reps <= 0:
        reps = 1 if args.workload != "cfg3" else 16
        need = int(np.ceil(args.min_seconds * 1e3 / max(args.steps * run_ms, 1e-6)))
        reps = min(max(reps, need), 256) if args.workload == "cfg3" else reps
    mb = _lib.microbench(local_rank) if rank == 0 else None
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.05)
    barrier()
    t0 = time.perf_counter()
    tm = Timed()
    run_steps(plans, jobs, world, args.steps, reps, flush_l2, tm)
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(tm.windows)
    wall_ms = (t1 - t0) * 1e3

    # e2e: public one-shot API from pinned host buffers
    jobs_pinned = pin_jobs(jobs)
    e2e_step(jobs_pinned, mode, local_rank, rank, world)
    barrier()
    te0 = time.perf_counter()
    e2e_pairs = 0
    e2e_steps = max(1, min(args.steps * reps, args.e2e_steps))
    for _ in range(e2e_steps):
        p_, h2d, d2h = e2e_step(jobs_pinned, mode, local_rank, rank, world)
        e2e_pairs += p_
    barrier()
    e2e_s = time.perf_counter() - te0
    del jobs_pinned

    # the same workload with one-sided sums only and nothing reused between passes: evaluations == interactions
    one_sided = None
    if (SYMMETRIC or reuse_requested()) and not args.no_one_sided:
        for p in plans:
            p.close()
        plans = [make_plan(j, mode, local_rank, rank, world, symmetric=False, reuse=False) for j in jobs]
        ot = Timed()
        run_steps(plans, jobs, world, 1, 1, flush_l2, Timed())
        barrier()
        run_steps(plans, jobs, world, max(1, min(args.steps, 2)), 1, flush_l2, ot)
        barrier()
        one_sided = ot
    for p in plans:
        p.close()
    plans = []

    # reduce over ranks: max time, sum of work; clocks: the slowest GPU, every reason seen
    t = torch.tensor([tm.dev_ms, wall_ms, e2e_s, -(clocks.get("sm_mhz") or 1e9)], dtype=torch.float64, device="cuda")
    w = torch.tensor([tm.pairs, e2e_pairs, tm.launches, clocks.get("samples") or 0,
                      sum(1 << k for k, (_, n) in enumerate(ClockSampler.REASONS) if n in clocks.get("reasons", []))],
                     dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(w) for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(gathered, w)
    else:
        gathered = [w]
    dev_ms_max, wall_ms_max, e2e_s_max, neg_min_mhz = t.tolist()
    pairs_all = sum(float(g[0]) for g in gathered)
    e2e_pairs_all = sum(float(g[1]) for g in gathered)
    launches_all = sum(float(g[2]) for g in gathered)
    reason_mask = 0
    for g in gathered:
        reason_mask |= int(g[4])
    clocks_all = dict(clocks, sm_mhz=(-neg_min_mhz if neg_min_mhz > -1e8 else None),
                      samples=int(min(float(g[3]) for g in gathered)),
                      reasons=sorted(n for k, (_, n) in enumerate(ClockSampler.REASONS) if reason_mask & (1 << k)),
                      scope="slowest of %d GPUs (median under load), fewest samples, union of reasons" % world)

    # sub-objects of the default line
    subs = {}
    if default_line:
        if world == 1:
            subs["cfg2"] = sub_cfg2(mode, local_rank, flush_l2, mb, clocks)
            subs["single_halo_1e6"] = sub_single_halo(mode, local_rank, flush_l2, mb, clocks)
            subs["f2py_level"] = sub_f2py_level(local_rank)
        else:
            subs["cfg2"] = subs["single_halo_1e6"] = subs["f2py_level"] = None      # single-GPU figures: see the N = 1 line
            subs["split_cfg4"] = sub_split_cfg4(mode, rank, local_rank, world, flush_l2, barrier, allmax)

    reuse_on = reuse_requested() and mode == "fast"
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        value = pairs_all / (dev_ms_max * 1e-3) / 1e9
        sm_mhz = clocks.get("sm_mhz") or mb["sm_clock_mhz"]
        kernel = {_lib.DRIVER_FUSED: "k_unbind_loop (persistent: the potential phases of the one launch per run, "
                                     "timed by the kernel's own %globaltimer stamps)",
                  _lib.DRIVER_ENQUEUE: "k_potential_fast (stand-alone launches, CUDA events)"}.get(
                      DRIVER_RAN.get("d"), "k_potential")
        peak, roofline = roofline_entry(tm, mb, sm_mhz, mode, kernel if mode == "fast" else "k_potential_exact")
        n_src = [len(j["members"][0]) + sum(len(g[1]) for g in j["groups"]) for j in jobs]
        n_tgt = [len(j["members"][0]) for j in jobs]
        # predicate-free path: sources are read once by the main tickets and once per axis-sorted copy
        alg_bytes_first_pass = sum(16 * s * 4 + 20 * t_ + 24 * t_ for s, t_ in zip(n_src, n_tgt))
        roofline["hbm"] = {"algorithmic_bytes_first_pass": alg_bytes_first_pass,
                           "achieved_GBs": alg_bytes_first_pass * tm.passes / max(len(jobs), 1) / (tm.pot_ms * 1e-3) / 1e9
                           if tm.pot_ms else None,
                           "peak_GBs": peaks.get("hbm_gbs"), "note": "far below HBM peak: the kernel is issue/MUFU bound"}
        roofline["traffic"] = None
        roofline["microbench"] = mb
        try:
            # DRAM bytes per launch of the same kernel at this workload, from the committed `ncu --set full` capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")))
            if tr.get("workload") == args.workload:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_note"] = tr["note"]
        except Exception:
            pass
        if world == 1 and not args.no_cpu:
            cp, cs, threads, calls = cpu_sample(jobs, args.cpu_pairs)
            cpu = {"value": cp / cs / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "first pass over a strided sample of the haloes in descending size: %.3g pairs in %d "
                             "oracle calls (%.1f s of CPU work); oracle/ f32seq, OpenMP" % (cp, calls, cs)}
        else:
            cpu = None          # reported at N = 1 only
        parity = None
        try:
            parity = json.load(open(os.path.join(ROOT, "profiles", "parity_r02.json")))
        except Exception:
            pass
        runs = args.steps * reps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(desc, mode=mode, reps_per_step=reps,
                           step="%d complete unbinding(s) of the workload from the pristine resident inputs" % reps,
                           l2="flushed before every run (256 MiB write)",
                           passes_per_run=tm.passes / max(tm.runs, 1), symmetric_self_term=bool(SYMMETRIC),
                           reuse=reuse_on, loop_driver={0: "auto", 1: "graph", 2: "enqueue", 3: "fused"}.get(DRIVER_RAN.get("d"), "?")),
            "catalogue_wall_ms" if args.workload == "cfg3" else "run_ms": dev_ms_max / runs,
            "wall_ms_per_step": wall_ms_max / args.steps,
            "timed_region_s": wall_ms_max * 1e-3,
            "interactions_per_step": pairs_all / args.steps,
            "clocks": clocks_all,
            "e2e": {"value": e2e_pairs_all / e2e_s_max / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d * reps,
                    "d2h_bytes_per_step": d2h * reps, "runs": e2e_steps, "ms_per_run": e2e_s_max / e2e_steps * 1e3,
                    "vs_resident": (e2e_pairs_all / e2e_s_max / 1e9) / value,
                    "how": "public plan API per job (catalogue jobs: unbind_catalogue with %s overlapped parts), "
                           "everything inside the timing: plan create + H2D from pinned host + device loop + D2H of "
                           "mask, potentials, energies, member lists + destroy; bytes are per step = %d runs"
                           % (e2e_parts_note(jobs), reps)},
            "gpu_launches": int(launches_all),
            "gpu_launches_note": "kernel launches of libhalma_unbind.so inside the timed region, all ranks; with the "
                                 "persistent loop kernel one launch is one complete unbinding",
            "one_sided": None if one_sided is None else {
                "value": one_sided.pairs / (one_sided.dev_ms * 1e-3) / 1e9, "unit": UNIT,
                "ms_per_run": one_sided.dev_ms / max(one_sided.runs, 1),
                "kernel": one_sided.pairs / (one_sided.pot_ms * 1e-3) / 1e9,
                "roofline_frac": one_sided.pairs / (one_sided.pot_ms * 1e-3) / 1e9 / peak,
                "scope": "rank 0" if world > 1 else "whole job",
                "note": "same workload with halma_unbind_config.symmetric = cache_external = incremental = 0: every "
                        "interaction the reference's loop visits is evaluated, in every pass"},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "limiter": limiter_note(tm, world),
        }
        line.update(subs)
        print(json.dumps(line), flush=True)
    if "c" in _COMM:
        _COMM.pop("c").close()
    if world > 1:
        dist.destroy_process_group()


DRIVER_RAN = {}


def limiter_note(t: Timed, world: int) -> str:
    other = t.dev_ms - t.pot_ms
    return ("per run: %.2f ms potential phases (MUFU-bound; tail of the last tickets included) + %.2f ms everything else "
            "(pack, O(N) energy / compaction phases at HBM speed, grid barriers, ticket tables); no collective in the "
            "catalogue path, so what keeps N-GPU efficiency below 1 is that fixed part and the shorter potential "
            "phases filling the machine less well, not communication" % (t.pot_ms / max(t.runs, 1), other / max(t.runs, 1)))


def _timed_plans(jobs, mode, device, flush_l2, steps, warm=2, **plan_kw):
    plans = [make_plan(j, mode, device, **plan_kw) for j in jobs]
    try:
        run_steps(plans, jobs, 1, warm, 1, flush_l2, Timed())
        t = Timed()
        run_steps(plans, jobs, 1, steps, 1, flush_l2, t)
        return t
    finally:
        for p in plans:
            p.close()


def sub_cfg2(mode, device, flush_l2, mb, clocks):
    """BASELINE configs[1]: one NFW galaxy halo, 2e5 stars + 5e5 gas cells, stellar + gas unbinding (round-1 headline)."""
    jobs, desc, _ = make_workload("cfg2", 0, 1)
    t = _timed_plans(jobs, mode, device, flush_l2, 5)
    sm_mhz = clocks.get("sm_mhz") or mb["sm_clock_mhz"]
    peak, roof = roofline_entry(t, mb, sm_mhz, mode, "potential phases")
    jp = pin_jobs(jobs)
    e2e_step(jp, mode, device)
    te = time.perf_counter()
    ep = 0
    for _ in range(3):
        ep += e2e_step(jp, mode, device)[0]
    es = time.perf_counter() - te
    o = _timed_plans(jobs, mode, device, flush_l2, 2, warm=1, symmetric=False, reuse=False)
    return {"workload": desc["workload"], "value": t.pairs / (t.dev_ms * 1e-3) / 1e9, "unit": UNIT,
            "ms_per_step": t.dev_ms / 5, "passes_per_step": t.passes / 5, "gpu_launches_per_step": t.launches / 5,
            "e2e": ep / es / 1e9, "roofline_frac": roof["frac"], "interactions_per_evaluation": roof["interactions_per_evaluation"],
            "one_sided": {"value": o.pairs / (o.dev_ms * 1e-3) / 1e9, "roofline_frac": o.pairs / (o.pot_ms * 1e-3) / 1e9 / peak}}


def sub_single_halo(mode, device, flush_l2, mb, clocks):
    """north_star's target: the all-pairs potential of ONE 1e6-particle halo at >= 60 % of the MUFU roofline.
    One potential pass (max_iter = 1) over 1e6 stars, one-sided and with the symmetric self-term."""
    from pyhalma_b200 import synth
    n = 1_000_000
    p = synth.config4(n)
    job = dict(kind="giant", offsets=single(n), members=(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass), groups=[],
               split=False, kw=dict(n_pre=0, split_classes=False, vb=None, kappa=9.0))
    sm_mhz = clocks.get("sm_mhz") or mb["sm_clock_mhz"]
    out = {"workload": "one Plummer halo of 1e6 stars, ONE potential pass + energy step + mask (max_iter = 1)", "n": n}
    for name, sym in (("one_sided", False), ("symmetric", True)):
        t = _timed_plans([job], mode, device, flush_l2, 3, warm=2, symmetric=sym, reuse=False, max_iter=1)
        peak, roof = roofline_entry(t, mb, sm_mhz, mode, "potential phase")
        out[name] = {"ms_per_pass": t.dev_ms / 3, "potential_ms": t.pot_ms / 3,
                     "Ginteractions_per_s": t.pairs / (t.dev_ms * 1e-3) / 1e9,
                     "G_evaluations_per_s": roof["achieved"], "roofline_frac": roof["frac"],
                     "interactions_per_evaluation": roof["interactions_per_evaluation"]}
    out["target"] = "north_star: >= 0.60 of the FP32 FMA/MUFU roofline (%.0f G evaluations/s measured peak)" % peak
    return out


def sub_f2py_level(device):
    """The zero-edit drop-in: RPS's sequence of f2py-level calls (halo_gas.py:306-450) at cfg2 sizes through
    fortran_modules.particle.particle.brute_force_binding_energy (host float32 arrays in and out per call)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_f2py_call
        return bench_f2py_call.rps_sequence(device)
    except Exception as exc:      # never lose the main line to a sub-object
        return {"error": repr(exc)[:300]}


def sub_split_cfg4(mode, rank, device, world, flush_l2, barrier, allmax):
    """BASELINE configs[3]: ONE 2e6-star halo, targets split over the GPUs, sources replicated, NCCL all-reduce of
    the potentials (+ corrections and two-sided sums) per pass; compared bit for bit with the single-GPU run."""
    try:
        jobs, desc, _ = make_workload("cfg4", rank, world)
        plan = make_plan(jobs[0], mode, device, rank, world)
        plan.run()
        barrier()
        t = Timed()
        run_steps([plan], jobs, world, 2, 1, flush_l2, t)
        barrier()
        res = plan.download()
        plan.close()
        ms = allmax(t.dev_ms / 2)
        comm = allmax(t.comm_ms / 2)
        out = None
        if rank == 0:
            # the same halo on one GPU (rank 0 alone; the others wait at the barrier)
            solo = dict(jobs[0], split=False)
            p1 = make_plan(solo, mode, device, 0, 1)
            p1.run()
            s1 = Timed()
            run_steps([p1], [solo], 1, 1, 1, flush_l2, s1)
            r1 = p1.download()
            p1.close()
            same = bool(np.array_equal(res.mask, r1.mask) and np.array_equal(res.be32.view(np.uint32), r1.be32.view(np.uint32))
                        and np.array_equal(res.idx_packed, r1.idx_packed) and res.halos[0].n_iter == r1.halos[0].n_iter)
            passes = t.passes / 2
            out = {"workload": desc["workload"], "n_gpus": world, "ms_per_step": ms, "passes": passes,
                   "value": t.pairs * world / 2 / (ms * 1e-3) / 1e9, "unit": UNIT,
                   "one_gpu_ms_per_step": s1.dev_ms, "speedup_vs_1gpu": s1.dev_ms / ms,
                   "bit_identical_to_1gpu": same,
                   "allreduce_ms_per_pass": comm / max(passes, 1), "allreduce_share_of_step": comm / ms,
                   "allreduce_bytes_per_pass": t.comm_bytes / 2 / max(passes, 1),
                   "note": "collectives are timed with CUDA events on the plan's stream (they include waiting for the "
                           "slowest rank to arrive); kernels 2-3 run replicated, so one grouped all-reduce per pass is "
                           "the whole exchange"}
        barrier()
        return out
    except Exception as exc:
        return {"error": repr(exc)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--reps", type=int, default=0,
                    help="complete unbindings of the workload per step (0: 16 for cfg3, more if K steps would "
                         "take less than --min-seconds; 1 for the other workloads)")
    ap.add_argument("--min-seconds", type=float, default=2.0, help="shortest timed region (cfg3)")
    ap.add_argument("--e2e-steps", type=int, default=8, help="runs of the end-to-end measurement")
    ap.add_argument("--e2e-streams", type=int, default=0,
                    help="catalogue workloads: parts of the one-shot call that overlap upload, sort and download "
                         "(0: chosen by the library from the size of the upload, unbind.auto_parts)")
    ap.add_argument("--symmetric", type=int, default=1, choices=[0, 1],
                    help="evaluate member x member pairs once for both particles (FAST mode)")
    ap.add_argument("--reuse", type=int, default=None, choices=[0, 1],
                    help="do not repeat work between passes: external sums cached, incremental passes (FAST mode; "
                         "default: the library's, HALMA_CACHE_EXT / HALMA_INCREMENTAL)")
    ap.add_argument("--driver", default=None, choices=["auto", "fused", "enqueue", "graph"],
                    help="loop driver (default: the library's = one persistent kernel per run on one GPU)")
    ap.add_argument("--no-sub", action="store_true", help="default workload without its sub-objects")
    ap.add_argument("--no-one-sided", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-pairs", type=float, default=1.3e11,
                    help="pairs in the cpu_baseline sample (~15 s on 16 cores)")
    ap.add_argument("--ref-pairs", type=float, default=1.5e10, help="pairs per step of --impl reference")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    global SYMMETRIC, E2E_STREAMS, REUSE, DRIVER
    REUSE = None if args.reuse is None else bool(args.reuse)
    SYMMETRIC = bool(args.symmetric) and args.mode == "fast"
    E2E_STREAMS = max(0, args.e2e_streams)
    DRIVER = args.driver
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
