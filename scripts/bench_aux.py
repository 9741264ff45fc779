"""Measurement of the rows next to the hot path: the snapshot
gather, halo_shape and sigma_projections.  All three are HBM-bound; each line reports the
CUDA-event time of the kernels (halma_last_kernel_ms), the algorithmic bytes they move,
the resulting GB/s against the measured HBM peak, the wall time of the public call (host
buffers in, host arrays out) and the CPU oracle timed on the same input.

    python scripts/bench_aux.py [--quick] > gpurun_out/bench_aux.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pyhalma_b200 import _lib, gather, synth  # noqa: E402
from pyhalma_b200.particle import particle  # noqa: E402


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6500.0, "fallback (B200_PROFILING.md)"


def best(fn, reps):
    out, wall, kern = None, 1e30, 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        wall = min(wall, time.perf_counter() - t0)
        kern = min(kern, float(_lib.lib().halma_last_kernel_ms()))
    return out, wall * 1e3, kern


def line(op, cfg, kern_ms, wall_ms, bytes_alg, cpu_ms, cpu_kind, extra=None):
    peak, src = hbm_peak()
    gbs = bytes_alg / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    d = {"op": op, "config": cfg, "kernel_ms": round(kern_ms, 4), "call_wall_ms": round(wall_ms, 3),
         "algorithmic_bytes": int(bytes_alg),
         "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                      "frac": round(gbs / peak, 4), "peak_source": src},
         "cpu_baseline": {"ms": round(cpu_ms, 2), "kind": cpu_kind, "speedup_vs_call": round(cpu_ms / wall_ms, 1)}}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    from oracle import gather as OG
    from oracle import oracle as O
    _lib.require_device(0)
    reps = 3 if a.quick else 7

    # ---- gather -----------------------------------------------------------------------------
    kw = (dict(n_levels=6, patches_per_level=3, max_cells=40, n_dm=400_000, n_st=300_000) if a.quick else
          dict(n_levels=8, patches_per_level=6, max_cells=96, n_dm=20_000_000, n_st=10_000_000))
    kw["background_frac"] = 0.97          # a cosmological box: 3 % of the particles belong to the central halo
    s = synth.amr_snapshot(**kw)
    n_dm, n_st = len(s.masclet_dm_data[0]), len(s.masclet_st_data[0])
    for path, env in (("brute force", "-1"), ("cell list", None)):
        if env is None:
            os.environ.pop("HALMA_GATHER_INDEX_MIN", None)
        else:
            os.environ["HALMA_GATHER_INDEX_MIN"] = env
        t0 = time.perf_counter()
        snap = gather.Snapshot(s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data, s.masclet_st_data)
        upload_ms = (time.perf_counter() - t0) * 1e3
        for R in (0.004, 0.02, 0.08):
            out, wall, kern = best(lambda: snap.gather(*s.centre, R, s.rho_B, s.rete), reps)
            ng, nd, ns = len(out[0]), len(out[8]), len(out[12])
            cells = ng * (2 * 2 + 20 + 64)      # flags of the selected cells only: a lower bound
            if path == "brute force":
                # every resident particle's position in the count and in the emit pass
                alg = 2 * 24 * (n_dm + n_st) + nd * 64 + ns * 80 + cells
            else:
                # candidates >= selected: index + position in both passes, index sort, row read and written
                alg = (nd + ns) * (2 * (4 + 24) + 16) + nd * 64 + ns * 80 + cells
            t0 = time.perf_counter()
            ref = OG.st_gas_dm_particles_inside(s.rete, s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data,
                                                s.masclet_st_data, None, None, *s.centre, R, s.rho_B)
            cpu = (time.perf_counter() - t0) * 1e3
            ok = all(np.array_equal(x, y) for x, y in zip(out, ref))
            line("snapshot_gather", {"ball_query": path, "cells": s.n_cells, "n_dm": n_dm, "n_st": n_st, "R_mpc": R,
                                     "selected": [ng, nd, ns]}, kern, wall, alg, cpu,
                 "port (numpy, brute-force ball query)",
                 {"bit_exact_vs_oracle": bool(ok), "snapshot_upload_ms": round(upload_ms, 1),
                  "note": "kernel_ms spans the whole gather incl. 3 host syncs per class; small selections are "
                          "latency-bound, not HBM-bound"})
        snap.close()

    # ---- halo_shape / sigma_projections ---------------------------------------------------------
    n = 400_000 if a.quick else 10_000_000
    rng = np.random.default_rng(1)
    p = (rng.normal(size=(3, n)) * np.array([[3e-3], [2e-3], [1e-3]])).astype(np.float32)
    m = rng.uniform(0.5e6, 2e6, n).astype(np.float32)
    out, wall, kern = best(lambda: particle.halo_shape(1, n, p[0], p[1], p[2], m), reps)
    t0 = time.perf_counter()
    ref = O.halo_shape(1, n, p[0], p[1], p[2], m, wide=True)
    cpu = (time.perf_counter() - t0) * 1e3
    line("halo_shape", {"npart": n}, kern, wall, 16 * n, cpu, "port (C, one thread, float64 sums)",
         {"max_rel_err_vs_oracle": float(np.abs(out / ref - 1).max())})

    v = (rng.normal(0, 60, size=(3, n))).astype(np.float32)
    v[0] -= 2e4 * p[1]
    v[1] += 2e4 * p[0]
    pl = np.arange(1, n + 1, dtype=np.int32)
    for n_cell in (25, 101):
        ll = 3e-3 * 4 / n_cell
        grid = ((np.arange(n_cell) - n_cell // 2) * ll).astype(np.float32)
        args = (1, n, grid, n_cell, pl, p[0], p[1], p[2], v[0], v[1], v[2], m, 0.0, 0.0, 0.0, 3e-3, 2.4e-3, 1.8e-3, ll)
        out, wall, kern = best(lambda: particle.sigma_projections(*args), reps)
        t0 = time.perf_counter()
        ref = O.sigma_projections(*args, wide=True)
        cpu = (time.perf_counter() - t0) * 1e3
        # per listed particle: index 4 + pos 12 + vel 12 + mass 4 in pass 1, index + vel + cell ids in
        # pass 2, index + pos + cell ids in pass 3, cell ids written once
        alg = n * (32 + (4 + 12 + 6) + (4 + 12 + 6) + 6)
        line("sigma_projections", {"npart": n, "n_cell": n_cell}, kern, wall, alg, cpu,
             "port (C, one thread, float64 sums)",
             {"max_rel_err_vs_oracle": float(np.abs(np.array(out) / np.array(ref) - 1).max())})


if __name__ == "__main__":
    main()
