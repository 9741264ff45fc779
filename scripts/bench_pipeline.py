"""Wall time of the potential stage of pyHALMA's halo loop (pyHALMA.py:1023-1067: gather +
RPS + most_bound_particle per halo) over a list of haloes of one resident snapshot:

  device    pipeline.halo_potential_stage: nothing but scalars returns to the host
  host      the drop-in functions with the reference's signatures and host arrays
            (st_gas_dm_particles_inside -> RPS -> most_bound_particle, fused plans)
  cpu       the oracle (numpy gather + C/OpenMP potential) on the first few haloes

    python scripts/bench_pipeline.py [--haloes 200] [--cpu-haloes 6] > gpurun_out/bench_pipeline.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pyhalma_b200 import _lib, gather, halo_gas, pipeline, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--haloes", type=int, default=200)
    ap.add_argument("--cpu-haloes", type=int, default=40)
    ap.add_argument("--cpu-seconds", type=float, default=25.0)
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--repeats", type=int, default=2)
    a = ap.parse_args()
    from oracle import gather as OG
    from oracle import oracle as O
    _lib.require_device(0)
    s = synth.amr_snapshot(n_levels=8, patches_per_level=4, max_cells=64, n_dm=2_000_000, n_st=1_000_000)
    rng = np.random.default_rng(11)
    centres = np.asarray(s.centre) + rng.normal(0, 0.004, (a.haloes, 3))
    radii = 10 ** rng.uniform(np.log10(0.006), np.log10(0.03), a.haloes)
    vb = synth.BULK_V
    mdm, nsp = 8e7, 1
    t0 = time.perf_counter()
    snap = gather.Snapshot(s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data, s.masclet_st_data)
    upload_s = time.perf_counter() - t0
    # untimed pass over the whole list: the stream-ordered memory pool grows to the working set of
    # the largest haloes (first-touch cudaMalloc costs hundreds of ms and is paid once per process)
    for c, r in zip(centres, radii):
        pipeline.halo_potential_stage(snap, *c, r, s.rho_B, s.rete, *vb, mdm, nsp, mode=a.mode)

    def device_pass():
        t0 = time.perf_counter()
        out = [pipeline.halo_potential_stage(snap, *c, r, s.rho_B, s.rete, *vb, mdm, nsp, mode=a.mode)
               for c, r in zip(centres, radii)]
        return out, time.perf_counter() - t0

    def host_one(c, r):
        g = halo_gas.st_gas_dm_particles_inside(s.rete, s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data,
                                                s.masclet_st_data, None, None, *c, r, s.rho_B, mass_to_sun=1.0)
        rps = halo_gas.RPS(*g[:16], *vb, 10 ** 9, mdm, nsp, mode=a.mode)
        mb = halo_gas.most_bound_particle(g[0], g[1], g[2], g[6], *g[8:12], *g[12:16], g[16], 10 ** 9, mdm,
                                          mode=a.mode) if len(g[12]) else None
        return rps, mb, (len(g[0]), len(g[8]), len(g[12]))

    def host_pass():
        t0 = time.perf_counter()
        out = [host_one(c, r) for c, r in zip(centres, radii)]
        return out, time.perf_counter() - t0

    host_one(centres[0], radii[0])          # uploads the cached snapshot of the drop-in path
    # alternate the two paths and keep the best pass of each (pool and clock state are shared)
    dev, dev_s = device_pass()
    host, host_s = host_pass()
    passes = {"device": [round(dev_s, 3)], "host": [round(host_s, 3)]}
    for _ in range(a.repeats - 1):
        d2, t = device_pass()
        passes["device"].append(round(t, 3))
        dev_s = min(dev_s, t)
        h2, t = host_pass()
        passes["host"].append(round(t, 3))
        host_s = min(host_s, t)
    snap.close()
    gather.release_cached_snapshot()

    same = all(np.allclose(d[0], h[0], rtol=1e-12) and (d[1] is None) == (h[1] is None)
               and (d[1] is None or d[1][3] == h[1][3]) for d, h in zip(dev, host))
    sizes = np.array([h[2] for h in host])
    pairs = float(np.sum(sizes[:, 0] * sizes.sum(1) + sizes[:, 2] * sizes.sum(1)))

    k = 0
    t0 = time.perf_counter()
    order = np.argsort(radii)            # smallest first; stop after --cpu-seconds
    for c, r in zip(centres[order], radii[order]):
        if k >= a.cpu_haloes or time.perf_counter() - t0 > a.cpu_seconds:
            break
        k += 1
        w = OG.st_gas_dm_particles_inside(s.rete, s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data,
                                          s.masclet_st_data, None, None, *c, r, s.rho_B)
        O.RPS(*w[:16], *vb, 10 ** 9, mdm, nsp)
        if len(w[12]):
            O.most_bound_particle(w[0], w[1], w[2], w[6], *w[8:12], *w[12:16], w[16], 10 ** 9, mdm)
    cpu_s = time.perf_counter() - t0
    sk = sizes[order[:k]]
    cpu_pairs = float(np.sum(sk[:, 0] * sk.sum(1) + sk[:, 2] * sk.sum(1)))

    print(json.dumps({
        "what": "potential stage of the halo loop (gather + RPS + most_bound_particle per halo), one B200",
        "haloes": a.haloes, "mode": a.mode,
        "snapshot": {"cells": s.n_cells, "n_dm": len(s.masclet_dm_data[0]), "n_st": len(s.masclet_st_data[0]),
                     "upload_s": round(upload_s, 3)},
        "mean_selected": {"gas": float(sizes[:, 0].mean()), "dm": float(sizes[:, 1].mean()),
                          "stars": float(sizes[:, 2].mean())},
        "interactions": pairs,
        "device_resident": {"wall_s": round(dev_s, 4), "ms_per_halo": round(1e3 * dev_s / a.haloes, 3),
                            "Ginteractions_per_s": round(pairs / dev_s / 1e9, 1)},
        "host_arrays": {"wall_s": round(host_s, 4), "ms_per_halo": round(1e3 * host_s / a.haloes, 3),
                        "Ginteractions_per_s": round(pairs / host_s / 1e9, 1)},
        "cpu_port": {"haloes": k, "which": "the k smallest radii", "wall_s": round(cpu_s, 3), "ms_per_halo": round(1e3 * cpu_s / k, 1),
                     "Ginteractions_per_s": round(cpu_pairs / cpu_s / 1e9, 2), "threads": O.max_threads()},
        "device_equals_host_path": bool(same), "pass_wall_s": passes,
    }), flush=True)


if __name__ == "__main__":
    main()
