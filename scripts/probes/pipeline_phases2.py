import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pyhalma_b200 import gather, halo_gas, pipeline, synth

s = synth.amr_snapshot(n_levels=8, patches_per_level=4, max_cells=64, n_dm=2_000_000, n_st=1_000_000)
rng = np.random.default_rng(11)
n = 14
centres = np.asarray(s.centre) + rng.normal(0, 0.004, (200, 3))
radii = 10 ** rng.uniform(np.log10(0.006), np.log10(0.03), 200)
snap = gather.Snapshot(s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data, s.masclet_st_data)
vb = synth.BULK_V
T = time.perf_counter
for order in ("device", "host", "device", "host"):
    tot = 0
    line = []
    for c, R in zip(centres[:n], radii[:n]):
        t0 = T()
        if order == "device":
            g = snap.gather_device(*c, R, s.rho_B, s.rete); t1 = T()
            pipeline.rps_on_device(g, *vb, 1); t2 = T()
            g = snap.gather_device(*c, R, s.rho_B, s.rete, dm_heavy_min=0.9e7)
            pipeline.most_bound_on_device(g); t3 = T()
        else:
            h = snap.gather(*c, R, s.rho_B, s.rete); t1 = T()
            halo_gas.RPS(*h[:16], *vb, 10**9, 8e7, 1); t2 = T()
            halo_gas.most_bound_particle(h[0], h[1], h[2], h[6], *h[8:12], *h[12:16], h[16], 10**9, 8e7); t3 = T()
        line.append("%.0f/%.0f/%.0f" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
        tot += t3 - t0
    print(order, "%.2f s" % tot, " ".join(line), flush=True)
