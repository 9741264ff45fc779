"""Where the time of pipeline.halo_potential_stage goes, against the host-array path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pyhalma_b200 import gather, halo_gas, pipeline, synth

s = synth.amr_snapshot(n_levels=8, patches_per_level=4, max_cells=64, n_dm=2_000_000, n_st=1_000_000)
snap = gather.Snapshot(s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data, s.masclet_st_data)
vb = synth.BULK_V
c, R = s.centre, 0.02
T = time.perf_counter
for rep in range(3):
    t0 = T(); g = snap.gather_device(*c, R, s.rho_B, s.rete); t1 = T()
    rps = pipeline.rps_on_device(g, *vb, 1); t2 = T()
    g = snap.gather_device(*c, R, s.rho_B, s.rete, dm_heavy_min=0.9e7); t3 = T()
    mb = pipeline.most_bound_on_device(g); t4 = T()
    print("device: gather %.1f  rps %.1f  gather2 %.1f  most_bound %.1f ms   n=%d/%d/%d" %
          (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), g.n_gas, g.n_dm + g.n_dm_light, g.n_st), flush=True)
for rep in range(3):
    t0 = T(); h = snap.gather(*c, R, s.rho_B, s.rete); t1 = T()
    rps = halo_gas.RPS(*h[:16], *vb, 10**9, 8e7, 1); t2 = T()
    mb = halo_gas.most_bound_particle(h[0], h[1], h[2], h[6], *h[8:12], *h[12:16], h[16], 10**9, 8e7); t3 = T()
    print("host:   gather %.1f  rps %.1f  most_bound %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)), flush=True)
