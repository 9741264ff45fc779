"""Short profiling target: one stellar + one gas unbinding of a cfg2-like halo (reduced so
that ncu's replays stay short), fast mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyhalma_b200 import synth
from pyhalma_b200.unbind import unbind_halo

n_star = int(os.environ.get("NCU_STARS", "100000"))
n_gas = int(os.environ.get("NCU_GAS", "250000"))
c = synth.config2(n_star, n_gas)
s, g = c.stars, c.gas
r = unbind_halo(s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass, pre=[g.pos_mass()], kappa=9.0, mode="fast")
print("stellar passes", r.n_iter, "bound", int(r.mask.sum()), "potential ms", r.stats.potential_ms)
M = s.mass.sum()
vb = (np.sum(s.mass * s.vx) / M, np.sum(s.mass * s.vy) / M, np.sum(s.mass * s.vz) / M)
r = unbind_halo(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass, post=[s.pos_mass()], split_classes=True, kappa=2.0,
                vb_fixed=vb, mode="fast")
print("gas passes", r.n_iter, "bound", int(r.mask.sum()), "potential ms", r.stats.potential_ms)
