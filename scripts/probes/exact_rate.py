"""Throughput of the EXACT potential kernel (one pass of a plan)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pyhalma_b200 import synth
from pyhalma_b200.unbind import UnbindPlan
n = 300_000
p = synth.plummer_stars(n, 30e-3, 1e6, np.random.default_rng(0))
with UnbindPlan(np.array([0, n], np.int64), mode="exact", max_iter=1) as plan:
    plan.upload_members(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)
    best = 1e30
    for _ in range(3):
        st = plan.run()
        best = min(best, st.potential_ms)
print("exact %d^2: %.2f ms %.0f G/s" % (n, best, st.pairs / best / 1e6))
