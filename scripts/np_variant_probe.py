"""One-pass throughput of the plan's potential kernel (predicate-free path) for the shape in
HALMA_FAST_VARIANT, on a pure-star halo and on the cfg2 gas job (lattice corrections)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pyhalma_b200 import synth
from pyhalma_b200.unbind import UnbindPlan

v = os.environ.get("HALMA_FAST_VARIANT", "auto")
line = ["variant=%s" % v]
n = int(os.environ.get("PROBE_N", 600_000))
p = synth.plummer_stars(n, 30e-3, 1e6, np.random.default_rng(0))
with UnbindPlan(np.array([0, n], np.int64), mode="fast", max_iter=1) as plan:
    plan.upload_members(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)
    best = 1e30
    for _ in range(4):
        st = plan.run()
        best = min(best, st.potential_ms)
    line.append("stars %d^2: %.2f ms %.0f G/s" % (n, best, st.pairs / best / 1e6))
c = synth.config2()
g, s = c.gas, c.stars
vb = [150.0, -80.0, 40.0]
with UnbindPlan(np.array([0, len(g)], np.int64), [np.array([0, len(s)], np.int64)], mode="fast", max_iter=1,
                split_classes=True, vb_fixed=True, kappa=2.0) as plan:
    plan.upload_members(g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass)
    plan.upload_group(0, s.mass, s.x, s.y, s.z)
    plan.set_vb(vb)
    best = 1e30
    for _ in range(4):
        st = plan.run()
        best = min(best, st.potential_ms)
    line.append("cfg2 gas %dx%d: %.2f ms %.0f G/s" % (len(g), len(g) + len(s), best, st.pairs / best / 1e6))
print("  ".join(line), flush=True)
