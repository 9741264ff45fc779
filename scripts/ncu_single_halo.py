"""Profiling target: ONE potential pass (max_iter = 1) over one Plummer halo of NCU_N stars, symmetric and one-sided
(two plans, two launches each: the first builds the sorted copies)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyhalma_b200 import synth
from pyhalma_b200.unbind import UnbindPlan
n = int(os.environ.get("NCU_N", "1000000"))
p = synth.config4(n)
for sym in (True, False):
    with UnbindPlan(np.array([0, n], np.int64), [], mode="fast", kappa=9.0, max_iter=1, symmetric=sym,
                    cache_external=False, incremental=False) as plan:
        plan.upload_members(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)
        plan.run()
        st = plan.run()
        print("symmetric" if sym else "one-sided", "pass ms %.2f potential ms %.2f evals %d" % (st.total_ms, st.potential_ms, st.evaluations))
