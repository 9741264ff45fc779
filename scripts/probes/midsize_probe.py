"""Per-pass time of one mid-size stellar halo under the current thresholds (env decides the path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pyhalma_b200 import synth
from pyhalma_b200.unbind import UnbindPlan
tag = "NPmin=%s variant=%s sym=%s" % (os.environ.get("HALMA_NP_MIN_PAIRS", "default"), os.environ.get("HALMA_FAST_VARIANT", "auto"),
                                     os.environ.get("HALMA_SYMMETRIC", "1"))
out = [tag]
for n in [int(v) for v in os.environ.get("PROBE_SIZES", "10000,20000,40000,60000").split(",")]:
    p = synth.plummer_stars(n, 2e-3 * (n / 1e4) ** (1 / 3), 1e6, np.random.default_rng(0))
    with UnbindPlan(np.array([0, n], np.int64), mode="fast", max_iter=1) as plan:
        plan.upload_members(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)
        best = tot = 1e30
        for _ in range(5):
            st = plan.run()
            best = min(best, st.potential_ms); tot = min(tot, st.total_ms)
    out.append("n=%d pot %.3f ms total %.3f ms (%.0f G/s, evals/pairs %.2f)" % (n, best, tot, st.pairs / tot / 1e6, st.evaluations / st.pairs))
print("  ".join(out), flush=True)
