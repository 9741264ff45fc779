"""Host-side cost of creating / filling / running / destroying a plan, by halo size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pyhalma_b200 import synth
from pyhalma_b200.unbind import UnbindPlan
T = time.perf_counter
for n in (1000, 10_000, 100_000, 800_000):
    p = synth.plummer_stars(n, 2e-3 * (n / 1e4) ** (1 / 3), 1e6, np.random.default_rng(0))
    acc = np.zeros(6)
    reps = 20 if n < 500_000 else 5
    for r in range(reps + 2):
        t0 = T(); plan = UnbindPlan(np.array([0, n], np.int64), mode="fast", max_iter=1)
        t1 = T(); plan.upload_members(p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)
        t2 = T(); st = plan.run()
        t3 = T(); plan.download()
        t4 = T(); plan.close()
        t5 = T()
        if r >= 2:
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, st.total_ms / 1e3]
    acc *= 1e3 / reps
    print("n=%7d  create %.2f  upload %.2f  run %.2f (device %.2f)  download %.2f  close %.2f ms" %
          (n, acc[0], acc[1], acc[2], acc[5], acc[3], acc[4]), flush=True)
