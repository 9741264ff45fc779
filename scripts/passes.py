"""Per-pass phase times of a workload's jobs with the persistent loop kernel:  python scripts/passes.py cfg2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
jobs, desc, _ = bench.make_workload(wl, 0, 1)
for job in jobs:
    plan = bench.make_plan(job, "fast", 0)
    for _ in range(3):
        st = plan.run()
    pp = plan.debug_pass_us()[:st.passes]
    print(wl, job["kind"], "total %.1f us, passes %d, evals %.4g, phases %s" % (st.total_ms * 1e3, st.passes, st.evaluations, [round(v * 1e3, 1) for v in st.phase_ms]))
    for k, t in enumerate(pp):
        print("   pass %d: potential %.1f  energy+compaction %.1f  tables %.1f us" % (k, *t))
    plan.close()
