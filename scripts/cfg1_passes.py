"""cfg1 (1e4 stars + 1e4 gas cells): per-pass phase times of the two jobs with the persistent loop kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
jobs, desc, _ = bench.make_workload("cfg1", 0, 1)
for job in jobs:
    plan = bench.make_plan(job, "fast", 0)
    for _ in range(3):
        st = plan.run()
    pp = plan.debug_pass_us()[:st.passes]
    print(job["kind"], "total %.1f us, passes %d, phases %s" % (st.total_ms * 1e3, st.passes, [round(v * 1e3, 1) for v in st.phase_ms]))
    for k, t in enumerate(pp):
        print("   pass %d: potential %.1f  energy+compaction %.1f  tables %.1f us" % (k, *t))
    plan.close()
