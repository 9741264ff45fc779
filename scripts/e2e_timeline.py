"""Where the one-shot catalogue call (host buffers in and out) spends its time: host-side phases of a single plan
and, under `ncu --metrics gpu__time_duration.sum`, the kernels it launches.
    python scripts/e2e_timeline.py [cfg3|cfg2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
jobs, desc, _ = bench.make_workload(wl, 0, 1)
jobs = bench.pin_jobs(jobs)
for rep in range(reps):
    tot = {}
    for job in jobs:
        t = [time.perf_counter()]
        plan = bench.make_plan(job, "fast", 0, upload=False); t.append(time.perf_counter())
        bench.upload_job(plan, job); t.append(time.perf_counter())
        st = plan.run(); t.append(time.perf_counter())
        res = plan.download(); t.append(time.perf_counter())
        plan.close(); t.append(time.perf_counter())
        names = ["create", "upload(enqueue)", "run(wall)", "download", "close"]
        for n, a, b in zip(names, t[:-1], t[1:]):
            tot[n] = tot.get(n, 0) + (b - a) * 1e3
        tot["run_device_ms"] = tot.get("run_device_ms", 0) + st.total_ms
        tot["loop_kernel_ms"] = tot.get("loop_kernel_ms", 0) + st.loop_ms
        tot["launches"] = tot.get("launches", 0) + st.launches
    tot["total"] = sum(v for k, v in tot.items() if k in ("create", "upload(enqueue)", "run(wall)", "download", "close"))
    print("rep", rep, {k: round(v, 2) for k, v in tot.items()}, flush=True)
