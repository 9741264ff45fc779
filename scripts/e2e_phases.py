"""Where does the one-shot (host buffers) path spend its time?  cfg2 jobs, phase by phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

import sys as _s
wl = _s.argv[1] if len(_s.argv) > 1 else "cfg2"
jobs, desc, _ = bench.make_workload(wl, 0, 1)
jobs = [dict(j, members=tuple(bench.pinned_copy(a) for a in j["members"]),
             groups=[(g[0],) + tuple(bench.pinned_copy(a) for a in g[1:]) for g in j["groups"]]) for j in jobs]
for rep in range(4):
    tot = {}
    for job in jobs:
        t = [time.perf_counter()]
        plan = bench.make_plan(job, "fast", 0, upload=False); t.append(time.perf_counter())
        bench.upload_job(plan, job); t.append(time.perf_counter())
        st = plan.run(); t.append(time.perf_counter())
        res = plan.download(); t.append(time.perf_counter())
        plan.close(); t.append(time.perf_counter())
        names = ["create", "upload(enqueue)", "run", "download", "close"]
        for n, a, b in zip(names, t[:-1], t[1:]):
            tot[n] = tot.get(n, 0) + (b - a) * 1e3
        tot["run_device_ms"] = tot.get("run_device_ms", 0) + st.total_ms
    print("rep", rep, {k: round(v, 2) for k, v in tot.items()}, flush=True)
