"""End-to-end (host buffers in and out) cost of ONE rank's share of the cfg3 catalogue on one GPU, for 1..4 overlapped
parts, next to the resident time of the same share and the host-side phases of a single plan:
    python scripts/e2e_share.py [world=8] [rank=0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
jobs, desc, _ = bench.make_workload("cfg3", rank, world)
jobs = bench.pin_jobs(jobs)
plan = bench.make_plan(jobs[0], "fast", 0)
for _ in range(4):
    st = plan.run()
print("share %d/%d: %d haloes, %d members; resident %.3f ms (loop %.3f)" % (rank, world, len(jobs[0]["offsets"]) - 1,
      len(jobs[0]["members"][0]), st.total_ms, st.loop_ms))
plan.close()
for streams in (1, 2, 3, 4):
    bench.E2E_STREAMS = streams
    ts = []
    for rep in range(8):
        t0 = time.perf_counter()
        bench.e2e_step(jobs, "fast", 0)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("parts %d: e2e ms/run %s  best %.3f median %.3f" % (streams, [round(t, 2) for t in ts], min(ts), sorted(ts)[len(ts) // 2]))
for rep in range(4):
    job = jobs[0]
    t = [time.perf_counter()]
    plan = bench.make_plan(job, "fast", 0, upload=False); t.append(time.perf_counter())
    bench.upload_job(plan, job); t.append(time.perf_counter())
    st = plan.run(); t.append(time.perf_counter())
    res = plan.download(); t.append(time.perf_counter())
    plan.close(); t.append(time.perf_counter())
    names = ["create", "upload(enqueue)", "run(wall)", "download", "close"]
    print("single plan rep", rep, {n: round((b - a) * 1e3, 3) for n, a, b in zip(names, t[:-1], t[1:])},
          "device total %.3f loop %.3f" % (st.total_ms, st.loop_ms))
