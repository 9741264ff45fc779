"""Measured MUFU.RSQ work of a profiled potential launch: sums the per-instruction counters of the SASS source page
of an .ncu-rep (captured with --import-source on) over the MUFU.RSQ instructions, and splits the warp-stall samples
by instruction class.  Cross-check of halma_run_stats.evaluations, which is an analytic count.

    python scripts/ncu_mufu_count.py gpurun_out/prof.ncu-rep [analytic_evaluations]
"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep = sys.argv[1]
    analytic = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    kernel = lines[start - 1] if start else ""
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    tot = {"mufu_rsq_thread": 0, "mufu_rsq_thread_pred_on": 0, "mufu_rsq_warp": 0, "all_warp": 0, "samples": 0}
    cls = {}
    for r in rows[1:]:
        if r and r[0] == "Address":
            break          # the next kernel of the report
        if len(r) < len(hdr) or not r[col["Source"]].strip():
            continue
        ins = r[col["Source"]].strip()
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        n_warp = int(r[col["Instructions Executed"]] or 0)
        n_thr = int(r[col["Thread Instructions Executed"]] or 0)
        n_on = int(r[col["Predicated-On Thread Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        tot["all_warp"] += n_warp
        tot["samples"] += smp
        if op.startswith("MUFU.RSQ"):
            tot["mufu_rsq_thread"] += n_thr
            tot["mufu_rsq_thread_pred_on"] += n_on
            tot["mufu_rsq_warp"] += n_warp
        key = op.split(".")[0]
        c = cls.setdefault(key, [0, 0])
        c[0] += n_warp
        c[1] += smp
    res = {"report": rep, "kernel": kernel.strip('"').split('","')[-1][:120], **tot,
           "lanes_per_mufu_warp_instruction": tot["mufu_rsq_thread"] / max(tot["mufu_rsq_warp"], 1),
           "top_opcodes_by_issue": sorted(((k, v[0], v[1]) for k, v in cls.items()), key=lambda t: -t[1])[:14]}
    if analytic:
        res["analytic_evaluations"] = analytic
        res["measured_over_analytic"] = tot["mufu_rsq_thread_pred_on"] / analytic
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
