"""cfg3 strong scaling emulated on ONE GPU: the catalogue run has no collective, so the time of rank r
of W is the time of partition r run alone.  Prints, per W, every partition's device time and the
efficiency T(1) / (W * max_r T(r, W)).

    python scripts/cfg3_parts.py [--worlds 1,8] [--steps 5] [--graph 0|1] [--one RANK/WORLD]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyhalma_b200 import sharding, synth  # noqa: E402
from pyhalma_b200.unbind import UnbindPlan  # noqa: E402


def run_part(cat, costs, world, rank, steps, graph):
    parts = sharding.lpt_partition(costs, world)
    off, cols = sharding.take_haloes(cat.offsets, [cat.x, cat.y, cat.z, cat.vx, cat.vy, cat.vz, cat.mass], parts[rank])
    plan = UnbindPlan(off, [], mode="fast", kappa=9.0, max_iter=64, use_graph=graph)
    plan.upload_members(*cols)
    for _ in range(3):
        plan.run()
    tot = pot = 0.0
    for _ in range(steps):
        st = plan.run()
        tot += st.total_ms
        pot += st.potential_ms
    per_pass = plan.debug_pass_us()[:st.passes]
    plan.close()
    return dict(per_pass_us=[[round(v, 1) for v in t] for t in per_pass], world=world, rank=rank, n_halo=len(parts[rank]), n=int(off[-1]), ms=tot / steps, pot_ms=pot / steps,
                passes=st.passes, launches=st.launches, pairs=st.pairs, evals=st.evaluations,
                phase_ms=[round(v, 4) for v in st.phase_ms], loop_ms=st.loop_ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", default="1,8")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--graph", type=int, default=None)
    ap.add_argument("--one", default=None, help="RANK/WORLD: run only this partition (ncu target)")
    a = ap.parse_args()
    graph = None if a.graph is None else bool(a.graph)
    cat = synth.config3()
    costs = sharding.halo_costs(cat.offsets)
    if a.one:
        r, w = [int(v) for v in a.one.split("/")]
        print(json.dumps(run_part(cat, costs, w, r, a.steps, graph)))
        return
    t1 = None
    for w in [int(v) for v in a.worlds.split(",")]:
        rows = [run_part(cat, costs, w, r, a.steps, graph) for r in range(w)]
        worst = max(r["ms"] for r in rows)
        if w == 1:
            t1 = worst
        print(json.dumps(dict(world=w, worst_ms=worst, mean_ms=float(np.mean([r["ms"] for r in rows])),
                              eff=None if t1 is None else t1 / (w * worst), rows=rows)), flush=True)


if __name__ == "__main__":
    main()
