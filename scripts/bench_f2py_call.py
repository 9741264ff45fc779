"""The f2py-level call (particle.brute_force_binding_energy -> halma_potential_f32, HOST float32 arrays in and
out) at cfg2 sizes: the direct path (predicated kernel) against the plan-backed path
(HALMA_POT_PLAN_MIN_PAIRS), for the three shapes the reference's callers have.  Wall time of the whole call
(H2D, kernels, D2H), best of 3 after one warm-up.  One JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pyhalma_b200 import particle, synth

f32 = np.float32
c = synth.config2()
s, g = c.stars, c.gas
cat = lambda k, *ps: f32(np.concatenate([getattr(p, k) for p in ps]))       # noqa: E731
S = [f32(a) for a in (s.mass, s.x, s.y, s.z)]
G = [f32(a) for a in (g.mass, g.x, g.y, g.z)]
GS = [cat(k, g, s) for k in ("mass", "x", "y", "z")]
cases = [("self: gas -> gas (RPS, halo_gas.py:306-328)", G, G[1:]),
         ("block: concat(gas, stars) -> stars (halo_properties.py:333-339)", GS, S[1:]),
         ("cross: stars -> gas (RPS, halo_gas.py:429-450)", S, G[1:])]


def call(src, tgt):
    return particle.brute_force_binding_energy(1, len(src[0]), *src, len(tgt[0]), *tgt, mode="fast")


for name, src, tgt in cases:
    out = {"case": name, "n_src": len(src[0]), "n_tgt": len(tgt[0]), "pairs": len(src[0]) * len(tgt[0])}
    res = {}
    for label, env in (("direct", "0"), ("plan", "1")):
        os.environ["HALMA_POT_PLAN_MIN_PAIRS"] = env
        res[label] = call(src, tgt)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            call(src, tgt)
            best = min(best, time.perf_counter() - t0)
        out[label + "_ms"] = best * 1e3
        out[label + "_Ginteractions_per_s"] = out["pairs"] / best / 1e9
    out["max_rel_diff"] = float(np.max(np.abs(res["plan"].astype(np.float64) / res["direct"] - 1)))
    out["speedup"] = out["direct_ms"] / out["plan_ms"]
    print(json.dumps(out), flush=True)
