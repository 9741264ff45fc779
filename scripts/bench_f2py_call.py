"""The f2py-level call (fortran_modules.particle.particle.brute_force_binding_energy -> halma_potential_f32,
HOST float32 arrays in and out) at cfg2 sizes.

  python scripts/bench_f2py_call.py            one JSON line per caller shape: direct path (predicated kernel)
                                               against the plan-backed path (predicate-free kernel)
  rps_sequence(device)                         bench.py's `f2py_level` sub-object: the calls RPS makes for one halo
                                               (halo_gas.py:306-450, exact sums: BRUTE_FORCE_LIM above every class)
                                               through the zero-edit drop-in, wall time of every call

Wall time of the whole call (H2D, sort, kernels, D2H), best of 3 after one warm-up.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

f32 = np.float32


def _case():
    from pyhalma_b200 import synth
    c = synth.config2()
    rng = synth.rng_for(2, 77)
    d = synth.dm_cloud(100_000, 0.03, 8e7, rng)
    return c.stars, c.gas, d


def _f32(p):
    return [f32(a) for a in (p.mass, p.x, p.y, p.z)]


def _best_ms(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def rps_sequence(device=0):
    """RPS's kernel calls for one cfg2-size halo through `from fortran_modules import particle` (the drop-in
    package, INTEGRATION.md level 1): gas -> gas, DM -> gas, stars -> gas, each with the reference's float32 casts
    (halo_gas.py:172-178) inside the timing, next to the fused device-resident plan that halo_gas.RPS uses."""
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    from fortran_modules import particle       # the drop-in for the f2py extension
    from pyhalma_b200 import _lib, halo_gas
    s, g, d = _case()
    mb = _lib.microbench(device)
    peak = mb["rsq_gops"]
    calls = []
    total_ms = total_pairs = 0.0
    for name, src in (("gas -> gas (halo_gas.py:306-328)", g), ("DM -> gas (:396-421)", d), ("stars -> gas (:429-450)", s)):
        def run():
            cast = [np.asarray(a).astype(f32) for a in (src.mass, src.x, src.y, src.z, g.x, g.y, g.z)]
            return particle.particle.brute_force_binding_energy(np.int32(1), np.int32(len(src)), *cast[:4],
                                                                np.int32(len(g)), *cast[4:], device=device)
        ms = _best_ms(run)
        pairs = len(src) * len(g)
        calls.append({"call": name, "n_src": len(src), "n_tgt": len(g), "ms": ms,
                      "Ginteractions_per_s": pairs / ms / 1e6, "frac_of_mufu_roofline": pairs / ms / 1e6 / peak})
        total_ms += ms
        total_pairs += pairs
    vx, vy, vz = (float(np.average(a, weights=s.mass)) for a in (s.vx, s.vy, s.vz))
    args = (g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass, np.full(len(g), 1e4), d.x, d.y, d.z, d.mass, s.x, s.y, s.z, s.mass,
            vx, vy, vz, 10 ** 9, 8e7, 1)
    fused_ms = _best_ms(lambda: halo_gas.RPS(*args, device=device))
    unfused_ms = _best_ms(lambda: halo_gas.RPS(*args, device=device, fused=False))
    return {"workload": "RPS's three kernel calls for one cfg2-size halo (5e5 gas cells; sources 5e5 gas, 1e5 DM, 2e5 stars), "
                        "host float32 arrays in and out of every call, reference casts inside the timing",
            "calls": calls, "sequence_ms": total_ms, "Ginteractions_per_s": total_pairs / total_ms / 1e6,
            "frac_of_mufu_roofline": total_pairs / total_ms / 1e6 / peak,
            "RPS_unfused_ms": unfused_ms, "RPS_fused_plan_ms": fused_ms,
            "note": "every call goes through a one-pass plan (predicate-free kernel; self calls also the symmetric "
                    "self-term, so their interactions/s exceed the one-sided roofline); RPS_unfused = "
                    "pyhalma_b200.halo_gas.RPS(fused=False), the reference's call structure incl. energy step and "
                    "mass sums on the host; RPS_fused_plan = the same outputs from one device-resident plan"}


def main():
    from pyhalma_b200 import particle as particle_mod
    particle = particle_mod if hasattr(particle_mod, "particle") else __import__("pyhalma_b200.particle", fromlist=["particle"])
    s, g, _ = _case()
    cat = lambda k: f32(np.concatenate([getattr(g, k), getattr(s, k)]))       # noqa: E731
    S, G = _f32(s), _f32(g)
    GS = [cat(k) for k in ("mass", "x", "y", "z")]
    cases = [("self: gas -> gas (RPS, halo_gas.py:306-328)", G, G[1:]),
             ("block: concat(gas, stars) -> stars (halo_properties.py:333-339)", GS, S[1:]),
             ("cross: stars -> gas (RPS, halo_gas.py:429-450)", S, G[1:]),
             ("cross: gas -> stars (most_bound_particle, halo_gas.py:517-533)", G, S[1:])]

    def call(src, tgt):
        return particle.particle.brute_force_binding_energy(1, len(src[0]), *src, len(tgt[0]), *tgt, mode="fast")

    for name, src, tgt in cases:
        out = {"case": name, "n_src": len(src[0]), "n_tgt": len(tgt[0]), "pairs": len(src[0]) * len(tgt[0])}
        res = {}
        for label, env in (("direct", "0"), ("plan", "1")):
            os.environ["HALMA_POT_PLAN_MIN_PAIRS"] = env
            res[label] = call(src, tgt)
            ms = _best_ms(lambda: call(src, tgt))
            out[label + "_ms"] = ms
            out[label + "_Ginteractions_per_s"] = out["pairs"] / ms / 1e6
        out["max_rel_diff"] = float(np.max(np.abs(res["plan"].astype(np.float64) / res["direct"] - 1)))
        out["speedup"] = out["direct_ms"] / out["plan_ms"]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
