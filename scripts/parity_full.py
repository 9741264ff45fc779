"""Full-size parity of the benchmarked configuration (BASELINE configs[1], "cfg2": 2e5 stars + 5e5 gas cells),
the complete iterative loop, against the CPU oracle -- both of its accumulations:

  f64acc  the reference's float32 terms summed in float64: what north_star's 1e-6 criteria are checked against
  f32seq  the reference's own arithmetic (in-order float32 sum, particle_subroutines.f90:497-510)

Two stages, because the oracle needs ~2.3e12 pair evaluations per accumulation (minutes on many cores) and
GPU-box minutes are better spent on the GPU:

  python scripts/parity_full.py oracle      CPU only: runs the oracle loops, caches the results under
                                            profiles/_parity_cache/ (git-ignored, travels to the GPU box)
  python scripts/parity_full.py gpu [--out F]   on a B200: runs the library with the shipped defaults (FAST: symmetric
                                            self-term, cached external sums, incremental passes, persistent loop
                                            kernel) and in EXACT mode, compares, writes profiles/parity_r02.json
                                            (read by bench.py) and prints a markdown table

For every job: final mask and member list, passes, Phi max relative error over the particles bound in both, and the
mismatching particles binned by energy margin |E| / max(KE, |PE|).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

CACHE = os.path.join(ROOT, "profiles", "_parity_cache")
BINS = [0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, np.inf]
BIN_NAMES = ["<1e-7", "1e-7..1e-6", "1e-6..1e-5", "1e-5..1e-4", "1e-4..1e-3", ">=1e-3"]


def jobs(n_star=200_000, n_gas=500_000):
    from pyhalma_b200 import synth
    c = synth.config2(n_star, n_gas)
    s, g = c.stars, c.gas
    M = s.mass.sum()
    vb = (float(np.sum(s.mass * s.vx) / M), float(np.sum(s.mass * s.vy) / M), float(np.sum(s.mass * s.vz) / M))
    return [("stellar", (s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass), dict(pre=[g.pos_mass()], kappa=c.factor_v ** 2),
             c.factor_v ** 2),
            ("gas", (g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass),
             dict(post=[s.pos_mass()], split_classes=True, kappa=2.0, vb_fixed=vb), 2.0)]


def cache_path(job, variant, tag):
    return os.path.join(CACHE, "cfg2%s_%s_%s.npz" % (tag, job, variant))


def stage_oracle(tag, sizes):
    from oracle import oracle as O
    os.makedirs(CACHE, exist_ok=True)
    for job, members, kw, _ in jobs(*sizes):
        for variant in ("f64acc", "f32seq"):
            path = cache_path(job, variant, tag)
            if os.path.exists(path):
                print("cached:", path, flush=True)
                continue
            t0 = time.time()
            o = O.unbind_halo(*members, variant=variant, **kw)
            np.savez_compressed(path, mask=np.packbits(o.mask), n=len(o.mask), be32=o.be32, energy=o.energy,
                                n_iter=o.n_iter, vb=np.asarray(o.vb), mass=o.mass, pairs=o.pairs,
                                seconds=time.time() - t0)
            print("%s %s: %d passes, %d bound of %d, %.0f s" % (job, variant, o.n_iter, int(o.mask.sum()), len(o.mask),
                                                                time.time() - t0), flush=True)


def binned(margin):
    h, _ = np.histogram(margin, BINS)
    return {k: int(v) for k, v in zip(BIN_NAMES, h)}


def compare(r, o, kappa, O):
    mask_o = np.unpackbits(o["mask"])[:int(o["n"])].astype(bool)
    be_o, en_o = o["be32"], o["energy"]
    both = r.mask & mask_o
    rel = float(np.abs(r.be32[both].astype(np.float64) / be_o[both].astype(np.float64) - 1).max()) if both.any() else 0.0
    diff = r.mask != mask_o
    margin = O.energy_margin(en_o, be_o, kappa)
    outside = int(np.count_nonzero(margin[diff] >= 1e-6))
    return {"phi_max_rel_err_bound_in_both": rel, "mask_mismatches": int(diff.sum()),
            "mask_mismatches_by_margin": binned(margin[diff]), "mask_mismatches_outside_1e-6": outside,
            "idx_identical": bool(np.array_equal(r.idx, np.flatnonzero(mask_o))),
            "passes_gpu": int(r.n_iter), "passes_oracle": int(o["n_iter"]),
            "bit_identical": bool(not diff.any() and np.array_equal(r.be32.view(np.uint32), be_o.view(np.uint32))),
            "vb_max_rel_err": float(np.max(np.abs(np.asarray(r.vb) - o["vb"]) / np.maximum(np.abs(o["vb"]), 1e-300))),
            "mass_rel_err": float(abs(r.mass - float(o["mass"])) / max(abs(float(o["mass"])), 1e-300))}


def stage_gpu(tag, sizes, out_json):
    from oracle import oracle as O
    from pyhalma_b200.unbind import unbind_halo
    report = {"workload": "cfg2%s: NFW galaxy halo, %d stars + %d gas cells, the complete iterative loop of both jobs"
                          % (tag, sizes[0], sizes[1]),
              "criteria": "north_star: mask and member indices identical outside |E|/max(KE,|PE|) < 1e-6; potentials and "
                          "derived quantities within 1e-6 relative (checked against f64acc); the differences against the "
                          "reference's own float32 in-order sum (f32seq) are its rounding drift, reported by margin",
              "jobs": {}}
    ok = True
    for job, members, kw, kappa in jobs(*sizes):
        o64 = np.load(cache_path(job, "f64acc", tag))
        o32 = np.load(cache_path(job, "f32seq", tag))
        row = {"n": len(members[0]), "oracle_seconds": {"f64acc": float(o64["seconds"]), "f32seq": float(o32["seconds"])}}
        t0 = time.time()
        fast = unbind_halo(*members, mode="fast", **kw)
        row["fast_ms"] = fast.stats.total_ms
        row["fast"] = {"options": "defaults: symmetric self-term, external sums cached, incremental passes, driver %d"
                                  % fast.stats.driver,
                       "vs_f64acc": compare(fast, o64, kappa, O), "vs_f32seq": compare(fast, o32, kappa, O)}
        ok &= row["fast"]["vs_f64acc"]["mask_mismatches_outside_1e-6"] == 0
        ok &= row["fast"]["vs_f64acc"]["phi_max_rel_err_bound_in_both"] < 1e-6
        exact = unbind_halo(*members, mode="exact", **kw)
        row["exact_ms"] = exact.stats.total_ms
        row["exact"] = {"vs_f32seq": compare(exact, o32, kappa, O)}
        ok &= row["exact"]["vs_f32seq"]["bit_identical"]
        mask64 = np.unpackbits(o64["mask"])[:int(o64["n"])].astype(bool)
        mask32 = np.unpackbits(o32["mask"])[:int(o32["n"])].astype(bool)
        bothb = mask64 & mask32
        row["reference_drift_f32seq_vs_f64acc"] = {
            "phi_max_rel": float(np.abs(o32["be32"][bothb].astype(np.float64) / o64["be32"][bothb] - 1).max()),
            "mask_mismatches": int((mask64 != mask32).sum()),
            "by_margin": binned(O.energy_margin(o64["energy"], o64["be32"], kappa)[mask64 != mask32])}
        row["gpu_seconds"] = time.time() - t0
        report["jobs"][job] = row
    report["pass"] = bool(ok)
    report["how"] = "scripts/parity_full.py (oracle stage on the CPU, gpu stage on a B200)"
    json.dump(report, open(out_json, "w"), indent=1)
    # markdown
    print("# Full-size parity, %s\n" % report["workload"])
    print("| job | mode | against | passes GPU / oracle | Phi max rel err | mask mismatches (by margin %s) | outside 1e-6 | "
          "idx identical | bit-identical |" % " / ".join(BIN_NAMES))
    print("|---|---|---|---|---|---|---|---|---|")
    for job, row in report["jobs"].items():
        for mode, ref in (("fast", "vs_f64acc"), ("fast", "vs_f32seq"), ("exact", "vs_f32seq")):
            c = row[mode][ref]
            print("| %s (%d) | %s | %s | %d / %d | %.2e | %d (%s) | %d | %s | %s |" % (
                job, row["n"], mode.upper(), ref[3:], c["passes_gpu"], c["passes_oracle"], c["phi_max_rel_err_bound_in_both"],
                c["mask_mismatches"], " / ".join(str(v) for v in c["mask_mismatches_by_margin"].values()),
                c["mask_mismatches_outside_1e-6"], "yes" if c["idx_identical"] else "no",
                "yes" if c["bit_identical"] else "no"))
        d = row["reference_drift_f32seq_vs_f64acc"]
        print("| %s | (reference's own float32 drift) | f32seq vs f64acc | | %.2e | %d (%s) | | | |" % (
            job, d["phi_max_rel"], d["mask_mismatches"], " / ".join(str(v) for v in d["by_margin"].values())))
    print("\nFAST %.0f + %.0f ms, EXACT %.0f + %.0f ms on the GPU; the oracle took %s s on the CPU.  pass = %s" % (
        report["jobs"]["stellar"]["fast_ms"], report["jobs"]["gas"]["fast_ms"], report["jobs"]["stellar"]["exact_ms"],
        report["jobs"]["gas"]["exact_ms"],
        " + ".join("%.0f" % v for r in report["jobs"].values() for v in r["oracle_seconds"].values()), report["pass"]))
    return 0 if ok else 1


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else "gpu"
    small = "--small" in sys.argv          # 1/5 of the sizes: a quick end-to-end check of the script itself
    tag, sizes = ("_fifth", (40_000, 100_000)) if small else ("", (200_000, 500_000))
    if stage == "oracle":
        stage_oracle(tag, sizes)
    else:
        out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(ROOT, "profiles", "parity_r02%s.json" % tag)
        sys.exit(stage_gpu(tag, sizes, out))
