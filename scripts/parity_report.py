"""Parity report: per configuration, Phi max relative error against the
float64-accumulating oracle, mask mismatches against it (must be 0 outside the 1e-6 band) and
against the reference's float32 arithmetic, binned by |E| / max(KE, |PE|).

    python scripts/parity_report.py > profiles/parity_r01.md        (on a B200)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as O
from pyhalma_b200 import synth
from pyhalma_b200.unbind import unbind_halo

BINS = [0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, np.inf]


def binned(margin):
    h, _ = np.histogram(margin, BINS)
    return " / ".join(str(int(v)) for v in h)


def report(name, members, kw, kappa):
    rows = []
    t0 = time.time()
    o64 = O.unbind_halo(*members, variant="f64acc", **kw)
    o32 = O.unbind_halo(*members, variant="f32seq", **kw)
    t_cpu = time.time() - t0
    for mode in ("exact", "fast"):
        r = unbind_halo(*members, mode=mode, **kw)
        both = r.mask & o64.mask
        rel64 = np.abs(r.be32[both].astype(np.float64) / o64.be32[both] - 1).max() if both.any() else 0.0
        b32 = r.mask & o32.mask
        rel32 = np.abs(r.be32[b32].astype(np.float64) / o32.be32[b32].astype(np.float64) - 1).max() if b32.any() else 0.0
        d64 = r.mask != o64.mask
        d32 = r.mask != o32.mask
        m64 = O.energy_margin(o64.energy, o64.be32, kappa)
        m32 = O.energy_margin(o32.energy, o32.be32, kappa)
        bit32 = bool(np.array_equal(r.be32.view(np.uint32), o32.be32.view(np.uint32)) and not d32.any())
        rows.append("| %s | %s | %d | %d / %d | %.2e | %d (%s) | %.2e | %d (%s) | %s |" % (
            name, mode, len(members[0]), r.n_iter, o32.n_iter, rel64, int(d64.sum()), binned(m64[d64]), rel32,
            int(d32.sum()), binned(m32[d32]), "yes" if bit32 else "no"))
    drift = np.abs(o32.be32[o32.mask & o64.mask].astype(np.float64) / o64.be32[o32.mask & o64.mask] - 1).max()
    rows.append("| %s | (reference f32seq vs f64acc) | | %d / %d | %.2e | %d (%s) | | | |" % (
        name, o32.n_iter, o64.n_iter, drift, int((o32.mask != o64.mask).sum()),
        binned(O.energy_margin(o64.energy, o64.be32, kappa)[o32.mask != o64.mask])))
    return rows, t_cpu


def main():
    print("# Parity report, round 1\n")
    print("Bins of |E|/max(KE,|PE|) for mismatching particles: [0,1e-7) / [1e-7,1e-6) / [1e-6,1e-5) / [1e-5,1e-4) / "
          "[1e-4,1e-3) / >=1e-3.  north_star: mask must match outside 1e-6 (bins 3-6 must be 0 against f64acc); "
          "potentials within 1e-6.\n")
    print("| case | mode | N | passes GPU / oracle | max rel err Phi vs f64acc | mask mismatches vs f64acc (bins) | "
          "max rel err Phi vs f32seq | mask mismatches vs f32seq (bins) | bit-identical to f32seq |")
    print("|---|---|---|---|---|---|---|---|---|")
    c = synth.config1()
    s, g, d = c.stars, c.gas, c.dm
    M = s.mass.sum()
    vb = (np.sum(s.mass * s.vx) / M, np.sum(s.mass * s.vy) / M, np.sum(s.mass * s.vz) / M)
    cases = [
        ("cfg1 stars (1e4, sources gas+stars+DM)", (s.x, s.y, s.z, s.vx, s.vy, s.vz, s.mass),
         dict(pre=[g.pos_mass()], post=[d.pos_mass()], kappa=9.0), 9.0),
        ("cfg1 gas (1e4 lattice cells, classes gas|DM|stars)", (g.x, g.y, g.z, g.vx, g.vy, g.vz, g.mass),
         dict(post=[d.pos_mass(), s.pos_mass()], split_classes=True, kappa=2.0, vb_fixed=vb), 2.0),
    ]
    c2 = synth.config2(40_000, 100_000)
    s2, g2 = c2.stars, c2.gas
    M2 = s2.mass.sum()
    vb2 = (np.sum(s2.mass * s2.vx) / M2, np.sum(s2.mass * s2.vy) / M2, np.sum(s2.mass * s2.vz) / M2)
    cases += [
        ("cfg2 / 5 stars (4e4, NFW, sources 1e5 gas + stars)", (s2.x, s2.y, s2.z, s2.vx, s2.vy, s2.vz, s2.mass),
         dict(pre=[g2.pos_mass()], kappa=9.0), 9.0),
        ("cfg2 / 5 gas (1e5 two-level lattice)", (g2.x, g2.y, g2.z, g2.vx, g2.vy, g2.vz, g2.mass),
         dict(post=[s2.pos_mass()], split_classes=True, kappa=2.0, vb_fixed=vb2), 2.0),
    ]
    total_cpu = 0.0
    for name, members, kw, kappa in cases:
        rows, t = report(name, members, kw, kappa)
        total_cpu += t
        print("\n".join(rows), flush=True)
    print("\nOracle CPU time for these cases: %.0f s.  Larger sizes are covered by sampled and property tests "
          "(tests/test_gpu_unbind.py::test_cfg2_full_size_*, test_cfg3_full_*)." % total_cpu)


if __name__ == "__main__":
    main()
