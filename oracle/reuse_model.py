"""Numpy model of the device loop's reuse between passes (TEST INFRASTRUCTURE, like the rest of oracle/).

The CUDA loop (pyhalma_b200/csrc, halma_unbind_config.cache_external / .incremental) does not
re-evaluate what cannot have changed between two passes of the fixed-point iteration of
SURVEY.md §3.4:

* the sum over the fixed external sources is evaluated by the first pass only;
* after a pass that removed at most a third of a halo's members (2 * n_removed <= n_left) the next
  pass evaluates survivors x removed members -- with the reference's predicate
  (fortran_modules/particle_subroutines.f90:499-501: a pair counts only if all three coordinates
  differ) -- and subtracts that from the complete potential it kept.

A full pass works on the PREDICATE-FREE sum (every pair with r > 0) and subtracts the pairs the
reference excludes from the current sources, like the correction tickets of the CUDA kernel do.

This model restates that bookkeeping with the FAST kernel's arithmetic granularity -- float32 terms,
float32 partial sums over 32 sources, float64 accumulation -- so that the CPU suite can pin (a) that
the member sets, pass counts and potentials stay inside the FAST tolerances of the plain loop
(oracle.unbind_halo, variant f64acc) and (b) the number of 1/r evaluations a one-sided run makes,
which the GPU tests compare with halma_run_stats.evaluations.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import oracle as O

CHUNK = 32      # sources per float32 partial sum (csrc/potential.cu: kFlushQuads * 4)
INCR_HEAVY = 0.8    # csrc/potential.cu::kIncrHeavy: an incremental pass in which the removed members carried more
                    # than this share of some member's kept potential is redone as a full predicated pass


def _terms(tx, ty, tz, sm, sx, sy, sz):
    """float32 m / r for every (target, source) pair and the three coordinate differences' zero tests."""
    dx = sx[None, :] - tx[:, None]
    dy = sy[None, :] - ty[:, None]
    dz = sz[None, :] - tz[:, None]
    r2 = (dz * dz + (dx * dx + dy * dy)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (sm[None, :] / np.sqrt(r2)).astype(np.float32)
    return t, r2, (dx == 0), (dy == 0), (dz == 0)


def _free_sum(tx, ty, tz, sm, sx, sy, sz, predicate: bool = False):
    """Predicate-free sum (pairs with r > 0) or, with predicate=True, the reference's sum (pairs whose three
    coordinates all differ); float32 partials over CHUNK sources, float64 total."""
    out = np.zeros(len(tx), np.float64)
    for a in range(0, len(sm), CHUNK):
        t, r2, ex, ey, ez = _terms(tx, ty, tz, sm[a:a + CHUNK], sx[a:a + CHUNK], sy[a:a + CHUNK], sz[a:a + CHUNK])
        t = np.where(~(ex | ey | ez) if predicate else (r2 > 0), t, np.float32(0))
        part = np.zeros(len(tx), np.float32)
        for k in range(t.shape[1]):
            part = part + t[:, k]
        out += part.astype(np.float64)
    return out


def _excluded_sum(tx, ty, tz, sm, sx, sy, sz):
    """Sum over the pairs the reference's predicate drops although r > 0 (a coordinate is shared)."""
    out = np.zeros(len(tx), np.float64)
    for a in range(0, len(sm), 4 * CHUNK):
        t, r2, ex, ey, ez = _terms(tx, ty, tz, sm[a:a + 4 * CHUNK], sx[a:a + 4 * CHUNK], sy[a:a + 4 * CHUNK],
                                   sz[a:a + 4 * CHUNK])
        drop = (ex | ey | ez) & (r2 > 0)
        out += np.where(drop, t, np.float32(0)).astype(np.float64).sum(axis=1)
    return out


@dataclass
class ModelResult:
    mask: np.ndarray
    idx: np.ndarray
    be32: np.ndarray
    energy: np.ndarray
    n_iter: int
    evaluations: int                 # 1/r evaluations of a one-sided run (symmetric = 0)
    passes: List[str] = field(default_factory=list)      # "full" / "incr" per pass


def unbind_halo(x, y, z, vx, vy, vz, mass, *, ext: Sequence = (), kappa: float = 9.0, vb_fixed=None,
                max_iter: int = 64, cache_external: bool = True, incremental: bool = True,
                heavy_guard: bool = True) -> ModelResult:
    """ext: sequence of (mass, x, y, z) fixed source groups (their order does not matter in FAST mode).
    heavy_guard=False switches the INCR_HEAVY safeguard off (to show what it is for)."""
    f32 = np.float32
    X, Y, Z, M = (np.asarray(a, np.float64).astype(f32) for a in (x, y, z, mass))
    vx, vy, vz, m64 = (np.asarray(a, np.float64) for a in (vx, vy, vz, mass))
    if len(ext):
        em, ex, ey, ez = (np.concatenate([np.asarray(g[k], np.float64) for g in ext]).astype(f32) for k in range(4))
    else:
        em = ex = ey = ez = np.zeros(0, f32)
    n_ext = len(em)
    N = len(X)
    idx = np.arange(N)
    phi_keep = np.zeros(N)          # the complete float64 potential of the previous pass
    phi_ext = np.zeros(N)
    be_out = np.zeros(N, f32)
    e_out = np.zeros(N)
    removed = np.zeros(0, np.int64)
    incr_next = False
    evals = 0
    it = 0
    kinds = []
    while len(idx) > 0 and it < max_iter:
        t = (X[idx], Y[idx], Z[idx])
        n = len(idx)
        fallback = False
        if incr_next:
            r = removed
            gone = _free_sum(*t, M[r], X[r], Y[r], Z[r], predicate=True)
            fallback = heavy_guard and bool(np.any(np.abs(gone) > INCR_HEAVY * np.abs(phi_keep[idx])))
            phi = phi_keep[idx] - gone
            if not fallback:
                evals += n * len(r)
                kinds.append("incr")
        if fallback:
            # the whole halo again, with the reference's predicate (the predicated kernel): complete in itself
            src = (np.concatenate((M[idx], em)), np.concatenate((X[idx], ex)), np.concatenate((Y[idx], ey)),
                   np.concatenate((Z[idx], ez)))
            phi = _free_sum(*t, *src, predicate=True)
            evals += n * (n + n_ext)
            kinds.append("fallback")
        elif not incr_next:
            phi = _free_sum(*t, M[idx], X[idx], Y[idx], Z[idx])
            evals += n * n
            cached = cache_external and it > 0
            if n_ext and not cached:
                e = _free_sum(*t, em, ex, ey, ez)
                evals += n * n_ext
                if cache_external:
                    phi_ext[idx] = e
                else:
                    phi = phi + e
            if cache_external:
                phi = phi + phi_ext[idx]
            src = (np.concatenate((M[idx], em)), np.concatenate((X[idx], ex)), np.concatenate((Y[idx], ey)),
                   np.concatenate((Z[idx], ez)))
            phi = phi - _excluded_sum(*t, *src)
            kinds.append("full")
        phi_keep[idx] = phi
        be = phi.astype(f32)
        if vb_fixed is None:
            Mtot = O.total_mass(idx, m64)
            vb = O.CM_velocity(Mtot, idx, vx, vy, vz, m64)
        else:
            vb = tuple(float(v) for v in vb_fixed)
        E = O.energy_step(be, vx[idx], vy[idx], vz[idx], vb[0], vb[1], vb[2], kappa)
        bound = E <= 0.
        be_out[idx] = be
        e_out[idx] = E
        it += 1
        new_idx = idx[bound]
        removed = idx[~bound]
        changed = len(new_idx) != len(idx)
        incr_next = incremental and not fallback and len(removed) > 0 and 2 * len(removed) <= len(new_idx)
        idx = new_idx
        if not changed:
            break
    mask = np.zeros(N, bool)
    mask[idx] = True
    return ModelResult(mask, idx, be_out, e_out, it, evals, kinds)


def expected_evaluations(n_history: Sequence[int], n_ext: int, *, cache_external: bool, incremental: bool,
                         symmetric: bool = False, tile: int = 128) -> int:
    """1/r evaluations of the device loop for a halo whose member count went n_history[0] -> [1] -> ...
    (oracle.UnbindResult.n_bound_history; one pass per entry but the last), with no pass falling back to
    the predicated kernel.  Mirrors csrc/loop_kernels.cu::k_halo_decide: a full pass evaluates the members
    against themselves (every pair of different 128-member tiles once with the symmetric self-term) and,
    unless cached, against the externals; an incremental pass evaluates survivors x removed."""
    total = 0
    incr_next = False
    for k in range(len(n_history) - 1):
        n = int(n_history[k])
        if n == 0:
            break
        if incr_next:
            total += n * (int(n_history[k - 1]) - n)
        else:
            tiles = (n + tile - 1) // tile
            if symmetric and tiles >= 2:
                last = n - (tiles - 1) * tile
                total += (n * n + (tiles - 1) * tile * tile + last * last) // 2
            else:
                total += n * n
            if not (cache_external and k > 0):
                total += n * n_ext
        rem = n - int(n_history[k + 1])
        incr_next = incremental and rem > 0 and 2 * rem <= int(n_history[k + 1])
    return total
