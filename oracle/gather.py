"""CPU oracle for the gather step BEFORE the hot path -- TEST
INFRASTRUCTURE ONLY, like oracle.py: only tests/, smoke() and bench.py's cpu_baseline leg
may import it.

Restates, with numpy:
  patch_to_particles            python_scripts/halo_gas.py:9-52   (numba @njit, float64)
  AMRgrid_to_particles          python_scripts/halo_gas.py:56-141
  parallel_inside               python_scripts/halo_gas.py:216-218
  st_gas_dm_particles_inside    python_scripts/halo_gas.py:223-277

Third-party pieces that are NOT under /root/reference (un-vendored `masclet_framework`, no
version pinned anywhere in the reference; path configured at pyHALMA.dat:52-53):
  tools.create_vector_levels(npatch)      level of every patch, index 0 = base grid
  tools.which_patches_inside_box(...)     indices (ascending, 0 always included) of the patches
                                          whose extent overlaps the box
are restated from their published behaviour.  The gather's result does not depend on how
tight the overlap test is: patch_to_particles keeps only the cells strictly inside the box
(:32-38), so any ascending superset of the overlapping patches yields the same particles in
the same order.
  scipy.spatial.KDTree.query_ball_point([cx,cy,cz], R) (halo_gas.py:255,269): the set of
points with squared distance <= R^2; single-point queries return the indices in tree
order, unsorted.  The oracle (and the GPU path) return them ASCENDING; parity for DM and
stars is therefore on the index SET (the order only permutes float32 partial sums of the
potential kernel, which is what the FAST/f64acc parity mode is insensitive to).

PARITY STATUS: pinned by tests/golden/gather_amr.npz, written by tests/golden/make_golden.py
running the reference's own halo_gas.st_gas_dm_particles_inside (real numba
patch_to_particles, real scipy KDTree) on pyhalma_b200.synth.amr_snapshot.
"""
from __future__ import annotations

import numpy as np


def create_vector_levels(npatch):
    """masclet_framework.tools.create_vector_levels: [0] + [l] * npatch[l] for l >= 1."""
    out = [np.zeros(1, dtype=np.int64)]
    for lev in range(1, len(npatch)):
        out.append(np.full(int(npatch[lev]), lev, dtype=np.int64))
    return np.concatenate(out)


def which_patches_inside_box(box, patchnx, patchny, patchnz, patchrx, patchry, patchrz, npatch, L, ncoarse):
    """masclet_framework.tools.which_patches_inside_box: patches that (partially) lie inside
    `box` = (xmin, xmax, ymin, ymax, zmin, zmax).  A patch of level l spans
    [r - res_l, r - res_l + n * res_l] per axis (r = centre of its first parent cell)."""
    levels = create_vector_levels(npatch)
    res = (L / ncoarse) / 2.0 ** levels
    keep = [0]
    for p in range(1, len(levels)):
        lo = np.array([patchrx[p], patchry[p], patchrz[p]]) - res[p]
        hi = lo + np.array([patchnx[p], patchny[p], patchnz[p]]) * res[p]
        if (lo[0] <= box[1] and hi[0] >= box[0] and lo[1] <= box[3] and hi[1] >= box[2]
                and lo[2] <= box[5] and hi[2] >= box[4]):
            keep.append(p)
    return keep


def patch_to_particles(patch_res, patch_rx, patch_ry, patch_rz, patch_nx, patch_ny, patch_nz, patch_delta,
                       patch_cr0amr, patch_solapst, patch_vx, patch_vy, patch_vz, patch_temp, box, rho_B):
    """halo_gas.py:9-52.  One particle at the centre of every cell that is strictly inside the
    box (:32,35,38), not refined and not overlapped (:40), in ix-major / iz-minor order; all
    float64 (numba: int + float32 -> float64, :48)."""
    x = (patch_rx - patch_res / 2) + np.arange(patch_nx) * patch_res           # :27,31
    y = (patch_ry - patch_res / 2) + np.arange(patch_ny) * patch_res
    z = (patch_rz - patch_res / 2) + np.arange(patch_nz) * patch_res
    sx = (x > box[0]) & (x < box[1])
    sy = (y > box[2]) & (y < box[3])
    sz = (z > box[4]) & (z < box[5])
    keep = (sx[:, None, None] & sy[None, :, None] & sz[None, None, :]
            & np.asarray(patch_cr0amr).astype(bool) & np.asarray(patch_solapst).astype(bool))
    ix, iy, iz = np.nonzero(keep)                        # C order = the reference's loop nest
    f64 = np.float64
    # :48.  numba lowers `patch_res**3` to patch_res*patch_res*patch_res (checked against numba
    # 0.65 on random values: never pow()), and evaluates the product left to right.
    res3 = patch_res * patch_res * patch_res
    mass = (1 + np.asarray(patch_delta)[ix, iy, iz].astype(f64)) * rho_B * res3
    return (x[ix], y[iy], z[iz], np.asarray(patch_vx)[ix, iy, iz].astype(f64),
            np.asarray(patch_vy)[ix, iy, iz].astype(f64), np.asarray(patch_vz)[ix, iy, iz].astype(f64), mass,
            np.asarray(patch_temp)[ix, iy, iz].astype(f64))


def AMRgrid_to_particles(L, ncoarse, grid_data, gas_data, Rrps, cx, cy, cz, rho_B):  # noqa: N802
    """halo_gas.py:56-141: cells of all patches of level >= 1 inside the box of half-width
    Rrps around (cx, cy, cz), patch by patch in ascending patch order; velocities x 3e5
    (c = 1 -> km/s, :136-138).  Returns x, y, z, vx, vy, vz, mass, temp."""
    npatch = grid_data[5]
    patchnx, patchny, patchnz = grid_data[6], grid_data[7], grid_data[8]
    patchrx, patchry, patchrz = grid_data[12], grid_data[13], grid_data[14]
    box = np.array([cx - Rrps, cx + Rrps, cy - Rrps, cy + Rrps, cz - Rrps, cz + Rrps])        # :87-88
    which = which_patches_inside_box(box, patchnx, patchny, patchnz, patchrx, patchry, patchrz, npatch, L, ncoarse)
    level = create_vector_levels(npatch)
    cols = [[np.zeros(0)] for _ in range(8)]
    for p in which:
        lev = level[p]
        if lev >= 1:                                                                            # :107
            res = (L / ncoarse) / 2 ** int(lev)                                                 # :108
            out = patch_to_particles(res, patchrx[p], patchry[p], patchrz[p], patchnx[p], patchny[p], patchnz[p],
                                     gas_data[0][p], gas_data[5][p], gas_data[6][p], gas_data[1][p],
                                     gas_data[2][p], gas_data[3][p], gas_data[4][p], box, rho_B)
            for c, o in zip(cols, out):
                c.append(o)
    x, y, z, vx, vy, vz, m, t = (np.concatenate(c) for c in cols)
    return x, y, z, vx * 3e5, vy * 3e5, vz * 3e5, m, t


def parallel_inside(array_x, array_y, array_z, R, cx, cy, cz):
    """halo_gas.py:216-218 (strict <)."""
    return np.sqrt((array_x - cx) ** 2 + (array_y - cy) ** 2 + (array_z - cz) ** 2) < R


def ball_indices(x, y, z, cx, cy, cz, R):
    """KDTree.query_ball_point([cx,cy,cz], R): squared distance <= R^2, ascending here."""
    d2 = (np.asarray(x) - cx) ** 2 + (np.asarray(y) - cy) ** 2 + (np.asarray(z) - cz) ** 2
    return np.nonzero(d2 <= R * R)[0]


def st_gas_dm_particles_inside(rete, L, ncoarse, grid_data, gas_data, masclet_dm_data, masclet_st_data,
                               st_kdtree, dm_kdtree, cx, cy, cz, R, rho_B, mass_to_sun=1.0):
    """halo_gas.py:223-277.  The KD-trees are accepted for signature parity and not used."""
    gx, gy, gz, gvx, gvy, gvz, gm, gt = AMRgrid_to_particles(L, ncoarse, grid_data, gas_data, R, cx, cy, cz, rho_B)
    inside = parallel_inside(gx, gy, gz, R, cx, cy, cz)                                         # :236
    gx, gy, gz, gvx, gvy, gvz, gm, gt = (a[inside] for a in (gx, gy, gz, gvx, gvy, gvz, gm, gt))
    gm = gm * rete ** 3                                                                         # :246
    k = ball_indices(masclet_dm_data[0], masclet_dm_data[1], masclet_dm_data[2], cx, cy, cz, R)
    dm = [np.asarray(masclet_dm_data[0])[k], np.asarray(masclet_dm_data[1])[k], np.asarray(masclet_dm_data[2])[k],
          (np.asarray(masclet_dm_data[3]) * mass_to_sun)[k]]                                    # :249-259
    k = ball_indices(masclet_st_data[0], masclet_st_data[1], masclet_st_data[2], cx, cy, cz, R)
    st = [np.asarray(masclet_st_data[0])[k], np.asarray(masclet_st_data[1])[k], np.asarray(masclet_st_data[2])[k],
          (np.asarray(masclet_st_data[6]) * mass_to_sun)[k], np.asarray(masclet_st_data[9])[k]]  # :262-274
    return (gx, gy, gz, gvx, gvy, gvz, gm, gt, *dm, *st)


GATHER_NAMES = ("gas_x", "gas_y", "gas_z", "gas_vx", "gas_vy", "gas_vz", "gas_mass", "gas_temp", "dm_x", "dm_y",
                "dm_z", "dm_mass", "st_x", "st_y", "st_z", "st_mass", "st_oripa")


def canonical_gather(out):
    """The 17-tuple of st_gas_dm_particles_inside with DM ordered by (x, y, z) and stars by id:
    the order-free form in which the reference (KD-tree order) and the oracle / GPU path
    (ascending index) are compared.  Gas keeps the reference's order."""
    out = [np.asarray(a) for a in out]
    kd = np.lexsort((out[10], out[9], out[8]))
    ks = np.argsort(out[16], kind="stable")
    return tuple(out[:8]) + tuple(a[kd] for a in out[8:12]) + tuple(a[ks] for a in out[12:17])


def digest(a) -> np.ndarray:
    """sha256 of the array's bytes (bit-exact comparison without storing the array)."""
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()
