"""Device-resident snapshot and the per-halo gather.

Host-side mirror of the reference's gather step, python_scripts/halo_gas.py:
  AMRgrid_to_particles          :56-141   (with patch_to_particles, :9-52)
  st_gas_dm_particles_inside    :223-277  (with parallel_inside, :216-218, and the KD-tree
                                           ball queries :255, :269)
The reference walks the AMR patches in Python for every halo.  Here the whole snapshot --
cell fields of every patch, DM and star particles -- is uploaded to HBM once (`Snapshot`)
and each halo is one `halma_snapshot_gather` call on the device.  Gas arrays come back
bit-identical to the reference's float64 arrays and in its order; DM and stars come back as
the same set in ascending index order (the reference's KD-tree order is unspecified).
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib


def create_vector_levels(npatch) -> np.ndarray:
    """Level of every patch, index 0 = the base grid (masclet_framework.tools.create_vector_levels,
    called at halo_gas.py:92): [0] + [l] * npatch[l] for l >= 1."""
    out = [np.zeros(1, dtype=np.int32)]
    for lev in range(1, len(npatch)):
        out.append(np.full(int(npatch[lev]), lev, dtype=np.int32))
    return np.concatenate(out)


def _c_f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)          # C order: ix slowest, like the loop nest at :30-38


def _c_u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a) != 0, dtype=np.uint8)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class Snapshot:
    """One simulation output resident on the GPU: AMR hierarchy + gas cell fields, and
    optionally the DM and star particles."""

    def __init__(self, L, ncoarse, grid_data, gas_data, masclet_dm_data=None, masclet_st_data=None, *,
                 mass_to_sun: float = 1.0, device: int = 0):
        npatch = grid_data[5]
        level = create_vector_levels(npatch)
        n_patch = len(level)
        nx, ny, nz = (np.ascontiguousarray(grid_data[k][:n_patch], dtype=np.int32) for k in (6, 7, 8))
        rx, ry, rz = (_f64(grid_data[k][:n_patch]) for k in (12, 13, 14))
        if not (len(nx) == len(ny) == len(nz) == len(rx) == len(ry) == len(rz) == n_patch):
            raise ValueError("grid_data arrays are shorter than 1 + sum(npatch)")
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.device = device
        self.n_patch = n_patch
        _lib.check(self._L.halma_snapshot_create(device, float(L), int(ncoarse), n_patch, level.ctypes.data,
                                                 nx.ctypes.data, ny.ctypes.data, nz.ctypes.data, rx.ctypes.data,
                                                 ry.ctypes.data, rz.ctypes.data, C.byref(self._h)))
        delta, vx, vy, vz, temp, cr0amr, solapst = (gas_data[k] for k in range(7))
        for p in range(1, n_patch):
            shape = (int(nx[p]), int(ny[p]), int(nz[p]))
            arrs = [_c_f32(a[p]) for a in (delta, vx, vy, vz, temp)] + [_c_u8(cr0amr[p]), _c_u8(solapst[p])]
            for a in arrs:
                if a.shape != shape:
                    raise ValueError("patch %d: field of shape %s, grid says %s" % (p, a.shape, shape))
            _lib.check(self._L.halma_snapshot_upload_patch(self._h, p, *[a.ctypes.data for a in arrs]))
        self.n_cells = int(self._L.halma_snapshot_cells(self._h))
        self.oripa_dtype = np.int64
        if masclet_dm_data is not None:
            self.upload_particles(0, masclet_dm_data[0], masclet_dm_data[1], masclet_dm_data[2],
                                  np.asarray(masclet_dm_data[3]) * mass_to_sun)                     # :252
        if masclet_st_data is not None:
            self.oripa_dtype = np.asarray(masclet_st_data[9]).dtype
            self.upload_particles(1, masclet_st_data[0], masclet_st_data[1], masclet_st_data[2],
                                  np.asarray(masclet_st_data[6]) * mass_to_sun, masclet_st_data[9])  # :265-266

    def upload_particles(self, kind: int, x, y, z, mass, ids=None) -> None:
        x, y, z, mass = map(_f64, (x, y, z, mass))
        if not (len(x) == len(y) == len(z) == len(mass)):
            raise ValueError("particle arrays differ in length")
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            if len(ids) != len(x):
                raise ValueError("ids differ in length from the positions")
            idp = ids.ctypes.data
        _lib.check(self._L.halma_snapshot_upload_particles(self._h, kind, len(x), x.ctypes.data, y.ctypes.data,
                                                           z.ctypes.data, mass.ctypes.data, idp))

    def _gather(self, cx, cy, cz, R, rho_B, rete, dm_heavy_min):
        counts = (C.c_int64 * 4)()
        _lib.check(self._L.halma_snapshot_gather(self._h, float(cx), float(cy), float(cz), float(R), float(rho_B),
                                                 float(rete) ** 3, float(dm_heavy_min), counts))
        return tuple(int(c) for c in counts)

    def gather_box(self, cx, cy, cz, R, rho_B, rete=1.0):
        """AMRgrid_to_particles (halo_gas.py:56-141): the 8 gas arrays of the cells strictly inside the box of
        half-width R (no sphere test), as host arrays in the reference's order."""
        counts = (C.c_int64 * 4)()
        _lib.check(self._L.halma_snapshot_gather_box(self._h, float(cx), float(cy), float(cz), float(R), float(rho_B),
                                                     float(rete) ** 3, counts))
        ng = int(counts[0])
        g = [np.empty(ng) for _ in range(8)]
        if ng:
            _lib.check(self._L.halma_snapshot_fetch(self._h, (C.c_void_p * 8)(*[a.ctypes.data for a in g]), None, None,
                                                    None, None))
        return tuple(g)

    def gather(self, cx, cy, cz, R, rho_B, rete=1.0, *, gas=True, dm=True, stars=True):
        """Returns the 17-tuple of st_gas_dm_particles_inside (halo_gas.py:276-277) as host arrays;
        classes that are switched off come back empty."""
        ng, nd, _, ns = self._gather(cx, cy, cz, R, rho_B, rete, -np.inf)
        g = [np.empty(ng if gas else 0) for _ in range(8)]
        d = [np.empty(nd if dm else 0) for _ in range(4)]
        s = [np.empty(ns if stars else 0) for _ in range(4)]
        sid = np.empty(ns if stars else 0, dtype=np.int64)

        def ptrs(arrs, on):
            return (C.c_void_p * len(arrs))(*[a.ctypes.data if on else None for a in arrs])

        _lib.check(self._L.halma_snapshot_fetch(self._h, ptrs(g, gas and ng), ptrs(d, dm and nd), None,
                                                ptrs(s, stars and ns), sid.ctypes.data if stars and ns else None))
        return (*g, *d, *s, sid.astype(self.oripa_dtype, copy=False))

    def gather_device(self, cx, cy, cz, R, rho_B, rete=1.0, *, dm_heavy_min=None) -> "DeviceGather":
        """The same selection left on the GPU.  dm_heavy_min splits the DM into the heavy species
        (mass >= dm_heavy_min, halo_gas.py:347) and the light one.  The handle is valid until the
        next gather on this snapshot."""
        counts = self._gather(cx, cy, cz, R, rho_B, rete, -np.inf if dm_heavy_min is None else dm_heavy_min)
        ptr = (C.c_void_p * 4)()
        sid = C.c_void_p()
        _lib.check(self._L.halma_snapshot_result_device(self._h, ptr, C.byref(sid), None))
        return DeviceGather(self, counts, [p or 0 for p in ptr], sid.value or 0, dm_heavy_min is not None)

    def fetch_star(self, k: int):
        """x, y, z, mass, id of gathered star k."""
        out = np.empty(4)
        sid = C.c_int64()
        _lib.check(self._L.halma_snapshot_fetch_star(self._h, int(k), out.ctypes.data, C.byref(sid)))
        return out[0], out[1], out[2], out[3], np.asarray(sid.value).astype(self.oripa_dtype)[()]

    def close(self) -> None:
        if self._h:
            self._L.halma_snapshot_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class DeviceGather:
    """Device addresses of one gather: column k of a group is at ptr + 8 * k * n.  Groups: gas
    (x, y, z, vx, vy, vz, mass, temp), dm (all DM, or the heavy species), dm_light, stars
    (x, y, z, mass)."""

    def __init__(self, snap, counts, ptrs, st_id_ptr, split):
        self.snap = snap
        self.n_gas, self.n_dm, self.n_dm_light, self.n_st = counts
        self._ptr = ptrs
        self.st_id_ptr = st_id_ptr
        self.split = split

    def col(self, group: int, k: int) -> int:
        n = (self.n_gas, self.n_dm, self.n_dm_light, self.n_st)[group]
        return self._ptr[group] + 8 * k * n if n else 0

    # (mass, x, y, z) address tuples in the argument order of upload_group
    def gas_source(self):
        return self.n_gas, (self.col(0, 6), self.col(0, 0), self.col(0, 1), self.col(0, 2))

    def dm_source(self, light=False):
        g = 2 if light else 1
        return (self.n_dm_light if light else self.n_dm), (self.col(g, 3), self.col(g, 0), self.col(g, 1), self.col(g, 2))

    def star_source(self):
        return self.n_st, (self.col(3, 3), self.col(3, 0), self.col(3, 1), self.col(3, 2))


# The reference passes the same snapshot objects for every halo of a snapshot (pyHALMA.py:930-1037);
# keep the device copy of the last one seen.  The strong references keep id() from being reused.
_cache: dict = {"gas": None, "full": None}


def snapshot_for(L, ncoarse, grid_data, gas_data, masclet_dm_data, masclet_st_data, mass_to_sun, device) -> Snapshot:
    """Device copy of the snapshot these objects describe.  Two slots: the grid + gas fields alone
    (AMRgrid_to_particles passes no particles) and grid + gas + particles; a request without particles is also
    served by the copy that has them, so alternating the two kinds of call per halo uploads nothing twice.
    A replaced snapshot is dropped, not closed: DeviceGather handles hold a reference, and the device memory goes
    when the last one does."""
    gkey = (id(grid_data), id(gas_data), float(L), int(ncoarse), int(device))
    want_particles = masclet_dm_data is not None or masclet_st_data is not None
    full = _cache["full"]
    if full is not None and full[0] == gkey and (not want_particles or full[1] == (
            id(masclet_dm_data), id(masclet_st_data), float(mass_to_sun))):
        return full[3]
    if not want_particles:
        gas = _cache["gas"]
        if gas is None or gas[0] != gkey:
            snap = Snapshot(L, ncoarse, grid_data, gas_data, None, None, device=device)
            _cache["gas"] = gas = (gkey, None, (grid_data, gas_data), snap)
        return gas[3]
    snap = Snapshot(L, ncoarse, grid_data, gas_data, masclet_dm_data, masclet_st_data, mass_to_sun=mass_to_sun,
                    device=device)
    _cache["full"] = (gkey, (id(masclet_dm_data), id(masclet_st_data), float(mass_to_sun)),
                      (grid_data, gas_data, masclet_dm_data, masclet_st_data), snap)
    if _cache["gas"] is not None and _cache["gas"][0] == gkey:
        _cache["gas"] = None          # the full copy serves those requests from now on
    return snap


def release_cached_snapshot() -> None:
    for k in ("gas", "full"):
        if _cache[k] is not None:
            _cache[k][3].close()
        _cache[k] = None


def default_mass_to_sun() -> float:
    """masclet_framework.units.mass_to_sun (halo_gas.py:252,265) when that un-vendored package is
    importable; callers without it pass mass_to_sun explicitly."""
    try:
        from masclet_framework import units      # type: ignore
    except Exception as exc:
        raise ImportError("masclet_framework is not importable: pass mass_to_sun=... (masses are multiplied by "
                          "units.mass_to_sun at halo_gas.py:252,265)") from exc
    return float(units.mass_to_sun)
