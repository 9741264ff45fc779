"""Device-resident iterative unbinding: Python face of the halma_plan_* C-ABI.

Semantics: each pass evaluates the reference's one-pass unbinding
(python_scripts/halo_properties.py:333-361 for stars, python_scripts/halo_gas.py:299-476
for gas) on the current member set, keeps the bound members in their original order, and
repeats until the set stops changing.  `max_iter=1` with a fixed bulk velocity IS the
reference function.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

G_CONST = (4.3 * 1e-3) * 1e-6          # halo_gas.py:459-460, halo_properties.py:345-346


def _f64(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("expected a rank-1 array")
    return a


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


@dataclass
class RunStats:
    total_ms: float
    potential_ms: float
    potential_launches: int
    launches: int
    passes: int
    pairs: int                        # (target, source) interactions: what the reference's loop visits
    evaluations: int = 0              # 1/r evaluations made for them (fewer than pairs in symmetric mode)
    driver: int = 0                   # _lib.DRIVER_* that ran (fused: the whole loop was one kernel launch)
    loop_ms: float = 0.0              # fused driver: duration of the persistent kernel by its own clock
    comm_ms: float = 0.0              # split mode: CUDA-event time of the NCCL collectives
    comm_bytes: int = 0               # split mode: bytes handed to the collectives
    phase_ms: tuple = ()              # fused driver: prologue, potential, energy + compaction, tables, epilogue


@dataclass
class HaloResult:
    n_bound: int
    n_iter: int
    converged: bool
    mass: float
    com: tuple
    vb: tuple
    pairs: int
    most_bound: int = -1              # local index of the member with the deepest potential (last pass)
    mass_initial: float = 0.0
    cold_bound_mass: float = 0.0      # needs upload_temp
    unbound_cold_mass: float = 0.0
    unbound_hot_mass: float = 0.0


class HaloResults:
    """Sequence of HaloResult over the array halma_plan_download filled; the Python objects are
    made on access (a catalogue has 1e4 haloes and most callers look at a few fields of a few)."""

    def __init__(self, raw, n):
        self._raw = raw
        self._n = n

    def __len__(self):
        return self._n

    def _make(self, h):
        r = self._raw[h]
        return HaloResult(int(r.n_bound), int(r.n_iter), bool(r.converged), float(r.mass), tuple(r.com), tuple(r.vb),
                          int(r.pairs), int(r.most_bound), float(r.mass_initial), float(r.cold_bound_mass),
                          float(r.unbound_cold_mass), float(r.unbound_hot_mass))

    def __getitem__(self, h):
        if isinstance(h, slice):
            return [self._make(k) for k in range(*h.indices(self._n))]
        if h < 0:
            h += self._n
        if not 0 <= h < self._n:
            raise IndexError(h)
        return self._make(h)

    def __iter__(self):
        return (self._make(h) for h in range(self._n))

    def field(self, name) -> np.ndarray:
        """One field of every halo as an array (n_bound, n_iter, converged, mass, pairs, most_bound,
        mass_initial, cold_bound_mass, unbound_cold_mass, unbound_hot_mass: shape (n,); com, vb: (n, 3))."""
        if self._n == 0:
            return np.zeros((0, 3) if name in ("com", "vb") else 0)
        a = np.frombuffer(self._raw, dtype=_HALO_DTYPE, count=self._n)
        return a[name].copy()


_HALO_DTYPE = np.dtype([("n_bound", np.int64), ("n_iter", np.int32), ("converged", np.int32), ("mass", np.float64),
                        ("com", np.float64, 3), ("vb", np.float64, 3), ("pairs", np.int64), ("most_bound", np.int64),
                        ("mass_initial", np.float64), ("cold_bound_mass", np.float64),
                        ("unbound_cold_mass", np.float64), ("unbound_hot_mass", np.float64)])


@dataclass
class CatalogueResult:
    offsets: np.ndarray
    mask: Optional[np.ndarray]
    be32: Optional[np.ndarray]
    energy: Optional[np.ndarray]
    idx_packed: Optional[np.ndarray]
    halos: Sequence[HaloResult]
    stats: Optional[RunStats] = None

    def members(self, h: int) -> np.ndarray:
        """Ascending local indices of the bound members of halo h."""
        a = int(self.offsets[h])
        return self.idx_packed[a:a + self.halos[h].n_bound]

    def halo_mask(self, h: int) -> np.ndarray:
        return self.mask[int(self.offsets[h]):int(self.offsets[h + 1])].astype(bool)


@dataclass
class UnbindResult:
    """Single-halo result, field-compatible with oracle.UnbindResult."""
    mask: np.ndarray
    idx: np.ndarray
    be32: np.ndarray
    energy: np.ndarray
    n_iter: int
    mass: float
    com: tuple
    vb: tuple
    pairs: int
    converged: bool
    stats: Optional[RunStats] = field(default=None, repr=False)


REUSE_DEFAULT = "1"          # on unless HALMA_CACHE_EXT=0 / HALMA_INCREMENTAL=0


def _reuse_default(name: str) -> bool:
    import os
    return os.environ.get(name, REUSE_DEFAULT) not in ("0", "")


class UnbindPlan:
    """A batch of haloes resident on one GPU.

    offsets: int64[n_halo+1] CSR over the concatenated member arrays.
    ext_offsets: one int64[n_halo+1] per external source group.
    Stellar layout (halo_properties.py:333-339): split_classes=False, groups[:n_pre] are
    summed before the members (gas), the rest after (DM).  Gas layout (halo_gas.py:301-450):
    split_classes=True, members (gas) first, then each group as its own float32 class.

    cache_external / incremental (FAST mode, plans on the predicate-free path, one GPU): do not repeat
    work whose result cannot have changed -- the sum over the fixed external sources is evaluated by the
    first pass only, and after a pass that removed at most a third of a halo's members the next pass
    evaluates survivors x removed and subtracts (see include/halma_unbind.h).  None: HALMA_CACHE_EXT /
    HALMA_INCREMENTAL.
    """

    def __init__(self, offsets, ext_offsets: Sequence = (), *, mode=None, n_pre: int = 0,
                 split_classes: bool = False, vb_fixed: bool = False, max_iter: int = 64,
                 G: float = G_CONST, kappa: float = 9.0, device: int = 0, rank: int = 0, n_ranks: int = 1,
                 use_graph: Optional[bool] = None, symmetric: Optional[bool] = None,
                 cache_external: Optional[bool] = None, incremental: Optional[bool] = None,
                 driver: Optional[str] = None):
        L = _lib.lib()
        self._L = L
        self.offsets = _i64(offsets)
        if self.offsets.ndim != 1 or len(self.offsets) < 1:
            raise ValueError("offsets must be int64[n_halo+1]")
        self.n_halo = len(self.offsets) - 1
        self.n = int(self.offsets[-1])
        self.ext_offsets = [_i64(e) for e in ext_offsets]
        for e in self.ext_offsets:
            if e.shape != self.offsets.shape:
                raise ValueError("every ext_offsets array must have n_halo+1 entries")
        if len(self.ext_offsets) > _lib.MAX_GROUPS:
            raise ValueError("at most %d external groups" % _lib.MAX_GROUPS)
        cfg = _lib.UnbindConfig()
        cfg.struct_size = C.sizeof(_lib.UnbindConfig)
        cfg.device = device
        cfg.mode = _lib.default_mode() if mode is None else _lib.mode_code(mode)      # None: HALMA_MODE
        cfg.n_groups = len(self.ext_offsets)
        cfg.n_pre = n_pre
        cfg.split_classes = int(bool(split_classes))
        cfg.vb_fixed = int(bool(vb_fixed))
        cfg.max_iter = int(max_iter)
        cfg.G = float(G)
        cfg.kappa = float(kappa)
        cfg.rank = rank
        cfg.n_ranks = n_ranks
        # loop driver (halma_unbind_config.use_graph): "auto" = the persistent loop kernel on one GPU
        # (one launch for the whole unbinding), "enqueue" = stand-alone kernels queued ahead by the host
        # (per-launch events; always in split mode), "graph" = CUDA-graph WHILE node, "fused" = auto or fail
        if driver is None:
            import os
            driver = "graph" if use_graph or (use_graph is None and os.environ.get("HALMA_GRAPH", "0") not in ("0", "")) \
                else os.environ.get("HALMA_DRIVER", "auto")
        cfg.use_graph = {"auto": _lib.DRIVER_AUTO, "graph": _lib.DRIVER_GRAPH, "enqueue": _lib.DRIVER_ENQUEUE,
                         "fused": _lib.DRIVER_FUSED}[str(driver).lower()]
        if symmetric is None:
            import os
            symmetric = os.environ.get("HALMA_SYMMETRIC", "1") not in ("0", "")      # on unless HALMA_SYMMETRIC=0
        cfg.symmetric = int(bool(symmetric))
        # work the loop does not repeat (halma_unbind_config.cache_external / .incremental)
        cfg.cache_external = int(_reuse_default("HALMA_CACHE_EXT") if cache_external is None else bool(cache_external))
        cfg.incremental = int(_reuse_default("HALMA_INCREMENTAL") if incremental is None else bool(incremental))
        self.cfg = cfg
        ptrs = (C.POINTER(C.c_int64) * max(1, len(self.ext_offsets)))()
        for g, e in enumerate(self.ext_offsets):
            ptrs[g] = e.ctypes.data_as(C.POINTER(C.c_int64))
        h = C.c_void_p()
        _lib.check(L.halma_plan_create(C.byref(cfg), self.n_halo, self.offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                       ptrs, C.byref(h)))
        self._h = h
        self._keep = []      # host arrays referenced by in-flight copies

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.halma_plan_destroy(self._h)
            self._h = None
            self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- inputs ---------------------------------------------------------------------------
    def upload_members(self, x, y, z, vx, vy, vz, mass):
        arrs = [_f64(a) for a in (x, y, z, vx, vy, vz, mass)]
        for a in arrs:
            if len(a) != self.n:
                raise ValueError("member arrays must have offsets[-1] = %d entries, got %d" % (self.n, len(a)))
        self._keep.append(arrs)
        _lib.check(self._L.halma_plan_upload_members(self._h, *[a.ctypes.data for a in arrs]))

    def upload_group(self, group: int, mass, x, y, z):
        arrs = [_f64(a) for a in (mass, x, y, z)]
        n = int(self.ext_offsets[group][-1])
        for a in arrs:
            if len(a) != n:
                raise ValueError("group %d arrays must have %d entries, got %d" % (group, n, len(a)))
        self._keep.append(arrs)
        _lib.check(self._L.halma_plan_upload_group(self._h, group, *[a.ctypes.data for a in arrs]))

    # Raw addresses (host or device, e.g. gather.DeviceGather columns): no conversion, no copy on the
    # host.  The caller keeps the memory alive and unchanged until run() returns.
    def upload_members_raw(self, x, y, z, vx, vy, vz, mass):
        _lib.check(self._L.halma_plan_upload_members(self._h, *[C.c_void_p(int(a)) for a in (x, y, z, vx, vy, vz, mass)]))

    def upload_group_raw(self, group: int, mass, x, y, z):
        _lib.check(self._L.halma_plan_upload_group(self._h, group, *[C.c_void_p(int(a)) for a in (mass, x, y, z)]))

    def upload_temp_raw(self, temp, cold_T: float = 5 * 1e4):
        _lib.check(self._L.halma_plan_upload_temp(self._h, C.c_void_p(int(temp)), float(cold_T)))

    def upload_temp(self, temp, cold_T: float = 5 * 1e4):
        """Member temperatures for the cold / hot mass sums of RPS (halo_gas.py:479-492)."""
        t = _f64(temp)
        if len(t) != self.n:
            raise ValueError("temp must have offsets[-1] entries")
        self._keep.append([t])
        _lib.check(self._L.halma_plan_upload_temp(self._h, t.ctypes.data, float(cold_T)))

    def set_vb(self, vb):
        vb = np.ascontiguousarray(vb, dtype=np.float64).reshape(-1)
        if len(vb) != 3 * self.n_halo:
            raise ValueError("vb must have 3*n_halo entries")
        self._keep.append([vb])
        _lib.check(self._L.halma_plan_set_vb(self._h, vb.ctypes.data))

    def sync(self):
        """Wait for the asynchronous uploads (and anything else) enqueued on the plan's stream."""
        _lib.check(self._L.halma_plan_sync(self._h))

    def join(self, unique_id: bytes):
        _lib.ensure_nccl_path()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _lib.check(self._L.halma_plan_join(self._h, buf))

    def use_comm(self, comm: "Communicator"):
        """Split mode with a communicator that outlives the plan."""
        self._comm = comm
        _lib.check(self._L.halma_plan_use_comm(self._h, comm._h))

    # -- execution ------------------------------------------------------------------------
    def run(self) -> RunStats:
        st = _lib.RunStats()
        _lib.check(self._L.halma_plan_run(self._h, C.byref(st)))
        self._keep = []
        return RunStats(st.total_ms, st.potential_ms, st.potential_launches, st.launches, st.passes, st.pairs,
                        st.evaluations, st.driver, st.loop_ms, st.comm_ms, st.comm_bytes, tuple(st.phase_ms))

    def debug_pass_us(self):
        """Tuning aid: (potential, energy + compaction, tables) in microseconds for the first 16 passes of the
        last run with the persistent loop kernel."""
        out = (C.c_uint32 * 48)()
        _lib.check(self._L.halma_plan_debug_pass_ns(self._h, out))
        return [tuple(out[3 * k + j] / 1e3 for j in range(3)) for k in range(16)]

    def download_into(self, mask_addr=0, be_addr=0, energy_addr=0, idx_addr=0, halos_addr=0) -> None:
        """halma_plan_download into caller-owned host memory (raw addresses; 0 = skip): mask uint8[n],
        be float32[n], energy float64[n], idx int32[n], halos halma_halo_result[n_halo]."""
        def vp(a):
            return C.c_void_p(int(a)) if a else None
        _lib.check(self._L.halma_plan_download(self._h, vp(mask_addr), vp(be_addr), vp(energy_addr), vp(idx_addr),
                                               C.cast(vp(halos_addr), C.POINTER(_lib.HaloResult)) if halos_addr else None))

    def download(self, mask=True, be=True, energy=True, idx=True, halos=True) -> CatalogueResult:
        n = self.n
        # result arrays live in pooled page-locked memory: D2H at link speed, no page faults
        m = _lib.pinned_empty(n, np.uint8) if mask else None
        b = _lib.pinned_empty(n, np.float32) if be else None
        e = _lib.pinned_empty(n, np.float64) if energy else None
        i = _lib.pinned_empty(n, np.int32) if idx else None
        hr = (_lib.HaloResult * max(1, self.n_halo))() if halos else None
        _lib.check(self._L.halma_plan_download(
            self._h, m.ctypes.data if mask else None, b.ctypes.data if be else None,
            e.ctypes.data if energy else None, i.ctypes.data if idx else None, hr))
        out = HaloResults(hr, self.n_halo) if halos else []
        return CatalogueResult(self.offsets, m, b, e, i, out)


class Communicator:
    """NCCL communicator shared by the split-mode plans of one process (creating one takes
    ~0.1 s).  unique_id: the 128 bytes made by rank 0 with nccl_unique_id() and broadcast by
    the caller (e.g. torch.distributed.broadcast_object_list)."""

    def __init__(self, unique_id: bytes, rank: int, n_ranks: int, device: int = 0):
        _lib.ensure_nccl_path()
        self._L = _lib.lib()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        h = C.c_void_p()
        _lib.check(self._L.halma_comm_create(device, rank, n_ranks, buf, C.byref(h)))
        self._h = h
        self.rank, self.n_ranks, self.device = rank, n_ranks, device

    def close(self):
        if getattr(self, "_h", None):
            self._L.halma_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    _lib.ensure_nccl_path()
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().halma_nccl_unique_id(buf))
    return buf.raw


def unbind_catalogue(offsets, x, y, z, vx, vy, vz, mass, *, groups: Sequence = (), n_pre: int = 0,
                     split_classes: bool = False, vb=None, kappa: float = 9.0, max_iter: int = 64,
                     mode=None, device: int = 0, G: float = G_CONST, temp=None,
                     cold_T: float = 5 * 1e4, symmetric: Optional[bool] = None, streams: int = 1,
                     cache_external: Optional[bool] = None, incremental: Optional[bool] = None) -> CatalogueResult:
    """Unbind every halo of a catalogue in one batched, device-resident run.

    groups: sequence of (ext_offsets, mass, x, y, z) external source groups.
    vb: None (bulk velocity recomputed from the bound set each pass) or float64[n_halo, 3].
    temp: optional member temperatures; fills the cold / hot mass sums of every HaloResult.
    symmetric: evaluate member x member pairs once for both particles (halma_unbind_config.symmetric).
    streams: > 1 cuts the catalogue into that many runs of consecutive haloes of about equal cost,
        each with its own plan and host thread, so that the upload, sort and download of one part
        overlap the device loop of another (haloes are independent; in FAST mode the j-split of a
        halo can differ from the one-plan run, i.e. potentials agree to ~1e-16 before their float32
        rounding rather than always bit for bit; EXACT mode is bit-identical).
        0: chosen from the size of the upload (auto_parts).
    """
    offsets = _i64(offsets)
    if streams == 0:
        streams = auto_parts(int(offsets[-1]) if len(offsets) else 0, sum(len(g[1]) for g in groups))
    if streams > 1 and len(offsets) - 1 >= 2 * streams:
        return _unbind_catalogue_parts(offsets, x, y, z, vx, vy, vz, mass, groups, n_pre, split_classes, vb, kappa,
                                       max_iter, mode, device, G, temp, cold_T, symmetric, streams,
                                       cache_external, incremental)
    plan = UnbindPlan(offsets, [g[0] for g in groups], mode=mode, n_pre=n_pre, split_classes=split_classes,
                      vb_fixed=vb is not None, max_iter=max_iter, G=G, kappa=kappa, device=device,
                      symmetric=symmetric, cache_external=cache_external, incremental=incremental)
    try:
        plan.upload_members(x, y, z, vx, vy, vz, mass)
        for k, g in enumerate(groups):
            plan.upload_group(k, g[1], g[2], g[3], g[4])
        if vb is not None:
            plan.set_vb(vb)
        if temp is not None:
            plan.upload_temp(temp, cold_T)
        stats = plan.run()
        res = plan.download()
        res.stats = stats
        return res
    finally:
        plan.close()


def auto_parts(n_members: int, n_external: int = 0) -> int:
    """Parts of the one-shot catalogue call.  A part more hides the upload of the parts behind it under the device loop
    of the one before and costs ~1.2 ms of its own (plan, one-off sort launches, a loop launch, a download), so it pays
    from ~3 ms of upload per part: 3 parts for the 9e6-member cfg3 catalogue on one GPU (507 MB, ~10 ms over PCIe:
    47.8 ms against 55.8 ms in one plan), 1 for the eighth of it one GPU of eight gets (7.8 ms against 10.2 ms in
    3 parts; scripts/e2e_share.py)."""
    upload_ms = (56.0 * n_members + 32.0 * n_external) / 50e6          # float64 rows over ~50 GB/s
    return max(1, min(3, 1 + int(upload_ms / 3.0)))


def cost_cuts(offsets, ext_offsets, parts: int, taper: bool = False) -> list:
    """Halo indices that cut a catalogue into at most `parts` runs of consecutive haloes of about
    equal cost N_h (N_h + N_ext,h): [0, ..., n_halo], strictly increasing.  taper: the first and the last
    run get half the cost of the others (the pipelined one-shot call cannot hide the first run's upload
    nor the last run's download, so they are kept short)."""
    offsets = _i64(offsets)
    nh = len(offsets) - 1
    nmem = np.diff(offsets).astype(np.float64)
    next_ = sum((np.diff(_i64(e)).astype(np.float64) for e in ext_offsets), np.zeros(nh))
    cum = np.concatenate(([0.0], np.cumsum(nmem * (nmem + next_))))
    w = np.ones(parts)
    if taper and parts >= 3:
        w[0] = w[-1] = 0.5
    frac = np.cumsum(w) / w.sum()
    cuts = [0] + [int(np.searchsorted(cum, cum[-1] * f)) for f in frac[:-1]] + [nh]
    return sorted(set(min(max(c, 0), nh) for c in cuts))


def _unbind_catalogue_parts(offsets, x, y, z, vx, vy, vz, mass, groups, n_pre, split_classes, vb, kappa, max_iter,
                            mode, device, G, temp, cold_T, symmetric, streams, cache_external=None,
                            incremental=None) -> CatalogueResult:
    import threading
    nh = len(offsets) - 1
    n = int(offsets[-1])
    members = [_f64(a) for a in (x, y, z, vx, vy, vz, mass)]
    groups = [(_i64(g[0]),) + tuple(_f64(a) for a in g[1:]) for g in groups]
    vb = None if vb is None else np.ascontiguousarray(vb, dtype=np.float64).reshape(nh, 3)
    temp = None if temp is None else _f64(temp)
    cuts = cost_cuts(offsets, [g[0] for g in groups], streams, taper=True)
    out_mask = _lib.pinned_empty(n, np.uint8)
    out_be = _lib.pinned_empty(n, np.float32)
    out_energy = _lib.pinned_empty(n, np.float64)
    out_idx = _lib.pinned_empty(n, np.int32)
    raw = (_lib.HaloResult * max(1, nh))()
    stats: list = [None] * (len(cuts) - 1)
    errors: list = []

    # Uploads go over the host link one part at a time, in order: started together they would share it, and the
    # first part's loop could not begin until a third of ALL uploads had landed.  A part's upload then overlaps
    # the device loop of the part before it, its download the loop of the part after it.
    turn = threading.Condition()
    state = {"next": 0}

    def part(k):
        try:
            a, b = cuts[k], cuts[k + 1]
            o0, o1 = int(offsets[a]), int(offsets[b])
            sub = [(g[0][a:b + 1] - g[0][a],) + tuple(arr[int(g[0][a]):int(g[0][b])] for arr in g[1:]) for g in groups]
            with UnbindPlan(offsets[a:b + 1] - o0, [g[0] for g in sub], mode=mode, n_pre=n_pre,
                            split_classes=split_classes, vb_fixed=vb is not None, max_iter=max_iter, G=G, kappa=kappa,
                            device=device, symmetric=symmetric, cache_external=cache_external,
                            incremental=incremental) as plan:
                with turn:
                    turn.wait_for(lambda: state["next"] == k or errors)
                try:
                    plan.upload_members(*[m[o0:o1] for m in members])
                    for gi, g in enumerate(sub):
                        plan.upload_group(gi, g[1], g[2], g[3], g[4])
                    if vb is not None:
                        plan.set_vb(vb[a:b])
                    if temp is not None:
                        plan.upload_temp(temp[o0:o1], cold_T)
                    plan.sync()
                finally:
                    with turn:
                        state["next"] = k + 1
                        turn.notify_all()
                stats[k] = plan.run()
                plan.download_into(out_mask.ctypes.data + o0, out_be.ctypes.data + 4 * o0,
                                   out_energy.ctypes.data + 8 * o0, out_idx.ctypes.data + 4 * o0,
                                   C.addressof(raw) + a * C.sizeof(_lib.HaloResult))
        except Exception as exc:          # re-raised in the caller's thread
            errors.append(exc)
            with turn:
                state["next"] = max(state["next"], k + 1)
                turn.notify_all()

    threads = [threading.Thread(target=part, args=(k,)) for k in range(len(cuts) - 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    res = CatalogueResult(offsets, out_mask, out_be, out_energy, out_idx, HaloResults(raw, nh))
    res.stats = RunStats(max(s.total_ms for s in stats), sum(s.potential_ms for s in stats),
                         sum(s.potential_launches for s in stats), sum(s.launches for s in stats),
                         max(s.passes for s in stats), sum(s.pairs for s in stats), sum(s.evaluations for s in stats))
    return res


def unbind_halo(x, y, z, vx, vy, vz, mass, *, pre: Sequence = (), post: Sequence = (),
                split_classes: bool = False, kappa: float = 9.0, vb_fixed=None, max_iter: int = 64,
                mode=None, device: int = 0, G: float = G_CONST, symmetric: Optional[bool] = None,
                cache_external: Optional[bool] = None, incremental: Optional[bool] = None) -> UnbindResult:
    """One halo.  pre / post: sequences of (mass, x, y, z) fixed source groups summed before /
    after the members (same keywords as oracle.unbind_halo)."""
    n = len(x)
    ext = list(pre) + list(post)
    groups = [(np.array([0, len(g[0])], np.int64), g[0], g[1], g[2], g[3]) for g in ext]
    res = unbind_catalogue(np.array([0, n], np.int64), x, y, z, vx, vy, vz, mass, groups=groups,
                           n_pre=len(pre), split_classes=split_classes,
                           vb=None if vb_fixed is None else np.asarray(vb_fixed, np.float64).reshape(1, 3),
                           kappa=kappa, max_iter=max_iter, mode=mode, device=device, G=G, symmetric=symmetric,
                           cache_external=cache_external, incremental=incremental)
    h = res.halos[0]
    return UnbindResult(res.mask.astype(bool), res.members(0).astype(np.int64), res.be32, res.energy, h.n_iter,
                        h.mass, h.com, h.vb, h.pairs, h.converged, res.stats)
