"""ctypes binding of libhalma_unbind.so (include/halma_unbind.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable,
calls raise.  Build the library with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C pyhalma_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhalma_unbind.so")

HALMA_OK = 0
ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_ALIGN, ERR_TOO_LARGE, ERR_STATE, ERR_NCCL = -1, -2, -3, -4, -5, -6, -7
MODE_FAST, MODE_EXACT = 0, 1
MAX_GROUPS = 4
ABI_VERSION = 2
DRIVER_AUTO, DRIVER_GRAPH, DRIVER_ENQUEUE, DRIVER_FUSED = 0, 1, 2, 3

EXPORTS = (
    "halma_last_error", "halma_abi_version", "halma_device_count", "halma_device_info",
    "halma_host_alloc", "halma_host_free", "halma_potential_f32", "halma_potential_workspace_bytes",
    "halma_potential_f32_dev", "halma_plan_create", "halma_plan_destroy", "halma_plan_upload_members",
    "halma_plan_upload_group", "halma_plan_upload_temp", "halma_plan_set_vb", "halma_plan_sync", "halma_nccl_unique_id", "halma_plan_join", "halma_comm_create", "halma_comm_destroy", "halma_plan_use_comm",
    "halma_plan_run", "halma_plan_download", "halma_plan_debug_pass_ns", "halma_unbind_halo", "halma_unbind_catalogue", "halma_microbench",
    "halma_halo_shape_f32", "halma_sigma_projections_f32",
    "halma_snapshot_create", "halma_snapshot_destroy", "halma_snapshot_cells", "halma_snapshot_upload_patch",
    "halma_snapshot_upload_particles", "halma_snapshot_gather", "halma_snapshot_gather_box", "halma_snapshot_fetch", "halma_last_kernel_ms",
    "halma_snapshot_result_device", "halma_snapshot_fetch_star", "halma_selftest_exact_arith",
    "halma_find_target_block",
)


class HalmaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libhalma_unbind error %d: %s" % (code, message))
        self.code = code


class UnbindConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("mode", C.c_int32),
                ("n_groups", C.c_int32), ("n_pre", C.c_int32), ("split_classes", C.c_int32),
                ("vb_fixed", C.c_int32), ("max_iter", C.c_int32), ("G", C.c_double),
                ("kappa", C.c_double), ("rank", C.c_int32), ("n_ranks", C.c_int32),
                ("use_graph", C.c_int32), ("symmetric", C.c_int32),
                ("cache_external", C.c_int32), ("incremental", C.c_int32)]


class HaloResult(C.Structure):
    _fields_ = [("n_bound", C.c_int64), ("n_iter", C.c_int32), ("converged", C.c_int32),
                ("mass", C.c_double), ("com", C.c_double * 3), ("vb", C.c_double * 3),
                ("pairs", C.c_int64), ("most_bound", C.c_int64), ("mass_initial", C.c_double),
                ("cold_bound_mass", C.c_double), ("unbound_cold_mass", C.c_double),
                ("unbound_hot_mass", C.c_double)]


class RunStats(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("potential_ms", C.c_double),
                ("potential_launches", C.c_int32), ("launches", C.c_int32), ("passes", C.c_int32),
                ("driver", C.c_int32), ("pairs", C.c_int64), ("evaluations", C.c_int64),
                ("loop_ms", C.c_double), ("comm_ms", C.c_double), ("comm_bytes", C.c_int64),
                ("phase_ms", C.c_double * 5)]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; raises if it was not built (never falls back to the CPU)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: the CUDA extension was not built (run __graft_entry__.build() or "
            "`make -C pyhalma_b200/csrc`). pyhalma_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.halma_last_error.restype = C.c_char_p
    L.halma_abi_version.restype = i32
    L.halma_device_count.argtypes = [C.POINTER(i32)]
    L.halma_device_info.argtypes = [i32, C.c_char_p, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    L.halma_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.halma_host_free.argtypes = [vp]
    L.halma_potential_f32.argtypes = [i32, i32, vp, vp, vp, vp, i64, vp, vp, vp, i64, vp]
    L.halma_potential_workspace_bytes.restype = i64
    L.halma_potential_workspace_bytes.argtypes = [i64, i64]
    L.halma_potential_f32_dev.argtypes = [i32, i32, vp, vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp]
    L.halma_plan_create.argtypes = [C.POINTER(UnbindConfig), i64, C.POINTER(i64), C.POINTER(C.POINTER(i64)),
                                    C.POINTER(vp)]
    L.halma_plan_destroy.argtypes = [vp]
    L.halma_plan_destroy.restype = None
    L.halma_plan_upload_members.argtypes = [vp] + [vp] * 7
    L.halma_plan_upload_group.argtypes = [vp, i32, vp, vp, vp, vp]
    L.halma_plan_upload_temp.argtypes = [vp, vp, C.c_double]
    L.halma_plan_set_vb.argtypes = [vp, vp]
    L.halma_plan_sync.argtypes = [vp]
    L.halma_nccl_unique_id.argtypes = [vp]
    L.halma_plan_join.argtypes = [vp, vp]
    L.halma_comm_create.argtypes = [i32, i32, i32, vp, C.POINTER(vp)]
    L.halma_comm_destroy.argtypes = [vp]
    L.halma_comm_destroy.restype = None
    L.halma_plan_use_comm.argtypes = [vp, vp]
    L.halma_plan_run.argtypes = [vp, C.POINTER(RunStats)]
    L.halma_plan_download.argtypes = [vp, vp, vp, vp, vp, C.POINTER(HaloResult)]
    L.halma_plan_debug_pass_ns.argtypes = [vp, vp]
    L.halma_microbench.argtypes = [i32, f64p]
    L.halma_halo_shape_f32.argtypes = [i32, vp, vp, vp, vp, i64, vp]
    L.halma_sigma_projections_f32.argtypes = ([i32, i64, vp, C.c_int32, vp, i64] + [vp] * 7 + [C.c_float] * 7 + [vp])
    L.halma_snapshot_create.argtypes = [i32, C.c_double, C.c_int32, i64] + [vp] * 7 + [C.POINTER(vp)]
    L.halma_snapshot_destroy.argtypes = [vp]
    L.halma_snapshot_destroy.restype = None
    L.halma_snapshot_cells.argtypes = [vp]
    L.halma_snapshot_cells.restype = i64
    L.halma_last_kernel_ms.restype = C.c_double
    L.halma_snapshot_upload_patch.argtypes = [vp, i64] + [vp] * 7
    L.halma_snapshot_upload_particles.argtypes = [vp, i32, i64] + [vp] * 5
    L.halma_snapshot_gather.argtypes = [vp] + [C.c_double] * 7 + [vp]
    L.halma_snapshot_gather_box.argtypes = [vp] + [C.c_double] * 6 + [vp]
    L.halma_snapshot_fetch.argtypes = [vp] * 6
    L.halma_snapshot_result_device.argtypes = [vp] * 4
    L.halma_snapshot_fetch_star.argtypes = [vp, i64, vp, vp]
    L.halma_selftest_exact_arith.argtypes = [i32, i64, C.c_uint64, C.POINTER(i64)]
    L.halma_find_target_block.argtypes = [vp, vp, vp, i64, vp, vp, vp, i64]
    L.halma_find_target_block.restype = i64
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("halma_last_error", "halma_potential_workspace_bytes", "halma_plan_destroy",
                        "halma_comm_destroy", "halma_snapshot_destroy", "halma_snapshot_cells", "halma_last_kernel_ms",
                        "halma_find_target_block"):
            fn.restype = i32
    if L.halma_abi_version() != ABI_VERSION:
        raise ImportError("libhalma_unbind ABI version mismatch")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != HALMA_OK:
        raise HalmaError(rc, lib().halma_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().halma_device_count(C.byref(n))
    return n.value if rc == HALMA_OK else 0


def require_device(device: int = 0) -> None:
    """Raise (loudly) when the product path cannot run on a GPU."""
    name = C.create_string_buffer(128)
    check(lib().halma_device_info(device, name, 128, None, None, None))


def device_info(device: int = 0) -> dict:
    name = C.create_string_buffer(128)
    sm, khz, mem = C.c_int(0), C.c_int(0), C.c_int64(0)
    check(lib().halma_device_info(device, name, 128, C.byref(sm), C.byref(khz), C.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sm.value, "clock_khz": khz.value, "mem_bytes": mem.value}


def microbench(device: int = 0) -> dict:
    out = (C.c_double * 8)()
    check(lib().halma_microbench(device, out))
    return {"rsq_per_clk_sm": out[0], "ffma_per_clk_sm": out[1], "ffma2_per_clk_sm": out[2],
            "sm_clock_mhz": out[3], "sm_count": int(out[4]), "rsq_gops": out[5], "ffma_gops": out[6],
            "ffma2_gops": out[7]}


class _PinnedBlock:
    """Page-locked host block exposing the array interface; goes back to the pool when the
    last numpy view of it dies."""

    def __init__(self, ptr: int, size: int, nbytes: int, dtype, count: int):
        self.ptr, self.size = ptr, size
        import numpy as np
        self.__array_interface__ = {"data": (ptr, False), "shape": (count,), "typestr": np.dtype(dtype).str,
                                    "version": 3}

    def __del__(self):
        try:
            _pinned_release(self.ptr, self.size)
        except Exception:
            pass


_PINNED_FREE: dict = {}
_PINNED_CACHED = [0]
_PINNED_CACHE_LIMIT = 8 << 30


def _pinned_release(ptr: int, size: int) -> None:
    if _lib is None:
        return
    if _PINNED_CACHED[0] + size <= _PINNED_CACHE_LIMIT:
        _PINNED_FREE.setdefault(size, []).append(ptr)
        _PINNED_CACHED[0] += size
    else:
        _lib.halma_host_free(C.c_void_p(ptr))


def pinned_empty(count: int, dtype):
    """numpy array in page-locked host memory from a size-bucketed pool (cudaMallocHost costs
    ~0.3 ms/MB, so result buffers are recycled).  Contents are uninitialised."""
    import numpy as np
    nbytes = max(int(count) * np.dtype(dtype).itemsize, 8)
    size = 1 << (nbytes - 1).bit_length() if nbytes < (1 << 20) else ((nbytes + (1 << 20) - 1) >> 20) << 20
    free = _PINNED_FREE.get(size)
    if free:
        ptr = free.pop()
        _PINNED_CACHED[0] -= size
    else:
        p = C.c_void_p()
        check(lib().halma_host_alloc(C.byref(p), size))
        ptr = p.value
    return np.asarray(_PinnedBlock(ptr, size, nbytes, dtype, int(count)))


def ensure_nccl_path() -> None:
    """Point libhalma_unbind at the NCCL build torch ships (split mode only).  Found without
    importing torch; an explicit HALMA_NCCL_LIB wins."""
    if os.environ.get("HALMA_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["HALMA_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def mode_code(mode) -> int:
    if mode in (MODE_FAST, MODE_EXACT) and not isinstance(mode, str):
        return int(mode)
    m = str(mode).lower()
    if m == "fast":
        return MODE_FAST
    if m == "exact":
        return MODE_EXACT
    raise ValueError("mode must be 'fast' or 'exact', got %r" % (mode,))


def default_mode() -> int:
    """HALMA_MODE=fast|exact selects the arithmetic of the drop-in entry points."""
    return mode_code(os.environ.get("HALMA_MODE", "fast"))
