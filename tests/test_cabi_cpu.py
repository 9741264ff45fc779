"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/halma_unbind.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from pyhalma_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "halma_unbind.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(halma_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 22
    for n in names:
        assert hasattr(L, n), "libhalma_unbind.so does not export %s" % n
    assert set(names) == set(_lib.EXPORTS)


def test_struct_sizes_match_header(tmp_path):
    assert ctypes.sizeof(_lib.UnbindConfig) == 72
    assert ctypes.sizeof(_lib.HaloResult) == 120
    assert ctypes.sizeof(_lib.RunStats) == 112
    # the same numbers from the header itself, through the C compiler
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "halma_unbind.h"\nint main(void){printf("%zu %zu %zu\\n", '
                   'sizeof(halma_unbind_config), sizeof(halma_halo_result), sizeof(halma_run_stats));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(_lib.UnbindConfig), ctypes.sizeof(_lib.HaloResult),
                                     ctypes.sizeof(_lib.RunStats)]


def test_abi_version():
    assert _lib.lib().halma_abi_version() == _lib.ABI_VERSION == 2


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from pyhalma_b200 import particle
    with pytest.raises(_lib.HalmaError) as ei:
        particle.brute_force_binding_energy(1, 2, np.ones(2), np.ones(2), np.ones(2), np.ones(2), 1, [0.], [0.], [0.])
    assert ei.value.code == _lib.ERR_NO_DEVICE


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.halma_potential_f32(0, 7, None, None, None, None, 0, None, None, None, 0, None) == _lib.ERR_INVALID
    assert L.halma_potential_f32(0, 0, None, None, None, None, -1, None, None, None, 0, None) == _lib.ERR_INVALID
    assert L.halma_potential_f32(0, 0, None, None, None, None, 0, None, None, None, 0, None) == _lib.HALMA_OK
    assert L.halma_potential_workspace_bytes(10, 10) >= 1024 + 8 * 12 * 8
    assert L.halma_halo_shape_f32(0, None, None, None, None, 3, None) == _lib.ERR_INVALID
    e3 = np.zeros(3, np.float32)
    assert L.halma_halo_shape_f32(0, None, None, None, None, 3, e3.ctypes.data) == _lib.ERR_INVALID
    g3 = np.zeros(3, np.float32)
    o5 = np.zeros(5, np.float32)
    bad = np.array([7], np.int32)
    one = np.zeros(1, np.float32)
    assert L.halma_sigma_projections_f32(0, 1, g3.ctypes.data, 3, bad.ctypes.data, 1, *[one.ctypes.data] * 7,
                                         0., 0., 0., 1., 1., 1., .5, o5.ctypes.data) == _lib.ERR_INVALID
    assert b"part_list" in L.halma_last_error()
    cfg = _lib.UnbindConfig()
    cfg.struct_size = 12          # wrong size -> ABI guard
    h = ctypes.c_void_p()
    off = (ctypes.c_int64 * 2)(0, 0)
    assert L.halma_plan_create(ctypes.byref(cfg), 1, off, None, ctypes.byref(h)) == _lib.ERR_INVALID
    assert b"size mismatch" in L.halma_last_error()


def test_f2py_shape_errors():
    from pyhalma_b200 import particle
    with pytest.raises(ValueError):
        particle.brute_force_binding_energy(1, 3, np.ones(2), np.ones(3), np.ones(3), np.ones(3), 1, [0.], [0.], [0.])
    with pytest.raises(ValueError):
        particle.serial_brute_force_binding_energy(2, np.ones(2), np.ones(2), np.ones(2), np.ones(2), 2, [0.], [0.], [0.])


def test_wrapper_empty_targets_needs_no_gpu():
    from pyhalma_b200 import halo_gas
    out = halo_gas.brute_force_binding_energy_fortran(np.ones(3), np.ones(3), np.ones(3), np.ones(3), [], [], [])
    assert out.shape == (0,) and out.dtype == np.float64        # halo_gas.py:169-170


def test_halo_results_view_matches_struct_layout():
    """unbind.HaloResults reads halma_halo_result records in place: its numpy dtype must mirror the
    ctypes structure (and through test_struct_sizes_match_header the C header)."""
    from pyhalma_b200.unbind import _HALO_DTYPE, HaloResults
    assert _HALO_DTYPE.itemsize == ctypes.sizeof(_lib.HaloResult)
    for (name, _t), f in zip(_lib.HaloResult._fields_, _HALO_DTYPE.names):
        assert name == f and getattr(_lib.HaloResult, name).offset == _HALO_DTYPE.fields[f][1], name
    hr = (_lib.HaloResult * 3)()
    hr[1].n_bound, hr[1].mass, hr[1].most_bound = 7, 2.5, 4
    hr[1].com[2] = 9.0
    r = HaloResults(hr, 3)
    assert len(r) == 3 and r[1].n_bound == 7 and r[-2].mass == 2.5 and r[1].com == (0.0, 0.0, 9.0)
    assert [h.most_bound for h in r] == [0, 4, 0] and len(r[0:2]) == 2
    assert list(r.field("n_bound")) == [0, 7, 0] and r.field("com").shape == (3, 3)
    with pytest.raises(IndexError):
        r[3]


def test_find_target_block_host_helper():
    """halma_potential_f32 runs a large call as a plan when the targets are a block of the sources
    (csrc/api.cu::potential_via_plan); the search for that block is plain host code."""
    L = _lib.lib()
    rng = np.random.default_rng(3)
    f32 = np.float32

    def find(src, tgt):
        s = [np.ascontiguousarray(a, f32) for a in src]
        t = [np.ascontiguousarray(a, f32) for a in tgt]
        return L.halma_find_target_block(*[a.ctypes.data for a in s], len(s[0]), *[a.ctypes.data for a in t], len(t[0]))

    x, y, z = rng.normal(size=(3, 1000))
    assert find((x, y, z), (x, y, z)) == 0                                        # self call (separate copies)
    assert find((x, y, z), (x[300:700], y[300:700], z[300:700])) == 300           # block in the middle
    assert find((x, y, z), (x[:10], y[:10], z[:10])) == 0 and find((x, y, z), (x[990:], y[990:], z[990:])) == 990
    assert find((x[:500], y[:500], z[:500]), (x, y, z)) == -1                     # more targets than sources
    y2 = y.copy()
    y2[650] += 1.0
    assert find((x, y, z), (x[300:700], y2[300:700], z[300:700])) == -1           # one coordinate differs
    assert find((x, y, z), (x[300:700][::-1], y[300:700][::-1], z[300:700][::-1])) == -1      # other order
    assert find((x, y, z), (x[:0], y[:0], z[:0])) == -1 and find((x[:0], y[:0], z[:0]), (x[:0], y[:0], z[:0])) == -1
    # a lattice: many sources share the first target's x; the block itself is still the first candidate
    gx, gy, gz = [a.ravel() for a in np.meshgrid(np.arange(8.0), np.arange(8.0), np.arange(8.0), indexing="ij")]
    assert find((gx, gy, gz), (gx, gy, gz)) == 0
    assert find((np.r_[x, gx], np.r_[y, gy], np.r_[z, gz]), (gx, gy, gz)) == 1000
    # ... and when it is not among the first 64 sources with that x the search gives up (direct path)
    assert find((np.r_[gx, gx[:100], gx], np.r_[gy, gy[:100] + 0.5, gy], np.r_[gz, gz[:100], gz]),
                (gx[:100], gy[:100] + 0.5, gz[:100])) in (-1, 512)
    # -0.0 and +0.0 are different bit patterns: not a block (the call then takes the direct path, same result)
    assert find(([0.0, 1.0], [0.0, 1.0], [0.0, 1.0]), ([-0.0, 1.0], [0.0, 1.0], [0.0, 1.0])) == -1
