"""Quick throughput probe of the potential kernel (device-resident, CUDA-event timed)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyhalma_b200 import _lib, synth

L = _lib.lib()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for mode in (0, 1):
    for n_src, n_tgt in ((20_000, 10_000), (200_000, 200_000), (700_000, 200_000), (1_000_000, 1_000_000)):
        if mode == 1 and n_src * n_tgt > 2e11:
            continue
        p = synth.plummer_stars(n_src, 5e-3, 1e6, rng)
        src = [torch.tensor(np.float32(a), device=dev) for a in (p.mass, p.x, p.y, p.z)]
        tgt = [t[:n_tgt].clone() for t in src[1:]]
        out = torch.zeros(n_tgt, dtype=torch.float32, device=dev)
        ws = torch.zeros(L.halma_potential_workspace_bytes(n_src, n_tgt) // 4 + 64, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        def run():
            _lib.check(L.halma_potential_f32_dev(0, mode, *[t.data_ptr() for t in src], n_src,
                                                 *[t.data_ptr() for t in tgt], n_tgt, out.data_ptr(), ws.data_ptr(), st))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("mode=%s n_src=%d n_tgt=%d  %.3f ms  %.1f Ginter/s" % ("fast" if mode == 0 else "exact", n_src, n_tgt, ms,
                                                                   n_src * n_tgt / ms / 1e6), flush=True)
