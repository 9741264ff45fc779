"""Throughput of the FAST potential kernel for the variant selected by HALMA_FAST_VARIANT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyhalma_b200 import _lib, synth

L = _lib.lib()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
v = os.environ.get("HALMA_FAST_VARIANT", "0")
out_line = ["variant=%s" % v]
for n_src, n_tgt in ((200_000, 200_000), (700_000, 500_000)):
    p = synth.plummer_stars(n_src, 5e-3, 1e6, rng)
    src = [torch.tensor(np.float32(a), device=dev) for a in (p.mass, p.x, p.y, p.z)]
    tgt = [t[:n_tgt].clone() for t in src[1:]]
    out = torch.zeros(n_tgt, dtype=torch.float32, device=dev)
    ws = torch.zeros(L.halma_potential_workspace_bytes(n_src, n_tgt) // 4 + 64, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def run():
        _lib.check(L.halma_potential_f32_dev(0, 0, *[t.data_ptr() for t in src], n_src,
                                             *[t.data_ptr() for t in tgt], n_tgt, out.data_ptr(), ws.data_ptr(), st))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out_line.append("%dx%d: %.2f ms %.0f G/s" % (n_src, n_tgt, ms, n_src * n_tgt / ms / 1e6))
    if n_src == 200_000:
        ref = out.double().sum().item()
        out_line.append("sum=%.9e" % ref)
print("  ".join(out_line), flush=True)
