import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pyhalma_b200 import gather, halo_gas, pipeline, synth
from pyhalma_b200.unbind import UnbindPlan

s = synth.amr_snapshot(n_levels=8, patches_per_level=4, max_cells=64, n_dm=2_000_000, n_st=1_000_000)
rng = np.random.default_rng(11)
centres = np.asarray(s.centre) + rng.normal(0, 0.004, (200, 3))
radii = 10 ** rng.uniform(np.log10(0.006), np.log10(0.03), 200)
snap = gather.Snapshot(s.L, s.ncoarse, s.grid_data, s.gas_data, s.masclet_dm_data, s.masclet_st_data)
T = time.perf_counter
one = lambda n: np.array([0, n], np.int64)

def mb_device(g, acc):
    ext = [g.gas_source(), g.dm_source(), g.dm_source(light=True)]
    zeros = np.zeros(g.n_st)
    t0 = T()
    plan = UnbindPlan(one(g.n_st), [one(n) for n, _ in ext], mode="fast", n_pre=3, split_classes=True, vb_fixed=True,
                      max_iter=1, kappa=1.0)
    t1 = T()
    z = zeros.ctypes.data
    plan.upload_members_raw(g.col(3, 0), g.col(3, 1), g.col(3, 2), z, z, z, g.col(3, 3))
    for k, (n, cols) in enumerate(ext):
        plan.upload_group_raw(k, *cols)
    plan.set_vb([0.0, 0.0, 0.0])
    t2 = T()
    st = plan.run()
    t3 = T()
    plan.download(mask=False, be=False, energy=False, idx=False)
    t4 = T()
    plan.close()
    t5 = T()
    for k, v in zip(("create", "upload", "run", "download", "close", "dev_total", "dev_pot"),
                    (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, st.total_ms / 1e3, st.potential_ms / 1e3)):
        acc[k] = acc.get(k, 0) + v

def mb_host(h, acc):
    gas_x, gas_y, gas_z, gas_mass = h[0], h[1], h[2], h[6]
    hv, lt = halo_gas._split_dm(*h[8:12], 8e7)
    ext = [(gas_mass, gas_x, gas_y, gas_z), (hv[3], hv[0], hv[1], hv[2]), (lt[3], lt[0], lt[1], lt[2])]
    nst = len(h[12]); zeros = np.zeros(nst)
    t0 = T()
    plan = UnbindPlan(one(nst), [one(len(g[0])) for g in ext], mode="fast", n_pre=3, split_classes=True, vb_fixed=True,
                      max_iter=1, kappa=1.0)
    t1 = T()
    plan.upload_members(h[12], h[13], h[14], zeros, zeros, zeros, h[15])
    for k, g in enumerate(ext):
        plan.upload_group(k, *g)
    plan.set_vb([0., 0., 0.])
    t2 = T()
    st = plan.run()
    t3 = T()
    plan.download(mask=False, be=False, energy=False, idx=False)
    t4 = T()
    plan.close()
    t5 = T()
    for k, v in zip(("create", "upload", "run", "download", "close", "dev_total", "dev_pot"),
                    (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, st.total_ms / 1e3, st.potential_ms / 1e3)):
        acc[k] = acc.get(k, 0) + v

n = 60
for order in ("device", "host", "device", "host"):
    acc = {}
    t0 = T()
    for c, R in zip(centres[:n], radii[:n]):
        if order == "device":
            g = snap.gather_device(*c, R, s.rho_B, s.rete, dm_heavy_min=0.9e7)
            if g.n_st: mb_device(g, acc)
        else:
            h = snap.gather(*c, R, s.rho_B, s.rete)
            if len(h[12]): mb_host(h, acc)
    print(order, "%.2f s" % (T() - t0), {k: round(v, 3) for k, v in acc.items()}, flush=True)
