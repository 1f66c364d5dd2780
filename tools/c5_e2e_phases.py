"""Where does a C5 (3-D TTI) end-to-end step spend its time beyond the device-resident step?"""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import judi_jl_b200 as J
import bench as B
shape = (301, 301, 201)
v = B.velocity(shape); v0 = B.smooth_model(v)
kw = dict(space_order=B.SO, nbl=B.NBL); kw.update(B.tti_params(v0))
m0 = np.asfortranarray((1 / v0 ** 2).astype(np.float32))
ctx = J.get_context(0)
def sync(): ctx.synchronize()
def T(label, fn):
    sync(); t0 = time.perf_counter(); r = fn(); sync(); print("%-34s %8.1f ms" % (label, 1e3 * (time.perf_counter() - t0)), flush=True); return r
for rep in range(3):
    print("--- rep", rep)
    mh = T("Model(host arrays)", lambda: J.Model((0., 0., 0.), B.SPACING, shape, m=m0, **kw))
    dt = float(mh.critical_dt); nt = int(np.floor(400.0 / dt)) + 1
    src, rec = B.geometry(shape, 0); q = B.ricker(nt, dt, B.F0)
    d = T("forward_rec (nt=%d)" % nt, lambda: J.forward_rec(mh, src, q, rec)[0])
    shots = [(src, q, rec, d)]
    T("fwi_objective", lambda: J.objectives.fwi_objective(mh, shots, t_sub=4, snapshot_region="physical"))
    T("del model", lambda: mh.__del__() if hasattr(mh, "__del__") else None)
