"""Per-phase wall times of one C5 (3-D TTI) FWI-gradient shot through the host-buffer path (what
bench.py --workload c5 times as e2e): where the end-to-end step loses against the device-resident one."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, smooth_model, ricker, geometry, tti_params, SPACING, F0, DT
shape = (301, 301, 201)
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ctx = J.get_context(0)
v0 = smooth_model(velocity(shape))
dtc = float(np.float32("%.3e" % (DT / np.sqrt(1 + 2 * (1 + 2 * 0.2)))))
src, rec = geometry(shape, 0)
q = ricker(nt, dtc, F0)
dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32, order="F")
m0 = np.asfortranarray((1 / v0 ** 2).astype(np.float32))
kw = dict(space_order=8, nbl=40, dt=dtc)
kw.update(tti_params(v0))
for a in [m0, q, dobs] + [v for v in kw.values() if isinstance(v, np.ndarray)]:
    J.host_register(a)
shots = [(src, q, rec, dobs)]
for rep in range(4):
    t0 = time.perf_counter()
    gm = J.Model((0., 0., 0.), SPACING, shape, m=m0, **kw)
    ctx.synchronize(); t1 = time.perf_counter()
    ctx.set_option("profile", 1 if rep >= 2 else 0)
    f, g = J.objectives.fwi_objective(gm, shots, t_sub=4, snapshot_region="physical")
    ctx.set_option("profile", 0)
    t2 = time.perf_counter()
    del gm
    ctx.synchronize(); t3 = time.perf_counter()
    print("rep %d: Model %.1f ms, fwi_objective %.1f ms, model destroy %.1f ms" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
