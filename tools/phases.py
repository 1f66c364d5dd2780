"""Per-phase wall times of one FWI-gradient shot of the bench workload (option "profile")."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, smooth_model, ricker, geometry, SHAPE, SPACING, DT, F0
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1042
ctx = J.get_context(0)
v0 = smooth_model(velocity(SHAPE))
src, rec = geometry(SHAPE, 0)
q = ricker(nt, DT, F0)
dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
import time
m0 = np.asfortranarray((1 / v0 ** 2).astype(np.float32))
for rep in range(4):
    if rep == 2:   # the same call with page-locked caller arrays (judi_host_register)
        print("--- host arrays registered", flush=True)
        J.host_register(m0); J.host_register(q); J.host_register(dobs)
        gbuf = J.host_register(np.zeros(gm.shape_pml, dtype=np.float32, order="F"))
    t0 = time.perf_counter()
    gm = J.Model((0., 0., 0.), SPACING, SHAPE, space_order=8, nbl=40, m=m0, dt=DT)
    ctx.synchronize(); t1 = time.perf_counter()
    ctx.set_option("profile", rep % 2)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4, snapshot_region="physical",
                             grad_out=gbuf if rep >= 2 else None)
    ctx.set_option("profile", 0)
    t2 = time.perf_counter()
    print("rep %d: Model %.1f ms, J_adjoint %.1f ms" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
