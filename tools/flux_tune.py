"""TTI / variable-density flux-kernel tuning sweep: thread-block shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, ricker
n = (301, 301, 201); nt = 22
v = velocity(n)
eps = (0.2 * (v - 1.5) / 4).astype(np.float32); dlt = (0.1 * (v - 1.5) / 4).astype(np.float32)
th = (np.pi / 6 * (v - 1.5) / 4).astype(np.float32)
ctx = J.get_context(0)
gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32),
             epsilon=eps, delta=dlt, theta=th, phi=np.float32(np.pi / 8))
q = ricker(nt, float(gm.critical_dt), 0.008)
src = np.array([[3750., 3750., 50.]], dtype=np.float32)
rec = np.array([[3000., 3000., 50.]], dtype=np.float32)
for mb in (0, 8, 10, 12):
    ctx.set_option("flux_minb", mb)
    d1 = J.forward_rec(gm, src, q, rec)[0]
    ctx.set_option("timing", 1); ctx.reset_stats()
    J.forward_rec(gm, src, q, rec)
    st = ctx.stencil_stats(); ctx.set_option("timing", 0)
    if mb == 0: ref = d1
    print("minb %3d: %.4f ms/step  same=%s" % (mb, st["ms"] / st["launches"], np.array_equal(d1, ref)), flush=True)
