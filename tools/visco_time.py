"""3-D visco-acoustic step times on the C5-size grid (generic kernels)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, ricker
n = (301, 301, 201); nt = 42
v = velocity(n)
qp = (3.516 * ((v * 1000.) ** 2.2) * 1e-6).astype(np.float32)
ctx = J.get_context(0)
src = np.array([[3750., 3750., 50.]], dtype=np.float32)
rec = np.array([[3700., 3750., 50.]], dtype=np.float32)
for name, kw in (("constant density", {}), ("density field", {"rho": (0.31 * (v * 1000.) ** 0.25).astype(np.float32)})):
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32), qp=qp, f0=0.008, **kw)
    q = ricker(nt, float(gm.critical_dt), 0.008)
    for fw in (True, False):
        J.forward_rec(gm, src, q, rec, fw=fw); ctx.synchronize()
        t0 = time.perf_counter()
        J.forward_rec(gm, src, q, rec, fw=fw)
        ctx.synchronize()
        print("visco %s %s: %.3f ms per step" % (name, "forward" if fw else "adjoint", 1e3 * (time.perf_counter() - t0) / (nt - 2)), flush=True)
