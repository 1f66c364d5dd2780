"""Variable-density single-pass kernel: z-chunk sweep on the C5-size grid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, ricker
n = (301, 301, 201); nt = 22
v = velocity(n)
rho = (0.31 * (v * 1000.) ** 0.25).astype(np.float32)
ctx = J.get_context(0)
gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32), rho=rho)
q = ricker(nt, float(gm.critical_dt), 0.008)
src = np.array([[3750., 3750., 50.]], dtype=np.float32)
rec = np.array([[3000., 3000., 50.]], dtype=np.float32)
ref = None
for fused, zc in [(0, 0), (1, 0), (1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (1, 8)]:
    ctx.set_option("varc_fused", fused); ctx.set_option("zchunks", zc)
    gm2 = gm
    d1 = J.forward_rec(gm2, src, q, rec)[0]
    ctx.set_option("timing", 1); ctx.reset_stats()
    J.forward_rec(gm2, src, q, rec)
    st = ctx.stencil_stats(); ctx.set_option("timing", 0)
    if ref is None: ref = d1
    print("fused %d zchunks %d: %.4f ms/step  rel diff vs two-pass %.2e" % (
        fused, zc, st["ms"] / st["launches"], np.linalg.norm(d1 - ref) / np.linalg.norm(ref)), flush=True)
