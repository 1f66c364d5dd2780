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
for blk in [(32, 4, 1), (32, 4, 2), (32, 4, 4), (32, 2, 4), (32, 2, 8), (16, 8, 4), (16, 4, 8), (16, 8, 2), (16, 4, 4),
            (32, 1, 8), (32, 1, 16), (16, 2, 16), (8, 8, 8), (8, 8, 4), (16, 16, 2), (32, 8, 2), (32, 8, 1)]:
    ctx.set_option("flux_block", blk[0] | (blk[1] << 8) | (blk[2] << 16))
    J.forward_rec(gm, src, q, rec)
    ctx.set_option("timing", 1); ctx.reset_stats()
    J.forward_rec(gm, src, q, rec)
    st = ctx.stencil_stats(); ctx.set_option("timing", 0)
    print("block %-12s: %.4f ms/step" % (blk, st["ms"] / st["launches"]), flush=True)
