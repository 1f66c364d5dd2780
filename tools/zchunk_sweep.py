"""3-D acoustic TMA kernel: z-chunk sweep on the C3 grid (plain forward steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, ricker
n = (401, 401, 201); nt = 62
ctx = J.get_context(0)
v = velocity(n)
gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32), dt=1.921)
q = ricker(nt, 1.921, 0.008)
src = np.array([[5000., 5000., 50.]], dtype=np.float32)
rec = np.array([[5100., 5000., 50.]], dtype=np.float32)
N = float(np.prod(gm.shape_pml))
for zc in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12]:
    ctx.set_option("zchunks", zc)
    J.forward_rec(gm, src, q, rec)
    best = 1e9
    for _ in range(3):
        ctx.set_option("timing", 1); ctx.reset_stats()
        J.forward_rec(gm, src, q, rec)
        st = ctx.stencil_stats(); ctx.set_option("timing", 0)
        best = min(best, st["ms"] / st["launches"])
    print("zchunks %2d: %.4f ms/step  alg %.0f GB/s" % (zc, best, 16 * N / best / 1e6), flush=True)
