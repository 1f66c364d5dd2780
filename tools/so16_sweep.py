"""3-D acoustic TMA kernel at space order 16: tile / ring variants on the C3 grid (plain forward steps).
usage: python tools/so16_sweep.py [tile options ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, ricker
n = (401, 401, 201); nt = 62
SO = int(os.environ.get("SO", "16"))
ctx = J.get_context(0)
v = velocity(n)
src = np.array([[5000., 5000., 50.]], dtype=np.float32)
rec = np.array([[5100., 5000., 50.]], dtype=np.float32)
ref = None
for tile in [int(a) for a in sys.argv[1:]] or [8, 21, 22, 23, 24, 25]:
    ctx.set_option("tile", tile)
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=SO, nbl=40, m=(1 / v ** 2).astype(np.float32))
    dt = float(gm.critical_dt)
    q = ricker(nt, dt, 0.008)
    N = float(np.prod(gm.shape_pml))
    d = J.forward_rec(gm, src, q, rec)[0]
    if ref is None:
        ref = d
    best = 1e9
    for _ in range(3):
        ctx.set_option("timing", 1); ctx.reset_stats()
        J.forward_rec(gm, src, q, rec)
        st = ctx.stencil_stats(); ctx.set_option("timing", 0)
        best = min(best, st["ms"] / st["launches"])
    print("tile %2d: %.4f ms/step  alg %.0f GB/s  rel diff vs first %.2e" % (
        tile, best, 16 * N / best / 1e6, np.linalg.norm(d - ref) / np.linalg.norm(ref)), flush=True)
    del gm
