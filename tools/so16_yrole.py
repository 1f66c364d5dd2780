"""acoustic3d_tma at space order 16 with the Y role (ctx option "tile": 8 = consumers read their own
y lines, 34 = Y warps with 4 rows per thread): plain forward steps and an FWI gradient
(snapshot-save + imaging-condition flavours) on the C3 grid; every variant must be bit-identical.
usage: python tools/so16_yrole.py [tile options ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, ricker
n = (401, 401, 201); nt = 62
ctx = J.get_context(0)
v = velocity(n)
src = np.array([[5000., 5000., 50.]], dtype=np.float32)
xr, yr = np.meshgrid(np.linspace(3000., 7000., 9), np.linspace(3000., 7000., 9), indexing="ij")
rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
ref = None
for tile in [int(a) for a in sys.argv[1:]] or [8, 34]:
    ctx.set_option("tile", tile)
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=16, nbl=40, m=(1 / v ** 2).astype(np.float32))
    dt = float(gm.critical_dt)
    q = ricker(nt, dt, 0.008)
    N = float(np.prod(gm.shape_pml))
    d = J.forward_rec(gm, src, q, rec)[0]
    dobs = (0.7 * d).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=2, snapshot_region="physical")
    g = np.array(g)
    if ref is None:
        ref = (d, g)
    best = 1e9
    for _ in range(3):
        ctx.set_option("timing", 1); ctx.reset_stats()
        J.forward_rec(gm, src, q, rec)
        st = ctx.stencil_stats(); ctx.set_option("timing", 0)
        best = min(best, st["ms"] / st["launches"])
    ctx.set_option("timing", 1); ctx.reset_stats()
    J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=2, snapshot_region="physical")
    st = ctx.stencil_stats(); ctx.set_option("timing", 0)
    print("tile %2d: plain %.4f ms/step = alg %.0f GB/s; gradient sweep (save + IC every 2nd step) %.4f ms/step; "
          "bitwise equal to the first: rec %s grad %s" % (
              tile, best, 16 * N / best / 1e6, st["ms"] / st["launches"],
              np.array_equal(d, ref[0]), np.array_equal(g, ref[1])), flush=True)
    del gm
ctx.set_option("tile", 8)
