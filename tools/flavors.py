"""Per-flavour stencil launch time on the C3 grid: plain / save+IC (t_sub=1) for several kernel variants."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
n = (401, 401, 201); nt = 50
ctx = J.get_context(0)
v = np.full(n, 2.0, dtype=np.float32); v[:, :, n[2] // 2:] = 3.0
gm = J.Model((0., 0., 0.), (25., 25., 25.), n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32))
dt = float(gm.critical_dt)
src = np.array([[5000., 5000., 50.]], dtype=np.float32)
xr, yr = np.meshgrid(np.linspace(0., 10000., 101), np.linspace(0., 10000., 101), indexing="ij")
rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
t = np.arange(nt) * dt
q = ((1 - 2 * (np.pi * 0.02 * (t - 20)) ** 2) * np.exp(-(np.pi * 0.02 * (t - 20)) ** 2)).astype(np.float32).reshape(nt, 1)
dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
N = float(np.prod(gm.shape_pml)); NP = float(np.prod(n))
def avg(fn):
    fn(); ctx.set_option("timing", 1); ctx.reset_stats(); out = fn(); st = ctx.stencil_stats(); ctx.set_option("timing", 0)
    return st["ms"] / st["launches"], out
tiles = [int(a) for a in sys.argv[1:]] or [6, 8, 9]
ref = None
for tile in tiles:
    ctx.set_option("tile", tile)
    ms0, d = avg(lambda: J.forward_rec(gm, src, q, rec))
    ms1, o1 = avg(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=1, snapshot_region="physical"))
    ms2, o2 = avg(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=1, snapshot_region="padded"))
    ms4, o4 = avg(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4, snapshot_region="physical"))
    if ref is None: ref = (d[0], o1[1])
    print("tile %d: plain %.4f ms (%.0f GB/s) | save+IC phys avg %.4f ms (alg %.0f GB/s) | padded avg %.4f ms (alg %.0f GB/s) | t_sub=4 phys avg %.4f | same=%s %s" % (
        tile, ms0, 16 * N / ms0 / 1e6, ms1, (16 * N + 8 * NP) / ms1 / 1e6, ms2, (16 * N + 8 * N) / ms2 / 1e6, ms4,
        np.array_equal(d[0], ref[0]), np.array_equal(o1[1], ref[1])), flush=True)
