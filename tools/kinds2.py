"""Secondary measurements on the C3 grid: space orders 4 / 16, ISIC imaging condition (LS-RTM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, smooth_model, ricker
ctx = J.get_context(0)


def geom(n):
    ex, ey = (n[0] - 1) * 25., (n[1] - 1) * 25.
    src = np.array([[ex / 2, ey / 2, 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0., ex, 101), np.linspace(0., ey, 101), indexing="ij")
    return src, np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)


def run(name, fn, bpp):
    fn()
    ctx.set_option("timing", 1); ctx.reset_stats()
    fn()
    st = ctx.stencil_stats(); ctx.set_option("timing", 0)
    for k, v in st["kinds"].items():
        ms = v["ms"] / v["launches"]
        print("%-32s %-6s %8.4f ms/launch  (%d launches)  alg %.0f GB/s" % (name, k, ms, v["launches"], bpp[k] / ms / 1e6), flush=True)


n = (401, 401, 201); nt = 30
v = velocity(n); v0 = smooth_model(v)
src, rec = geom(n)
N = 481. * 481 * 281; NP = float(np.prod(n))
dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
which = sys.argv[1:] or ['so', 'ic']
for so in ((4, 16) if 'so' in which else ()):
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=so, nbl=40, m=(1 / v0 ** 2).astype(np.float32))
    q = ricker(nt, float(gm.critical_dt), 0.008)
    run("iso so%d t_sub=1 physical" % so,
        lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=1, snapshot_region="physical"),
        {"plain": 16 * N, "save": 16 * N + 4 * NP, "ic": 16 * N + 12 * NP})
    run("iso so%d forward" % so, lambda: J.forward_rec(gm, src, q, rec), {"plain": 16 * N})
    del gm
dm = (1 / v ** 2 - 1 / v0 ** 2).astype(np.float32)
gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v0 ** 2).astype(np.float32), dm=dm, dt=1.921)
q = ricker(nt, 1.921, 0.008)
for ic in (("isic", "fwi") if 'ic' in which else ()):
    run("born so8 ic=%s" % ic, lambda: J.born_rec(gm, src, q, rec, ic=ic), {"born": 32 * N})
    run("J' so8 ic=%s t_sub=1 padded" % ic,
        lambda: J.J_adjoint(gm, src, q, rec, dobs, is_residual=True, t_sub=1, ic=ic), {"plain": 16 * N, "save": 20 * N, "ic": 28 * N})
