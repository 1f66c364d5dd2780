"""Single-pass 3-D TTI kernel against the two-pass form on the C5 grid: time per step of the
plain / save / IC flavours and relative difference of shot record and gradient.
usage: python tools/tti_tune.py [nt] [variants e.g. 0,1,2] [zchunks e.g. 0,1,2] [space order] [rho]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, ricker
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 40
variants = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
zcs = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
so = int(sys.argv[4]) if len(sys.argv) > 4 else 8
rho = len(sys.argv) > 5
n = (301, 301, 201)
v = velocity(n)
eps = (0.2 * (v - 1.5) / 4).astype(np.float32); dlt = (0.1 * (v - 1.5) / 4).astype(np.float32)
th = (np.pi / 6 * (v - 1.5) / 4).astype(np.float32)
ctx = J.get_context(0)
dm = (0.05 * (1 / v ** 2) * np.sin(np.arange(n[2]) / 7.)[None, None, :]).astype(np.float32)
kw = {"rho": ((v + .5) / 2).astype(np.float32)} if rho else {}
print("space order %d%s" % (so, ", density field" if rho else ""))
gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=so, nbl=40, m=(1 / v ** 2).astype(np.float32), **kw,
             epsilon=eps, delta=dlt, theta=th, phi=np.float32(np.pi / 8), dm=dm)
dt = float(gm.critical_dt)
q = ricker(nt, dt, 0.03)
src = np.array([[3750., 3750., 50.]], dtype=np.float32)
xr, yr = np.meshgrid(np.linspace(3500., 4000., 11), np.linspace(3500., 4000., 11), indexing="ij")
rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
ref = None
for fused in variants:
    for zc in (zcs if fused else [0]):
        ctx.set_option("tti_fused", fused); ctx.set_option("zchunks", zc)
        d1 = J.forward_rec(gm, src, q, rec)[0]
        f, g = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=1, snapshot_region="physical")[:2]
        ctx.set_option("timing", 1); ctx.reset_stats()
        J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=1, snapshot_region="physical")
        st = ctx.stencil_stats(); ctx.reset_stats()
        dl = J.born_rec(gm, src, q, rec)[0]
        stb = ctx.stencil_stats(); ctx.set_option("timing", 0)
        if ref is None: ref = (d1, g, dl)
        ms = {k: v_["ms"] / v_["launches"] for k, v_ in st["kinds"].items()}
        ms["born"] = stb["ms"] / stb["launches"]
        print("tti_fused %d zchunks %d: %s  |d|=%.4e |g|=%.4e  rel diff vs first: rec %.2e grad %.2e born %.2e" % (
            fused, zc, " ".join("%s %.4f ms" % kv for kv in sorted(ms.items())),
            np.linalg.norm(d1), np.linalg.norm(g),
            np.linalg.norm(d1 - ref[0]) / np.linalg.norm(ref[0]),
            np.linalg.norm(g - ref[1]) / np.linalg.norm(ref[1]),
            np.linalg.norm(dl - ref[2]) / np.linalg.norm(ref[2])), flush=True)
