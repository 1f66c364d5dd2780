"""Round-2 per-flavour stencil launch times (CUDA events inside the library) on the BASELINE grids.
usage: python tools/kinds_r2.py [so] [ext] [isic] [--nt N]
  so    TTI / variable density at space orders 4, 8, 16 on the C5 grid (single-pass vs two-pass)
  ext   3-D extended / full-wavefield sources and extended receiver on the C3 grid (TMA vs generic)
  isic  isic / fwi Born step and imaging step on the C3 grid
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, smooth_model, ricker

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nt = 30
if "--nt" in sys.argv:
    nt = int(sys.argv[sys.argv.index("--nt") + 1])
which = args or ["so", "ext", "isic"]
ctx = J.get_context(0)


def geom(n, nr=101):
    ex, ey = (n[0] - 1) * 25., (n[1] - 1) * 25.
    src = np.array([[ex / 2, ey / 2, 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0., ex, nr), np.linspace(0., ey, nr), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
    return src, rec


def run(name, fn):
    fn()
    ctx.set_option("timing", 1)
    ctx.reset_stats()
    fn()
    st = ctx.stencil_stats()
    ctx.set_option("timing", 0)
    print("%-44s %s" % (name, "  ".join("%s %.4f ms (%d)" % (k, v["ms"] / v["launches"], v["launches"])
                                        for k, v in st["kinds"].items())), flush=True)


if "so" in which:
    n = (301, 301, 201)
    v = velocity(n)
    src, rec = geom(n)
    for so in (4, 8, 16):
        for kind in ("iso", "tti", "rho", "ttirho"):
            if kind == "ttirho" and so == 16:
                continue
            kw = dict(m=(1 / v ** 2).astype(np.float32))
            if kind in ("tti", "ttirho"):
                kw.update(epsilon=(0.2 * (v - 1.5) / 4).astype(np.float32), delta=(0.1 * (v - 1.5) / 4).astype(np.float32),
                          theta=(np.pi / 6 * (v - 1.5) / 4).astype(np.float32), phi=np.float32(np.pi / 8))
            if kind in ("rho", "ttirho"):
                kw["rho"] = ((v + .5) / 2).astype(np.float32)
            gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=so, nbl=40, **kw)
            dt = float(gm.critical_dt)
            q = ricker(nt, dt, 0.008)
            dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
            for opt, label in ((1, "single-pass"), (0, "two-pass")):
                if kind == "iso" and not opt:
                    continue
                ctx.set_option("varc_fused" if kind == "rho" else "tti_fused", opt if kind == "rho" else (2 if opt else 0))
                run("%s so%d %s" % (kind, so, label),
                    lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=2, snapshot_region="physical"))
            ctx.set_option("tti_fused", 2); ctx.set_option("varc_fused", 1)
            del gm

if "ext" in which or "isic" in which:
    n = (401, 401, 201)
    v = velocity(n); v0 = smooth_model(v)
    src, rec = geom(n)
    q = ricker(nt, 1.921, 0.008)
    dm = (1 / v ** 2 - 1 / v0 ** 2).astype(np.float32)
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v0 ** 2).astype(np.float32), dm=dm, dt=1.921)
    y = np.random.default_rng(0).standard_normal((nt, rec.shape[0])).astype(np.float32)
    if "ext" in which:
        w = np.zeros(gm.shape_pml, np.float32)
        w[200:280, 200:280, 100:140] = 1.0
        for kern, label in ((1, "TMA"), (0, "generic")):
            ctx.set_option("acoustic3d_kernel", kern)
            run("forward_rec_w (%s)" % label, lambda: J.forward_rec_w(gm, w, q, rec))
            run("adjoint_w (%s)" % label, lambda: J.adjoint_w(gm, rec, y, q, fw=False))
            run("born_rec_w (%s)" % label, lambda: J.born_rec_w(gm, w, q, rec))
        ctx.set_option("acoustic3d_kernel", 1)
    if "isic" in which:
        for ic in ("as", "isic", "fwi"):
            run("born_rec ic=%s" % ic, lambda: J.born_rec(gm, src, q, rec, ic=ic))
            run("J_adjoint ic=%s t_sub=4" % ic,
                lambda: J.J_adjoint(gm, src, q, rec, y, is_residual=True, ic=ic, t_sub=4))
