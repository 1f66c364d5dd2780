"""Per-flavour stencil launch times (CUDA events inside the library) on the BASELINE grids:
plain / +save / +IC / Born for 3-D acoustic (C3/C4), TTI and variable density (C5 grid).
usage: python tools/kinds.py [iso] [born] [tti] [rho] [--nt N]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, smooth_model, ricker

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nt = 42
if "--nt" in sys.argv:
    nt = int(sys.argv[sys.argv.index("--nt") + 1])
which = args or ["iso", "born", "tti", "rho"]
ctx = J.get_context(0)
RES = {}


def geom(n):
    ex, ey = (n[0] - 1) * 25., (n[1] - 1) * 25.
    src = np.array([[ex / 2, ey / 2, 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0., ex, 101), np.linspace(0., ey, 101), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
    return src, rec


def run(name, fn, bpp):
    fn()
    ctx.set_option("timing", 1)
    ctx.reset_stats()
    fn()
    st = ctx.stencil_stats()
    ctx.set_option("timing", 0)
    out = {}
    for k, v in st["kinds"].items():
        ms = v["ms"] / v["launches"]
        out[k] = ms
        b = bpp.get(k)
        print("%-28s %-6s %8.4f ms/launch  (%d launches)%s" % (
            name, k, ms, v["launches"],
            "  alg %.0f GB/s" % (b / ms / 1e6) if b else ""), flush=True)
    RES[name] = out


if "iso" in which or "born" in which:
    n = (401, 401, 201)
    v = velocity(n); v0 = smooth_model(v)
    src, rec = geom(n)
    q = ricker(nt, 1.921, 0.008)
    dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
    N = 481. * 481 * 281; NP = float(np.prod(n))
    if "iso" in which:
        gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v0 ** 2).astype(np.float32), dt=1.921)
        for region, nr in (("physical", NP), ("padded", N)):
            run("iso so8 t_sub=1 " + region,
                lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=1, snapshot_region=region),
                {"plain": 16 * N, "save": 16 * N + 4 * nr, "ic": 16 * N + 12 * nr})
        del gm
    if "born" in which:
        dm = (1 / v ** 2 - 1 / v0 ** 2).astype(np.float32)
        gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v0 ** 2).astype(np.float32), dm=dm, dt=1.921)
        run("born so8", lambda: J.born_rec(gm, src, q, rec), {"born": 32 * N})
        run("lsrtm so8 t_sub=4 physical",
            lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, t_sub=4, snapshot_region="physical"),
            {"born": 32 * N, "plain": 16 * N, "ic": 16 * N + 12 * NP})
        del gm

if "tti" in which or "rho" in which:
    n = (301, 301, 201)
    v = velocity(n)
    src, rec = geom(n)
    N = 381. * 381 * 281; NP = float(np.prod(n))
    if "tti" in which:
        eps = (0.2 * (v - 1.5) / 4).astype(np.float32); dlt = (0.1 * (v - 1.5) / 4).astype(np.float32)
        th = (np.pi / 6 * (v - 1.5) / 4).astype(np.float32)
        gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32),
                     epsilon=eps, delta=dlt, theta=th, phi=np.float32(np.pi / 8))
        dt = float(gm.critical_dt)
        q = ricker(nt, dt, 0.008)
        dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
        run("tti so8 t_sub=1 physical",
            lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=1, snapshot_region="physical"),
            {"plain": 44 * N, "save": 44 * N + 8 * NP, "ic": 44 * N + 16 * NP})
        run("tti so8 forward", lambda: J.forward_rec(gm, src, q, rec), {"plain": 44 * N})
        del gm
    if "rho" in which:
        rho = (0.31 * (v * 1000.) ** 0.25).astype(np.float32)
        gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32), rho=rho)
        dt = float(gm.critical_dt)
        q = ricker(nt, dt, 0.008)
        run("rho so8 forward", lambda: J.forward_rec(gm, src, q, rec), {"plain": 20 * N})
        del gm

os.makedirs("gpurun_out", exist_ok=True)
json.dump(RES, open("gpurun_out/kinds.json", "w"), indent=1)
