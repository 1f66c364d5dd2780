"""2-D shots (C2 grid, 401x121 @25 m, nt ~ 1136): forward + FWI gradient wall time per physics."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import ricker
ctx = J.get_context(0)
n = (401, 121); d = (25., 25.)
x = np.arange(n[0])[:, None] * 25.; z = np.arange(n[1])[None, :] * 25.
v = np.clip(1.5 + (z / 3000.) * 3.0 + 0.4 * np.sin(2 * np.pi * x / 5000.) * (z / 3000.), 1.5, 5.5).astype(np.float32)
v[:, :21] = 1.5
src = np.array([[5000., 50.]], dtype=np.float32)
rec = np.stack([np.linspace(0, 10000., 401), np.full(401, 50.)], axis=1).astype(np.float32)
blocks = [int(b) for b in os.environ.get("PERSIST_BLOCKS", "0").split(",")]
ctx.set_option("persist2d", int(os.environ.get("PERSIST2D", "1")))
ic = os.environ.get("IC", "as")      # imaging condition: as / isic / fwi
for kind in [a for a in sys.argv[1:]] or ["iso", "rho", "tti"]:
    kw = {}
    if kind == "rho": kw["rho"] = (0.31 * (v * 1000.) ** 0.25).astype(np.float32)
    if kind in ("visco", "viscorho"):
        kw["qp"] = (3.516 * ((v * 1000.) ** 2.2) * 1e-6).astype(np.float32); kw["f0"] = 0.008
        if kind == "viscorho": kw["rho"] = (0.31 * (v * 1000.) ** 0.25).astype(np.float32)
    if kind == "tti":
        kw.update(epsilon=(0.2 * (v - 1.5) / 4).astype(np.float32), delta=(0.1 * (v - 1.5) / 4).astype(np.float32),
                  theta=(np.pi / 6 * (v - 1.5) / 4).astype(np.float32))
    gm = J.Model((0., 0.), d, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32), **kw)
    dt = float(gm.critical_dt); nt = int(3000. / dt) + 1
    q = ricker(nt, dt, 0.008)
    d0, _ = J.forward_rec(gm, src, q, rec); dobs = (0.9 * d0).astype(np.float32)
    for nb in blocks:
        ctx.set_option("persist_blocks", nb)
        J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, ic=ic); ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): J.forward_rec(gm, src, q, rec)
        ctx.synchronize(); t1 = time.perf_counter()
        for _ in range(3): J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, ic=ic)
        ctx.synchronize(); t2 = time.perf_counter()
        print("%s 2-D 481x201 nt=%d blocks<=%d ic=%s: forward %.2f ms  fwi gradient %.2f ms  (%.2f us per step pair)" % (
            kind, nt, nb, ic, 1e3 * (t1 - t0) / 3, 1e3 * (t2 - t1) / 3, 1e6 * (t2 - t1) / 3 / nt), flush=True)
        if kind == "iso":
            # LS-RTM: Born forward sweep + adjoint (lsrtm_objective, one shot)
            gb = J.Model((0., 0.), d, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32),
                         dm=(0.02 * np.sin(z / 300.) * np.ones_like(v)).astype(np.float32))
            J.born_rec(gb, src, q, rec, ic=ic); ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(3): J.born_rec(gb, src, q, rec, ic=ic)
            ctx.synchronize(); t1 = time.perf_counter()
            for _ in range(3): J.J_adjoint(gb, src, q, rec, dobs, return_obj=True, ic=ic, born_fwd=True)
            ctx.synchronize(); t2 = time.perf_counter()
            print("%s 2-D 481x201 nt=%d ic=%s: born_rec %.2f ms  lsrtm gradient %.2f ms" % (
                kind, nt, ic, 1e3 * (t1 - t0) / 3, 1e3 * (t2 - t1) / 3), flush=True)
