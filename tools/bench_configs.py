"""Timing of the secondary BASELINE configs (C2 2-D FWI, C4 3-D Born/LS-RTM, C5 3-D TTI) -- short runs."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import velocity, smooth_model, ricker
RES = []
def rep(name, **kw):
    RES.append(dict(name=name, **kw)); print("%-40s" % name, " ".join("%s=%s" % (k, ("%.4g" % v) if isinstance(v, float) else v) for k, v in kw.items()), flush=True)
ctx = J.get_context(0)

def timed(fn, reps=2):
    fn(); ctx.synchronize()
    ctx.set_option("timing", 1); ctx.reset_stats()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    ctx.synchronize(); wall = (time.perf_counter() - t0) / reps
    st = ctx.stencil_stats(); ctx.set_option("timing", 0)
    return wall, st["ms"] / max(st["launches"], 1), st["launches"] // reps, st["points"] / reps, out

def c2_2d():
    n = (401, 121); d = (25., 25.)
    x = np.arange(n[0])[:, None] * 25.; z = np.arange(n[1])[None, :] * 25.
    v = np.clip(1.5 + (z / 3000.) * 3.0 + 0.4 * np.sin(2 * np.pi * x / 5000.) * (z / 3000.), 1.5, 5.5).astype(np.float32)
    v[:, :21] = 1.5
    gm = J.Model((0., 0.), d, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32))
    dt = float(gm.critical_dt); nt = int(3000. / dt) + 1
    q = ricker(nt, dt, 0.008)
    src = np.array([[5000., 50.]], dtype=np.float32)
    rec = np.stack([np.linspace(0, 10000., 401), np.full(401, 50.)], axis=1).astype(np.float32)
    d0, _ = J.forward_rec(gm, src, q, rec); dobs = (0.9 * d0).astype(np.float32)
    wall, ms, nl, pts, _ = timed(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True), reps=4)
    N = np.prod(gm.shape_pml)
    rep("C2 2D fwi shot 481x201 nt=%d" % nt, wall_ms=wall * 1e3, stencil_ms=ms, launches=nl, us_per_step=wall * 1e6 / (2 * nt), gpts=2 * nt * N / wall / 1e9, shots_per_s=1 / wall)

def c3_forward(n=(401, 401, 201), nt=1042):
    v = velocity(n)
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=np.asfortranarray((1 / v ** 2).astype(np.float32)), dt=1.921)
    q = ricker(nt, 1.921, 0.008)
    src = np.array([[5000., 5000., 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0., 10000., 101), np.linspace(0., 10000., 101), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
    N = float(np.prod(gm.shape_pml))
    wall, ms, nl, pts, out = timed(lambda: J.forward_rec(gm, src, q, rec), reps=2)
    rep("C3 3D forward shot nt=%d (10201 rec)" % nt, wall_ms=wall * 1e3, ms_per_step=wall * 1e3 / (nt - 2),
        gpts=(nt - 2) * N / wall / 1e9, shots_per_s=1 / wall, norm=float(np.linalg.norm(out[0])))


def c4_born(n=(401, 401, 201), nt=120):
    v = velocity(n); v0 = smooth_model(v)
    dm = (1 / v ** 2 - 1 / v0 ** 2).astype(np.float32)
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v0 ** 2).astype(np.float32), dm=dm, dt=1.921)
    q = ricker(nt, 1.921, 0.008)
    src = np.array([[5000., 5000., 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0., 10000., 101), np.linspace(0., 10000., 101), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
    N = float(np.prod(gm.shape_pml))
    for kern in (1, 0):
        ctx.set_option("acoustic3d_kernel", kern)
        wall, ms, nl, pts, out = timed(lambda: J.born_rec(gm, src, q, rec), reps=1)
        rep("C4 3D born_rec kern=%d nt=%d" % (kern, nt), wall_ms=wall * 1e3, ms_per_step=wall * 1e3 / (nt - 2), gpts=2 * (nt - 2) * N / wall / 1e9, norm=float(np.linalg.norm(out[0])))
    ctx.set_option("acoustic3d_kernel", 1)
    dl = out[0]
    wall, ms, nl, pts, out = timed(lambda: J.J_adjoint(gm, src, q, rec, dl, is_residual=True, t_sub=4, snapshot_region="physical"), reps=1)
    rep("C4 3D J' (is_residual) nt=%d" % nt, wall_ms=wall * 1e3, ms_per_step_pair=wall * 1e3 / (nt - 2), gnorm=float(np.linalg.norm(out[0])))
    wall, ms, nl, pts, out = timed(lambda: J.J_adjoint(gm, src, q, rec, dl, return_obj=True, born_fwd=True, t_sub=4, snapshot_region="physical"), reps=1)
    rep("C4 3D lsrtm_objective shot nt=%d" % nt, wall_ms=wall * 1e3, ms_per_step_pair=wall * 1e3 / (nt - 2), f=float(out[0]))

def c5_tti(n=(301, 301, 201), nt=60):
    v = velocity(n)
    eps = (0.2 * (v - 1.5) / 4).astype(np.float32); dlt = (0.1 * (v - 1.5) / 4).astype(np.float32)
    th = (np.pi / 6 * (v - 1.5) / 4).astype(np.float32)
    gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32), epsilon=eps, delta=dlt, theta=th, phi=np.float32(np.pi / 8))
    dt = float(gm.critical_dt)
    q = ricker(nt, dt, 0.008)
    src = np.array([[3750., 3750., 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0., 7500., 76), np.linspace(0., 7500., 76), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
    N = float(np.prod(gm.shape_pml))
    wall, ms, nl, pts, out = timed(lambda: J.forward_rec(gm, src, q, rec), reps=1)
    rep("C5 3D TTI forward nt=%d dt=%.3f" % (nt, dt), wall_ms=wall * 1e3, ms_per_step=wall * 1e3 / (nt - 2), gpts=(nt - 2) * N / wall / 1e9, gbs_at_44B=44 * (nt - 2) * N / wall / 1e9)
    dobs = (0.9 * out[0]).astype(np.float32)
    wall, ms, nl, pts, out = timed(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4, snapshot_region="physical"), reps=1)
    rep("C5 3D TTI fwi gradient shot nt=%d" % nt, wall_ms=wall * 1e3, ms_per_step_pair=wall * 1e3 / (nt - 2), f=float(out[0]))

if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    if "c2" in which: c2_2d()
    if "c3" in which: c3_forward()
    if "c4" in which: c4_born()
    if "c5" in which: c5_tti()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(RES, open("gpurun_out/bench_configs.json", "w"), indent=1)
