"""BASELINE config C2: 2-D small-overthrust-shaped acoustic FWI gradient, 16 shots, so=8,
fwi_objective on one B200 (host arrays in, host gradient out)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import ricker
from scipy.ndimage import gaussian_filter1d
ctx = J.get_context(0)
n = (401, 121); d = (25., 25.)
x = np.arange(n[0])[:, None] * 25.; z = np.arange(n[1])[None, :] * 25.
v = np.clip(1.5 + (z / 3000.) * 3.0 + 0.4 * np.sin(2 * np.pi * x / 5000.) * (z / 3000.), 1.5, 5.5).astype(np.float32)
v[:, :21] = 1.5
v0 = gaussian_filter1d(v, sigma=10, axis=1, mode="nearest").astype(np.float32); v0[:, :21] = 1.5
kw = dict(space_order=8, nbl=40)
kind = os.environ.get("KIND", "iso")
if kind == "rho":
    kw["rho"] = (0.31 * (v * 1000.) ** 0.25).astype(np.float32)
if kind == "tti":
    kw.update(epsilon=(0.2 * (v - 1.5) / 4).astype(np.float32), delta=(0.1 * (v - 1.5) / 4).astype(np.float32),
              theta=(np.pi / 6 * (v - 1.5) / 4).astype(np.float32))
mt = J.Model((0., 0.), d, n, m=(1 / v ** 2).astype(np.float32), **kw)
dt = float(mt.critical_dt); nt = int(3000. / dt) + 1
q = ricker(nt, dt, 0.008)
rec = np.stack([np.linspace(0, 10000., 401), np.full(401, 50.)], axis=1).astype(np.float32)
shots = []
for sx in np.linspace(500., 9500., 16):
    src = np.array([[sx, 50.]], dtype=np.float32)
    dobs, _ = J.forward_rec(mt, src, q, rec)
    shots.append((src, q, rec, dobs))
m0 = np.asfortranarray((1 / v0 ** 2).astype(np.float32))
lsrtm = os.environ.get("LSRTM", "0") == "1"     # LS-RTM objective (Born forward sweeps) instead of FWI
dm0 = np.asfortranarray((1 / v ** 2 - 1 / v0 ** 2).astype(np.float32))
def run():
    gm = J.Model((0., 0.), d, n, m=m0, dt=dt, **kw)
    if lsrtm:
        return J.objectives.lsrtm_objective(gm, shots, dm0, nlind=True)
    return J.objectives.fwi_objective(gm, shots)
ctx.set_option("persist_blocks", int(os.environ.get("PERSIST_BLOCKS", "0")))
ctx.set_option("persist_vec", int(os.environ.get("PERSIST_VEC", "1")))
run(); ctx.synchronize()
reps = 5
t0 = time.perf_counter()
for _ in range(reps): f, g = run()
ctx.synchronize()
wall = (time.perf_counter() - t0) / reps
N = 481 * 201
print("C2 (%s) %s, 16 shots" % (kind, "lsrtm_objective" if lsrtm else "fwi_objective"))
print("C2 fwi_objective, 16 shots, nt=%d: %.1f ms per gradient (%.2f ms per shot, %.1f shots/s, %.1f Gpt/s)  f=%.4e |g|=%.4e" % (
    nt, 1e3 * wall, 1e3 * wall / 16, 16 / wall, 16 * 2 * nt * N / wall / 1e9, float(f), float(np.linalg.norm(g))), flush=True)
ctx.set_option("profile", 1)
run()
ctx.set_option("profile", 0)
# per-phase wall times of one shot
gm = J.Model((0., 0.), d, n, m=m0, dt=dt, **kw)
ctx.set_option("profile", 1)
J.J_adjoint(gm, *shots[3], return_obj=True)
ctx.set_option("profile", 0)
