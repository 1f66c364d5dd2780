"""2-D extended-source modelling (forward_rec_w / adjoint_w / J_adjoint(ws=...)) on the C2 grid:
persistent kernel (PERSIST2D=1) against launch-per-step (PERSIST2D=0)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
from bench import ricker
ctx = J.get_context(0)
ctx.set_option("persist2d", int(os.environ.get("PERSIST2D", "1")))
n = (401, 121); d = (25., 25.)
z = np.arange(n[1])[None, :] * 25.
v = np.clip(1.5 + (z / 3000.) * 3.0, 1.5, 5.5).astype(np.float32) * np.ones(n, dtype=np.float32)
gm = J.Model((0., 0.), d, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32))
dt = float(gm.critical_dt); nt = int(3000. / dt) + 1
q = ricker(nt, dt, 0.008)
rec = np.stack([np.linspace(0, 10000., 401), np.full(401, 50.)], axis=1).astype(np.float32)
w = np.zeros(gm.shape_pml, dtype=np.float32); w[200:280, 60:90] = np.random.default_rng(0).standard_normal((80, 30)).astype(np.float32)
def wall(fn, reps=3):
    fn(); ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.synchronize(); return 1e3 * (time.perf_counter() - t0) / reps
d0, _ = J.forward_rec_w(gm, w, q, rec)
print("persist2d=%s nt=%d: forward_rec_w %.2f ms  adjoint_w %.2f ms  J_adjoint(ws) %.2f ms  |d|=%.4e" % (
    os.environ.get("PERSIST2D", "1"), nt,
    wall(lambda: J.forward_rec_w(gm, w, q, rec)), wall(lambda: J.adjoint_w(gm, rec, d0, q)),
    wall(lambda: J.J_adjoint(gm, None, q, rec, d0, ws=w, is_residual=True)), float(np.linalg.norm(d0))), flush=True)
