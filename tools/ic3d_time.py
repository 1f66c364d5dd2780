"""C3 grid (481x481x281 padded): wall time of an FWI / LS-RTM gradient per imaging condition and of the
on-the-fly DFT gradient, short time axis (nt = 64), t_sub = 4.  Shows what the isic / fwi / DFT paths
cost relative to the fused "as" path of the single-pass kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
n = (401, 401, 201); nt = 64
ctx = J.get_context(0)
v = np.full(n, 2.0, dtype=np.float32); v[:, :, n[2] // 2:] = 3.0
dm = np.zeros(n, dtype=np.float32); dm[:, :, n[2] // 2] = 0.01
gm = J.Model((0., 0., 0.), (25., 25., 25.), n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32), dm=dm)
dt = float(gm.critical_dt)
src = np.array([[5000., 5000., 50.]], dtype=np.float32)
xr, yr = np.meshgrid(np.linspace(0., 10000., 101), np.linspace(0., 10000., 101), indexing="ij")
rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
t = np.arange(nt) * dt
q = ((1 - 2 * (np.pi * 0.02 * (t - 20)) ** 2) * np.exp(-(np.pi * 0.02 * (t - 20)) ** 2)).astype(np.float32).reshape(nt, 1)
dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
def wall(fn, reps=2):
    fn(); ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.synchronize(); return 1e3 * (time.perf_counter() - t0) / reps
for ic in ("as", "isic", "fwi"):
    a = wall(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4, ic=ic))
    b = wall(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4, ic=ic, born_fwd=True))
    print("ic=%s nt=%d t_sub=4: fwi gradient %.1f ms (%.3f ms per step pair)  lsrtm gradient %.1f ms (%.3f)" % (
        ic, nt, a, a / (nt - 2), b, b / (nt - 2)), flush=True)
for ic in ("as", "isic"):
    for sub in (1, 4):
        a = wall(lambda: J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, ic=ic, freq_list=[0.005, 0.008, 0.012, 0.016], dft_sub=sub))
        print("DFT (4 frequencies, dft_sub=%d) ic=%s: gradient %.1f ms (%.3f ms per step pair)" % (sub, ic, a, a / (nt - 2)), flush=True)
