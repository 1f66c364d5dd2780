"""Short forward run on the C3 grid for ncu (few steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
n = (401, 401, 201); nt = int(sys.argv[1]) if len(sys.argv) > 1 else 12
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = J.get_context(0); ctx.set_option("tile", tile)
v = np.full(n, 2.0, dtype=np.float32); v[:, :, n[2] // 2:] = 3.0
gm = J.Model((0., 0., 0.), (25., 25., 25.), n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32))
dt = float(gm.critical_dt)
src = np.array([[5000., 5000., 50.]], dtype=np.float32)
xr, yr = np.meshgrid(np.linspace(100., 9900., 51), np.linspace(100., 9900., 51), indexing="ij")
rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
t = np.arange(nt) * dt
q = ((1 - 2 * (np.pi * 0.008 * (t - 125)) ** 2) * np.exp(-(np.pi * 0.008 * (t - 125)) ** 2)).astype(np.float32).reshape(nt, 1)
d, _ = J.forward_rec(gm, src, q, rec)
print("ok", float(np.abs(d).max()))
