"""Short FWI-gradient run on the C3 grid for ncu (few steps, all four launch flavours)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import judi_jl_b200 as J
n = (401, 401, 201); nt = int(sys.argv[1]) if len(sys.argv) > 1 else 24
region = sys.argv[2] if len(sys.argv) > 2 else "physical"
v = np.full(n, 2.0, dtype=np.float32); v[:, :, n[2] // 2:] = 3.0
gm = J.Model((0., 0., 0.), (25., 25., 25.), n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32))
dt = float(gm.critical_dt)
src = np.array([[5000., 5000., 50.]], dtype=np.float32)
xr, yr = np.meshgrid(np.linspace(0., 10000., 101), np.linspace(0., 10000., 101), indexing="ij")
rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
t = np.arange(nt) * dt
q = ((1 - 2 * (np.pi * 0.02 * (t - 20)) ** 2) * np.exp(-(np.pi * 0.02 * (t - 20)) ** 2)).astype(np.float32).reshape(nt, 1)
dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4, snapshot_region=region)
print("ok", float(f), float(np.abs(g).max()))
