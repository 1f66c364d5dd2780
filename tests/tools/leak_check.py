import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import judi_jl_b200 as J
from conftest import model_args, geometry
from oracle import oracle as O
free0 = None
for it in range(40):
    kind = ["iso", "rho", "tti", "visco"][it % 4]; nd = 2 + (it // 4) % 2
    args = model_args(kind, nd, 8)
    gm = J.Model(**args); src, rec = geometry(args)
    dt = float(gm.critical_dt); q = O.ricker_wavelet(99 * dt, dt, 0.02)[:100]
    d, _ = J.forward_rec(gm, src, q, rec)
    J.J_adjoint(gm, src, q, rec, 0.5 * d, return_obj=True, t_sub=2)
    J.born_rec(gm, src, q, rec)
    del gm
    f, t = torch.cuda.mem_get_info()
    if it == 7: free0 = f
    if it % 8 == 7: print(it, "free GB %.3f" % (f / 1e9), flush=True)
print("drift MB", (free0 - f) / 1e6)
