"""Tiny run through every kernel family (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import judi_jl_b200 as J
from conftest import model_args, geometry
nt = 24
def ricker(nt, dt, f0):
    t = np.arange(nt, dtype=np.float32) * np.float32(dt); r = np.pi * f0 * (t - 10 * dt)
    return ((1 - 2 * r ** 2) * np.exp(-r ** 2)).astype(np.float32).reshape(nt, 1)
for kind, nd, kw in [("iso", 2, {}), ("tti", 2, {}), ("rho", 2, {}), ("iso", 3, dict(n=(37, 33, 29), nbl=6)), ("rho", 3, dict(n=(21, 19, 17), nbl=5)),
                     ("tti", 3, dict(n=(21, 19, 17), nbl=5)), ("iso", 3, dict(n=(21, 19, 17), nbl=5, fs=True)), ("iso", 2, dict(fs=True))]:
    args = model_args(kind, nd, 8, **kw)
    gm = J.Model(**args); src, rec = geometry(args)
    ext = [(s - 1) * h for s, h in zip(args["shape"], args["spacing"])]
    rec = np.clip(rec, 0, np.array(ext, dtype=np.float32))
    src = np.clip(src, 0, np.array(ext, dtype=np.float32))
    dt = float(gm.critical_dt); q = ricker(nt, dt, 0.03)
    d, I = J.forward_rec(gm, src, q, rec, illum=True)
    qa, _ = J.forward_rec(gm, rec, d, src, fw=False)
    dl, _ = J.born_rec(gm, src, q, rec)
    f, g, Iu, Iv = J.J_adjoint(gm, src, q, rec, 0.5 * d, return_obj=True, illum=True, t_sub=2)
    f2, g2, _, _ = J.J_adjoint(gm, src, q, rec, 0.5 * d, return_obj=True, checkpointing=True, n_checkpoints=3)
    f3, g3, _, _ = J.J_adjoint(gm, src, q, rec, 0.5 * d, return_obj=True, born_fwd=True, nlind=True, snapshot_region="physical")
    w = np.zeros(gm.shape_pml, dtype=np.float32); w[tuple(slice(s // 2 - 2, s // 2 + 2) for s in gm.shape_pml)] = 1
    dw, _ = J.forward_rec_w(gm, w, q, rec); wa, _ = J.adjoint_w(gm, rec, dw, q, fw=False)
    if kind == "iso" and not kw.get("fs"):
        gi, _, _ = J.J_adjoint(gm, src, q, rec, d, is_residual=True, ic="isic")
        di, _ = J.born_rec(gm, src, q, rec, ic="fwi")
    u, _ = J.forward_no_rec(gm, src, q[:8])
    print(kind, nd, kw, "ok", float(np.abs(d).max()), float(f), float(np.abs(g).max()), flush=True)
print("ALL OK")
