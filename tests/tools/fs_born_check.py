"""Free surface + Born with a perturbation that is non-zero on the surface row: library vs oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import judi_jl_b200 as J
from conftest import model_args, geometry, rel_l2
from oracle import oracle as O
O.build()
for kind, ndim in (("iso", 2), ("rho", 2), ("tti", 2), ("iso", 3)):
    for p2 in (1, 0):
        J.get_context(0).set_option("persist2d", p2)
        args = model_args(kind, ndim, 8, fs=True)
        args["dm"] = np.full(args["shape"], 0.05, dtype=np.float32)
        gm, om = J.Model(**args), O.Model(precision="f32", **args)
        src, rec = geometry(args)
        dt = float(gm.critical_dt); nt = 220
        q = O.ricker_wavelet((nt - 1) * dt, dt, 0.02)[:nt]
        for ic in ("as", "isic"):
            d, _ = J.born_rec(gm, src, q, rec, ic=ic)
            do, _ = O.born_rec(om, src, q, rec, ic=ic)
            print("%s %d-D persist2d=%d ic=%s: born_rec rel-L2 vs oracle %.2e" % (kind, ndim, p2, ic, rel_l2(d, do)), flush=True)
