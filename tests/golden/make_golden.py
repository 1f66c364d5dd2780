"""Generates tests/golden/oracle_golden.npz.

The reference cannot be executed in the build container (no Devito, no Julia) and ships no
golden vectors for this path, so these fixtures are produced by the fp64 CPU
oracle itself: they pin the oracle (and, through the GPU parity tests, the CUDA path) against
accidental change; they do NOT pin it against Devito ("parity unpinned").

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import model_args, geometry  # noqa: E402
from oracle import oracle as O  # noqa: E402

NT = 200


def run(kind, ndim, so=8):
    args = model_args(kind, ndim, so)
    if ndim == 3:
        args = model_args(kind, ndim, so, n=(25, 23, 21), nbl=6)
    model = O.Model(precision="f64", **args)
    src, rec = geometry(args)
    dt = float(model.critical_dt)
    q = O.ricker_wavelet((NT - 1) * dt, dt, 0.02)[:NT]
    d, _ = O.forward_rec(model, src, q, rec)
    dl, _ = O.born_rec(model, src, q, rec)
    dobs = 0.9 * d
    f, g, _, _ = O.J_adjoint(model, src, q, rec, dobs, return_obj=True)
    sel = slice(None, None, 5)
    return {"d": d[:, sel].astype(np.float32), "dl": dl[:, sel].astype(np.float32),
            "f": np.float64(f), "g_norm": np.float64(np.linalg.norm(g)),
            "g_slice": g[(slice(None, None, 4),) * ndim].astype(np.float32),
            "dt": np.float32(dt)}


if __name__ == "__main__":
    out = {}
    for kind in ("iso", "rho", "tti"):
        for ndim in (2, 3):
            for k, v in run(kind, ndim).items():
                out["%s%d_%s" % (kind, ndim, k)] = v
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", len(out), "arrays")
