"""Regression pins of the oracle against committed fixtures (tests/golden/make_golden.py).
These are self-generated (fp64 oracle): the reference has no golden vectors for this path and
cannot run here, so they guard against drift, not against Devito -- parity stays 'unpinned'."""
import os

import numpy as np
import pytest

from conftest import model_args, geometry, rel_l2
from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))
NT = 200


@pytest.mark.parametrize("precision,tol", [("f64", 1e-6), ("f32", 1e-4)])
@pytest.mark.parametrize("kind", ["iso", "rho", "tti"])
@pytest.mark.parametrize("ndim", [2, 3])
def test_oracle_reproduces_golden(kind, ndim, precision, tol):
    args = model_args(kind, ndim, 8) if ndim == 2 else model_args(kind, ndim, 8, n=(25, 23, 21), nbl=6)
    model = O.Model(precision=precision, **args)
    src, rec = geometry(args)
    dt = float(model.critical_dt)
    key = "%s%d_" % (kind, ndim)
    assert np.float32(dt) == G[key + "dt"]
    q = O.ricker_wavelet((NT - 1) * dt, dt, 0.02)[:NT]
    d, _ = O.forward_rec(model, src, q, rec)
    dl, _ = O.born_rec(model, src, q, rec)
    assert rel_l2(d[:, ::5], G[key + "d"]) < tol
    assert rel_l2(dl[:, ::5], G[key + "dl"]) < 10 * tol
    f, g, _, _ = O.J_adjoint(model, src, q, rec, 0.9 * d, return_obj=True)
    assert abs(f - G[key + "f"]) <= 100 * tol * abs(G[key + "f"])
    assert rel_l2(g[(slice(None, None, 4),) * ndim], G[key + "g_slice"]) < 100 * tol


def test_fast_path_is_bit_identical_to_generic_path():
    """The vectorised constant-density loops used for the CPU baseline are the same arithmetic."""
    for ndim in (2, 3):
        args = model_args("iso", ndim, 8, with_dm=False)
        model = O.Model(precision="f32", **args)
        src, rec = geometry(args)
        dt = float(model.critical_dt)
        q = O.ricker_wavelet(149 * dt, dt, 0.02)[:150]
        try:
            O.set_fast(1)
            d1, _ = O.forward_rec(model, src, q, rec)
            O.set_fast(0)
            d0, _ = O.forward_rec(model, src, q, rec)
        finally:
            O.set_fast(1)
        assert np.linalg.norm(d0) > 0 and np.array_equal(d0, d1)
