"""GPU-vs-oracle parity on the EXACT grids of BASELINE.json's configs:

  C1  2-D two-layer 401x201 @ 10 m, full record (T = 1000 ms): forward + adjoint + dot test
  C2  2-D overthrust-shaped 401x121 @ 25 m, one full shot (T = 3000 ms) of fwi_objective
  C3  3-D 401x401x201 @ 25 m (padded 481x481x281), 10 201 receivers: forward record and the
      t_sub = 4 gradient over the first time steps (the oracle runs this grid at a few Gpt/s)
  C5  3-D TTI 301x301x201 (padded 381x381x281): forward record and gradient, first time steps

Tolerances are the north-star ones: shot records 1e-4, gradients 1e-3 relative L2; dot tests are
reported against 1e-5 (see test_dot_products.py for the per-case account).
"""
import os

import numpy as np
import pytest

from conftest import rel_l2
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _threads():
    O.set_threads(os.cpu_count() or 1)


def test_c1_two_layer_forward_adjoint_dot():
    import judi_jl_b200 as J
    _threads()
    n, d = (401, 201), (10., 10.)
    v = np.full(n, 1.5, np.float32)
    v[:, 101:] = 2.5
    args = dict(origin=(0., 0.), spacing=d, shape=n, space_order=8, nbl=40,
                m=(1 / v ** 2).astype(np.float32))
    gm, om = J.Model(**args), O.Model(precision="f32", **args)
    o64 = O.Model(precision="f64", **args)
    assert gm.critical_dt == om.critical_dt
    dt = float(gm.critical_dt)
    q = O.ricker_wavelet(1000., dt, 0.010)
    nt = q.shape[0]
    src = np.array([[2000., 20.]], np.float32)
    rec = np.stack([np.arange(401, dtype=np.float32) * 10, np.full(401, 20., np.float32)], 1)
    dg, _ = J.forward_rec(gm, src, q, rec)
    do, _ = O.forward_rec(om, src, q, rec)
    d64, _ = O.forward_rec(o64, src, q, rec)
    e32, e64 = rel_l2(dg, do), rel_l2(dg, d64)
    y = np.random.default_rng(0).standard_normal(dg.shape).astype(np.float32)
    qa, _ = J.forward_rec(gm, rec, y, src, fw=False)
    qao, _ = O.forward_rec(om, rec, y, src, fw=False)
    a, b = np.vdot(dg.astype(np.float64), y), np.vdot(q.astype(np.float64), qa)
    dot = abs((a - b) / (a + b))
    ao, bo = np.vdot(do.astype(np.float64), y), np.vdot(q.astype(np.float64), qao)
    print("C1 nt=%d: shot vs f32 oracle %.2e, vs f64 oracle %.2e (f32 oracle vs f64 %.2e); adjoint "
          "%.2e; dot test %.2e (f32 oracle %.2e)" % (nt, e32, e64, rel_l2(do, d64), rel_l2(qa, qao),
                                                      dot, abs((ao - bo) / (ao + bo))))
    assert e32 < 1e-4 and e64 < 1e-4
    assert rel_l2(qa, qao) < 1e-4
    assert dot < 1e-5


def _c2_velocity():
    n, d = (401, 121), (25., 25.)
    x = (np.arange(n[0], dtype=np.float32) * d[0]).reshape(-1, 1)
    z = (np.arange(n[1], dtype=np.float32) * d[1]).reshape(1, -1)
    v = 1.5 + (z / 3000.) * 3.0 + 0.4 * np.sin(2 * np.pi * x / 5000.) * (z / 3000.)
    v = np.clip(v, 1.5, 5.5).astype(np.float32)
    v[:, :21] = 1.5
    from scipy.ndimage import gaussian_filter1d
    v0 = gaussian_filter1d(v, sigma=10, axis=1, mode="nearest").astype(np.float32)
    v0[:, :21] = 1.5
    return n, d, v, v0


def test_c2_small_overthrust_one_full_shot_of_fwi_objective():
    import judi_jl_b200 as J
    _threads()
    n, d, v, v0 = _c2_velocity()
    mk = lambda vv: dict(origin=(0., 0.), spacing=d, shape=n, space_order=8, nbl=40,
                         m=(1 / vv ** 2).astype(np.float32), dt=np.float32(2.353))
    gt, ot = J.Model(**mk(v)), O.Model(precision="f32", **mk(v))
    g0, o0 = J.Model(**mk(v0)), O.Model(precision="f32", **mk(v0))
    o64 = O.Model(precision="f64", **mk(v0))
    dt = float(g0.critical_dt)
    assert g0.critical_dt == o0.critical_dt
    q = O.ricker_wavelet(3000., dt, 0.008)
    nt = q.shape[0]
    src = np.array([[np.linspace(500., 9500., 16)[5], 50.]], np.float32)
    rec = np.stack([np.linspace(0., 10000., 401).astype(np.float32), np.full(401, 50., np.float32)], 1)
    dobs, _ = O.forward_rec(ot, src, q, rec)
    dg, _ = J.forward_rec(gt, src, q, rec)
    f, g, _, _ = J.J_adjoint(g0, src, q, rec, dobs, return_obj=True)
    fo, go, _, _ = O.J_adjoint(o0, src, q, rec, dobs, return_obj=True)
    f64, g64, _, _ = O.J_adjoint(o64, src, q, rec, dobs, return_obj=True)
    print("C2 nt=%d: shot %.2e; objective %.2e (vs f64 %.2e); gradient vs f32 oracle %.2e, vs f64 "
          "oracle %.2e (f32 oracle vs f64 %.2e)"
          % (nt, rel_l2(dg, dobs), abs(float(f) / float(fo) - 1), abs(float(f) / float(f64) - 1),
             rel_l2(g, go), rel_l2(g, g64), rel_l2(go, g64)))
    assert rel_l2(dg, dobs) < 1e-4
    assert abs(float(f) / float(fo) - 1) < 1e-4
    assert rel_l2(g, go) < 1e-3 and rel_l2(g, g64) < 1e-3


def _3d(shape, tti):
    import bench as B
    v = B.velocity(shape)
    v0 = B.smooth_model(v)
    kw = dict(origin=(0., 0., 0.), spacing=B.SPACING, shape=shape, space_order=8, nbl=B.NBL,
              m=(1 / v0 ** 2).astype(np.float32))
    if tti:
        p = B.tti_params(v0)
        kw.update(epsilon=np.ascontiguousarray(p["epsilon"]), delta=np.ascontiguousarray(p["delta"]),
                  theta=np.ascontiguousarray(p["theta"]),
                  phi=np.full(shape, p["phi"], np.float32))
    src, rec = B.geometry(shape, 27)
    return kw, src, rec


@pytest.mark.parametrize("name,shape,tti,nt", [("C3", (401, 401, 201), False, 45),
                                              ("C5", (301, 301, 201), True, 37)])
def test_c3_c5_grids_first_time_steps(name, shape, tti, nt):
    """The full BASELINE grid, every tile / z-chunk / halo configuration the bench runs, over the
    first time steps: shot record on all 10 201 receivers and the t_sub = 4 gradient."""
    import judi_jl_b200 as J
    _threads()
    kw, src, rec = _3d(shape, tti)
    gm, om = J.Model(**kw), O.Model(precision="f32", **kw)
    assert gm.critical_dt == om.critical_dt
    assert tuple(gm.shape_pml) == tuple(s + 80 for s in shape)
    dt = float(gm.critical_dt)
    # a 40 Hz wavelet so that the field has left the source cell within the sample
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.040)[:nt]
    dg, _ = J.forward_rec(gm, src, q, rec)
    do, _ = O.forward_rec(om, src, q, rec)
    assert np.linalg.norm(do) > 0
    e_d = rel_l2(dg, do)
    dobs = (0.7 * do).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4)
    fo, go, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, t_sub=4)
    e_g = rel_l2(g, go)
    _, gp, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4,
                              snapshot_region="physical")
    crop = tuple(slice(l, n_ - r) for (l, r), n_ in zip(gm.padsizes, g.shape))
    print("%s grid %s nt=%d: shot %.2e, objective %.2e, gradient %.2e"
          % (name, gm.shape_pml, nt, e_d, abs(float(f) / float(fo) - 1), e_g))
    assert e_d < 1e-4
    assert abs(float(f) / float(fo) - 1) < 1e-4
    assert e_g < 1e-3
    assert np.array_equal(gp[crop], g[crop])     # the bench's physical-domain snapshots
