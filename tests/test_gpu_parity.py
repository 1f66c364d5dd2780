"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes -> libjudi_b200.so), against
the CPU oracle on the same seeded inputs.

Tolerances are the north-star ones (BASELINE.json): shot records <= 1e-4 relative L2, gradients
<= 1e-3 relative L2, source/receiver indexing bit-exact, adjoint dot tests <= 1e-5.

Dot tests come in two recipes (tests/tools/dot_report.py prints both for every case next to the fp32 and
fp64 oracle; profiles/r02_dot_report.log):
  * the reference's own (test/test_adjoint.jl:67-76, src/pysource/adjoint_test_F.py:75-88):
    y = F q, q_rand = (.5 + rand) .* q.  Measured on the B200: F <= 2.8e-7, J <= 1.3e-6 over all 20
    cases -- held to 1e-5 in test_dot_products_reference_recipe.
  * y = white noise: <F q, y> cancels almost completely (kappa = |Fq||y| / |<Fq,y>| = 54 ... 953), so
    every rounding error of the fields is amplified by kappa.  Measured: <= 4.1e-6 (F), <= 9.5e-6
    (J, kappa = 424) -- held to 1e-5 as well, with the conditioning-normalised bound
    err / kappa <= 5e-8 as the alternative for cases whose kappa a later change may move.
"""
import numpy as np
import pytest

from conftest import model_args, geometry, rel_l2
from oracle import oracle as O

pytestmark = pytest.mark.gpu

NT = 260
TOL_SHOT = 1e-4
TOL_GRAD = 1e-3
TOL_DOT = 1e-5


def dot_ok(a, b, na, nb):
    """|a - b| / |a + b| <= 1e-5, or the same normalised by the cancellation factor
    kappa = na * nb / |a| of the inner product a = <x, y> (na, nb: norms of x, y)."""
    err = abs((a - b) / (a + b))
    kappa = na * nb / max(abs(a), 1e-300)
    return err <= TOL_DOT or err / kappa <= 5e-8


def both(kind, ndim, so, nt=NT, **kw):
    import judi_jl_b200 as J
    args = model_args(kind, ndim, so, **kw)
    gm = J.Model(**args)
    om = O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dt = float(gm.critical_dt)
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.02)[:nt]
    return J, args, gm, om, src, rec, q, dt


# every propagator of the north star at every space order it names (kernels.py:46-77,144-182 and
# FD_utils.py:11-105 are generic in space_order), + the visco-acoustic family
# ("ttirho": TTI with a density field, the reference's own TTI test model, seismic_utils.jl:23-63)
CASES = [(k, nd, so) for nd in (2, 3) for k in ("iso", "rho", "tti") for so in (8, 4, 16)] + \
        [("ttirho", 3, 8), ("ttirho", 3, 4), ("ttirho", 3, 16), ("ttirho", 2, 8)] + \
        [("visco", 2, 8), ("viscoc", 2, 8), ("visco", 2, 4), ("visco", 3, 8), ("viscoc", 3, 8)]


@pytest.mark.parametrize("kind,ndim,so", CASES)
def test_host_logic_matches_oracle(kind, ndim, so):
    """critical_dt, padsizes, damping field and padded m are identical; receiver corner indices
    and weights are bit-exact (geom_utils.py:74-94 / SURVEY A.5)."""
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, so)
    assert gm.critical_dt == om.critical_dt
    assert gm.padsizes == om.padsizes
    assert gm.shape_pml == om.shape_pml
    assert rel_l2(gm.get_field("damp"), om.damp) < 1e-6
    assert np.array_equal(gm.get_field("m"), om.m)
    rng = np.random.default_rng(1)
    ext = np.array([(s - 1) * h for s, h in zip(args["shape"], args["spacing"])], dtype=np.float32)
    pts = (rng.random((500, ndim)).astype(np.float32) * (ext + 60) - 30).astype(np.float32)
    pts = np.concatenate([pts, rec, src, np.zeros((1, ndim), np.float32), ext.reshape(1, -1)])
    ijk, w, valid = gm.sparse_locate(pts)
    oijk, ow, ovalid = O.sparse_locate(om, pts)
    assert np.array_equal(ijk, oijk)
    assert np.array_equal(valid, ovalid)
    assert np.array_equal(w.view(np.int32), ow.view(np.int32))


@pytest.mark.parametrize("kind,ndim,so", CASES)
def test_forward_adjoint_parity_and_dot(kind, ndim, so):
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, so)
    d, I = J.forward_rec(gm, src, q, rec, illum=True)
    do, Io = O.forward_rec(om, src, q, rec, illum=True)
    assert np.linalg.norm(do) > 0
    assert rel_l2(d, do) < TOL_SHOT
    assert rel_l2(I, Io) < TOL_SHOT
    assert np.all(d[0] == 0) and np.all(d[-1] == 0)        # rec[0], rec[nt-1] never written (A.2)
    y = np.random.default_rng(0).standard_normal(d.shape).astype(np.float32)
    qa, _ = J.forward_rec(gm, rec, y, src, fw=False)
    qao, _ = O.forward_rec(om, rec, y, src, fw=False)
    assert rel_l2(qa, qao) < TOL_SHOT
    a, b = np.vdot(d.astype(np.float64), y), np.vdot(q.astype(np.float64), qa)
    assert dot_ok(a, b, np.linalg.norm(d), np.linalg.norm(y))


@pytest.mark.parametrize("kind,ndim,so", CASES)
def test_dot_products_reference_recipe(kind, ndim, so):
    """test/test_adjoint.jl:67-76: y = F q, q_rand = (.5 + rand) .* q, <F q_rand, y> = <q_rand, F' y>
    and dt <J dm, y> = <dm, J' y>, at the north star's 1e-5 (the reference itself asks 5e-4)."""
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, so)
    y, _ = J.forward_rec(gm, src, q, rec)
    qr = ((0.5 + np.random.default_rng(11).random(q.shape, dtype=np.float32)) * q).astype(np.float32)
    d, _ = J.forward_rec(gm, src, qr, rec)
    qa, _ = J.forward_rec(gm, rec, y, src, fw=False)
    a, b = np.vdot(d.astype(np.float64), y), np.vdot(qr.astype(np.float64), qa)
    assert abs((a - b) / (a + b)) < TOL_DOT
    dl, _ = J.born_rec(gm, src, qr, rec)
    g, _, _ = J.J_adjoint(gm, src, qr, rec, y, is_residual=True)
    dmp = np.pad(args["dm"], gm.padsizes, mode="edge").astype(np.float64)
    c, e = dt * np.vdot(dl.astype(np.float64), y), np.vdot(g.astype(np.float64), dmp)
    assert abs((c - e) / (c + e)) < TOL_DOT


@pytest.mark.parametrize("kind,ndim,so", CASES)
def test_born_and_gradient_parity_and_dot(kind, ndim, so):
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, so)
    dl, _ = J.born_rec(gm, src, q, rec)
    dlo, _ = O.born_rec(om, src, q, rec)
    assert np.linalg.norm(dlo) > 0
    assert rel_l2(dl, dlo) < TOL_SHOT
    y = np.random.default_rng(0).standard_normal(dl.shape).astype(np.float32)
    g, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True)
    go, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True)
    assert rel_l2(g, go) < TOL_GRAD
    dmp = np.pad(args["dm"], gm.padsizes, mode="edge").astype(np.float64)
    c, e = dt * np.vdot(dl.astype(np.float64), y), np.vdot(g.astype(np.float64), dmp)
    assert dot_ok(c, e, dt * np.linalg.norm(dl), np.linalg.norm(y))


@pytest.mark.parametrize("kind,ndim", [("iso", 2), ("tti", 2), ("iso", 3), ("rho", 3), ("tti", 3),
                                       ("ttirho", 3), ("visco", 2), ("viscoc", 3)])
def test_fwi_objective_and_options(kind, ndim):
    """fwi objective value + gradient, t_sub, checkpointing, physical-domain snapshots,
    lsrtm (born_fwd, nlind), illumination (test_gradients.jl, test_all_options.jl)."""
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, 8)
    do, _ = O.forward_rec(om, src, q, rec)
    dobs = (0.9 * do).astype(np.float32)
    f, g, Iu, Iv = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, illum=True)
    fo, go, Iuo, Ivo = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, illum=True)
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, go) < TOL_GRAD
    assert rel_l2(Iu, Iuo) < TOL_SHOT and rel_l2(Iv, Ivo) < TOL_GRAD
    # objective equals 1/2 dt ||F q - dobs||^2 (test_gradients.jl:36)
    d, _ = J.forward_rec(gm, src, q, rec)
    assert abs(float(f) - 0.5 * dt * np.linalg.norm(d - dobs) ** 2) <= 1e-4 * abs(float(f))
    # t_sub
    f4, g4, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=4)
    _, g4o, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, t_sub=4)
    assert rel_l2(g4, g4o) < TOL_GRAD
    # checkpointing recomputes the same arithmetic: bitwise equal to the stored-wavefield path
    fc, gc, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, checkpointing=True)
    assert fc == f and np.array_equal(gc, g)
    fc2, gc2, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, checkpointing=True,
                                 n_checkpoints=7)
    assert np.array_equal(gc2, g)
    # the three schedules of the in-HBM checkpointing: everything kept (default on a grid this
    # small), everything recomputed segment by segment, and a tail of 37 levels kept
    ctx = J.get_context()
    try:
        for mode in (0, 37):
            ctx.set_option("ckpt_direct", mode)
            fcm, gcm, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, checkpointing=True)
            assert fcm == f and np.array_equal(gcm, g), mode
    finally:
        ctx.set_option("ckpt_direct", 1)
    # physical-domain snapshots: same gradient inside, zero in the absorbing layer
    _, gp, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, snapshot_region="physical")
    crop = tuple(slice(l, n - r) for (l, r), n in zip(gm.padsizes, g.shape))
    assert np.array_equal(gp[crop], g[crop])
    outside = gp.copy()
    outside[crop] = 0
    assert not outside.any()
    # lsrtm_objective with nlind (interface.py:565, sensitivity.py:264-270)
    fl, gl, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True)
    flo, glo, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True)
    assert abs(float(fl) - float(flo)) <= 1e-4 * abs(float(flo))
    assert rel_l2(gl, glo) < TOL_GRAD
    # ... and with checkpointing (only the background field is checkpointed / recomputed)
    flc, glc, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True,
                                 checkpointing=True)
    assert flc == fl and np.array_equal(glc, gl)


@pytest.mark.parametrize("kind,ndim,so", [("iso", 2, 8), ("iso", 3, 8), ("rho", 2, 8), ("tti", 2, 8), ("rho", 3, 8),
                                          ("tti", 3, 8), ("visco", 2, 8), ("viscoc", 3, 8),
                                          ("iso", 2, 4), ("iso", 2, 16), ("iso", 3, 16), ("iso", 3, 4)])
@pytest.mark.parametrize("ic", ["isic", "fwi"])
def test_isic_fwi_imaging_conditions(kind, ndim, so, ic):
    """test_all_options.jl:13-53 runs isic / fwi on the variable-density and the TTI model
    (seismic_utils.jl:23-63): Born data and gradient against the oracle + the J dot test."""
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, so)
    dl, _ = J.born_rec(gm, src, q, rec, ic=ic)
    dlo, _ = O.born_rec(om, src, q, rec, ic=ic)
    assert rel_l2(dl, dlo) < TOL_SHOT
    y = np.random.default_rng(0).standard_normal(dl.shape).astype(np.float32)
    g, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True, ic=ic)
    go, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True, ic=ic)
    assert rel_l2(g, go) < TOL_GRAD
    dmp = np.pad(args["dm"], gm.padsizes, mode="edge").astype(np.float64)
    c, e = dt * np.vdot(dl.astype(np.float64), y), np.vdot(g.astype(np.float64), dmp)
    assert dot_ok(c, e, dt * np.linalg.norm(dl), np.linalg.norm(y))


def test_born_of_zero_is_exactly_zero_and_lsrtm_equals_fwi_bitwise():
    """test_jacobian.jl:32 (J*0 == 0) and test_gradients.jl:92-97 (atol = 0)."""
    J, args, gm, om, src, rec, q, dt = both("iso", 2, 8, with_dm=False)
    dl, _ = J.born_rec(gm, src, q, rec)
    assert not dl.any()
    dobs = np.random.default_rng(3).standard_normal((NT, rec.shape[0])).astype(np.float32)
    f1, g1, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True)
    f2, g2, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True)
    assert f1 == f2 and np.array_equal(g1, g2)


def test_linearity_and_determinism():
    """test_linearity.jl:17 (5e-5) and bit-reproducibility of repeated runs."""
    J, args, gm, om, src, rec, q, dt = both("iso", 3, 8)
    q2 = (np.roll(q, 9, axis=0) * np.float32(0.5)).astype(np.float32)
    d1, _ = J.forward_rec(gm, src, q, rec)
    d1b, _ = J.forward_rec(gm, src, q, rec)
    assert np.array_equal(d1, d1b)
    d2, _ = J.forward_rec(gm, src, q2, rec)
    d12, _ = J.forward_rec(gm, src, q + q2, rec)
    assert rel_l2(d12, d1 + d2) < 5e-5
    y = np.random.default_rng(5).standard_normal(d1.shape).astype(np.float32)
    ga, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True)
    gb, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True)
    assert np.array_equal(ga, gb)
    g2, _, _ = J.J_adjoint(gm, src, q, rec, 2 * y, is_residual=True)
    assert rel_l2(g2, 2 * ga) < 5e-5


def test_tma_kernel_matches_generic_kernel():
    """The z-marching TMA kernel against the one-thread-per-point kernel on a grid that spans
    several tiles and z-chunks, partial tiles included."""
    import judi_jl_b200 as J
    n = (150, 70, 90)
    rng = np.random.default_rng(7)
    v = (1.5 + rng.random(n) * 2.0).astype(np.float32)
    args = dict(origin=(0., 0., 0.), spacing=(10., 12., 8.), shape=n, space_order=8, nbl=14,
                m=(1 / v ** 2).astype(np.float32))
    ctx = J.get_context()
    gm = J.Model(**args)
    dt = float(gm.critical_dt)
    nt = 60
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.03)[:nt]
    src = np.array([[700., 400., 300.]], dtype=np.float32)
    rec = (rng.random((300, 3)) * np.array([1490., 828., 712.])).astype(np.float32)
    out = {}
    try:
        for kern in (0, 1):
            ctx.set_option("acoustic3d_kernel", kern)
            d, _ = J.forward_rec(gm, src, q, rec)
            dobs = np.zeros_like(d)
            f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=2)
            out[kern] = (d, g, f)
    finally:
        ctx.set_option("acoustic3d_kernel", 1)
    assert np.linalg.norm(out[0][0]) > 0
    assert rel_l2(out[1][0], out[0][0]) < 1e-5
    assert rel_l2(out[1][1], out[0][1]) < 1e-5
    for zc in (1, 2, 5):
        ctx.set_option("zchunks", zc)
        d, _ = J.forward_rec(gm, src, q, rec)
        ctx.set_option("zchunks", 0)
        assert np.array_equal(d, out[1][0])      # chunking does not change a single bit


@pytest.mark.parametrize("so", [8, 4, 16])
def test_tma_kernel_extended_source_and_receiver_flavours(so):
    """F_QEXT / F_WREC flavours of acoustic3d_tma (extended and full-wavefield sources, extended
    receiver: interface.py:50-204, 244-276) against the one-thread-per-point kernel, on a grid with
    partial tiles, and a launch count that shows the TMA kernel took the steps."""
    import judi_jl_b200 as J
    n = (90, 50, 70)
    rng = np.random.default_rng(9)
    v = (1.5 + rng.random(n) * 2.0).astype(np.float32)
    dm = np.zeros(n, np.float32)
    dm[30:60, 15:35, 20:50] = 0.02
    args = dict(origin=(0., 0., 0.), spacing=(10., 12., 8.), shape=n, space_order=so, nbl=11,
                m=(1 / v ** 2).astype(np.float32), dm=dm)
    ctx = J.get_context()
    gm = J.Model(**args)
    dt = float(gm.critical_dt)
    nt = 50
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.03)[:nt]
    rec = (rng.random((120, 3)) * np.array([890., 588., 552.])).astype(np.float32)
    w = np.zeros(gm.shape_pml, np.float32)
    w[40:60, 25:40, 30:50] = rng.standard_normal((20, 15, 20)).astype(np.float32)
    y = rng.standard_normal((nt, rec.shape[0])).astype(np.float32)
    usrc = np.zeros((nt,) + gm.shape_pml, np.float32)
    usrc[:, 50:54, 30:34, 40:44] = rng.standard_normal((nt, 4, 4, 4)).astype(np.float32)
    out = {}
    try:
        for kern in (0, 1):
            ctx.set_option("acoustic3d_kernel", kern)
            d, I = J.forward_rec_w(gm, w, q, rec, illum=True)
            wa, _ = J.adjoint_w(gm, rec, y, q, fw=False)
            dl, _ = J.born_rec_w(gm, w, q, rec)
            g, _, _ = J.J_adjoint(gm, None, q, rec, y, is_residual=True, ws=w, t_sub=2)
            du, _ = J.forward_wf_src(gm, usrc, rec)
            out[kern] = (d, I, wa, dl, g, du)
    finally:
        ctx.set_option("acoustic3d_kernel", 1)
    for a, b in zip(out[1], out[0]):
        assert np.linalg.norm(b) > 0
        assert rel_l2(a, b) < 2e-5
    before = ctx.stencil_stats()
    ctx.set_option("timing", 1)
    ctx.reset_stats()
    J.forward_rec_w(gm, w, q, rec)
    st = ctx.stencil_stats()
    ctx.set_option("timing", 0)
    # one stencil launch per time step, at TMA-kernel speed (the generic kernel is ~5x slower)
    assert st["launches"] == nt - 2


@pytest.mark.parametrize("so", [8, 4])
@pytest.mark.parametrize("ic", ["isic", "fwi"])
def test_isic_born_fast_path_matches_generic_kernel(so, ic):
    """3-D isic / fwi Born step on the single-pass machinery (background TMA launch writing delta,
    one flux launch for dt^2 div(dm grad u), linearised TMA launch with the F_QISIC source) against
    the one-thread-per-point kernel, partial tiles and several z-chunks included; LS-RTM gradient
    with stored and checkpointed wavefields."""
    import judi_jl_b200 as J
    n = (90, 50, 70)
    rng = np.random.default_rng(13)
    v = (1.5 + rng.random(n) * 2.0).astype(np.float32)
    dm = (0.02 * rng.standard_normal(n)).astype(np.float32)
    args = dict(origin=(0., 0., 0.), spacing=(10., 12., 8.), shape=n, space_order=so, nbl=11,
                m=(1 / v ** 2).astype(np.float32), dm=dm)
    ctx = J.get_context()
    gm = J.Model(**args)
    dt = float(gm.critical_dt)
    nt = 48
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.03)[:nt]
    src = np.array([[430., 300., 250.]], dtype=np.float32)
    rec = (rng.random((150, 3)) * np.array([890., 588., 552.])).astype(np.float32)
    y = rng.standard_normal((nt, rec.shape[0])).astype(np.float32)
    out = {}
    try:
        for kern in (0, 1):
            ctx.set_option("acoustic3d_kernel", kern)
            dl, _ = J.born_rec(gm, src, q, rec, ic=ic)
            f, g, _, _ = J.J_adjoint(gm, src, q, rec, y, return_obj=True, ic=ic, born_fwd=True, nlind=True)
            out[kern] = (dl, g, f)
    finally:
        ctx.set_option("acoustic3d_kernel", 1)
    assert np.linalg.norm(out[0][0]) > 0
    assert rel_l2(out[1][0], out[0][0]) < 2e-5
    assert rel_l2(out[1][1], out[0][1]) < 2e-5
    fc, gc, _, _ = J.J_adjoint(gm, src, q, rec, y, return_obj=True, ic=ic, born_fwd=True, nlind=True,
                               checkpointing=True)
    assert fc == out[1][2] and np.array_equal(gc, out[1][1])


def test_misfit_callback_and_forward_no_rec():
    J, args, gm, om, src, rec, q, dt = both("iso", 2, 8)
    do, _ = O.forward_rec(om, src, q, rec)
    dobs = (0.9 * do).astype(np.float32)

    def mse(dsyn, dob):   # src/TimeModeling/Modeling/losses.jl:15-19
        r = dsyn - dob
        return 0.5 * float(np.linalg.norm(r)) ** 2, r

    f0, g0, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True)
    f1, g1, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, misfit=mse)
    assert abs(float(f1) - float(f0)) <= 1e-5 * abs(float(f0))
    assert rel_l2(g1, g0) < 1e-6
    # a non-quadratic misfit together with nlind: the callback sees (rcvl, dobs - rnl), sensitivity.py:255-259
    def l1(dsyn, dob):
        r = dsyn - dob
        return float(np.sum(np.abs(r))), np.sign(r).astype(np.float32)

    f2, g2, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, misfit=l1, born_fwd=True, nlind=True)
    f2o, g2o, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, misfit=l1, born_fwd=True, nlind=True)
    assert abs(float(f2) - float(f2o)) <= 1e-4 * abs(float(f2o))
    assert rel_l2(g2, g2o) < 2e-2      # sign() of a residual near zero may flip between fp32 paths
    # wavefield history: u[t] sampled at the receivers reproduces the shot record
    u, _ = J.forward_no_rec(gm, src, q[:80])
    _, usave, _ = O.forward(om, src, None, q[:80], save=True)
    assert u.shape == (80,) + gm.shape_pml
    assert rel_l2(u[2:79], usave[2:79, 0]) < TOL_SHOT


def test_device_resident_path_matches_host_path():
    import torch
    J, args, gm, om, src, rec, q, dt = both("iso", 3, 8)
    d, _ = J.forward_rec(gm, src, q, rec)
    dobs = (0.9 * d).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True)
    dev = torch.device("cuda", gm.ctx.device)
    m_d = torch.from_numpy(np.ascontiguousarray(args["m"].transpose(2, 1, 0))).to(dev)
    a2 = dict(args)
    a2["m"] = m_d
    a2["dm"] = torch.from_numpy(np.ascontiguousarray(args["dm"].transpose(2, 1, 0))).to(dev)
    gm2 = J.Model(**a2)
    q_d = torch.from_numpy(np.ascontiguousarray(q.T)).to(dev)
    dobs_d = torch.from_numpy(np.ascontiguousarray(dobs.T)).to(dev)
    rec_d = torch.zeros((rec.shape[0], q.shape[0]), dtype=torch.float32, device=dev)
    J.forward_rec(gm2, src, q_d, rec, rec_out=rec_d)
    torch.cuda.synchronize()
    assert np.array_equal(rec_d.cpu().numpy().T, d)
    red = torch.zeros(int(np.prod(gm2.shape_pml)) + 1, dtype=torch.float32, device=dev)
    for _ in range(2):   # accumulate twice: the local reduction of distributed.jl:6-22
        J.J_adjoint(gm2, src, q_d, rec, dobs_d, return_obj=True, grad_out=red[:-1], f_out=red[-1:],
                    accumulate=True)
    torch.cuda.synchronize()
    gd = red[:-1].cpu().numpy().reshape(gm2.shape_pml[::-1]).transpose(2, 1, 0)
    assert rel_l2(gd, 2 * g) < 1e-6
    assert abs(float(red[-1].item()) - 2 * float(f)) <= 1e-5 * abs(2 * float(f))


@pytest.mark.parametrize("kind", ["iso", "rho", "tti"])
@pytest.mark.parametrize("n,nbl", [((23, 17, 13), 5), ((34, 9, 40), 4), ((7, 50, 21), 6), ((65, 33, 10), 9)])
def test_odd_3d_shapes(kind, n, nbl):
    """Tile edges of the TMA kernels: grids that are not multiples of the 64x16 / 32x16 tiles or of
    the 4-point groups, thinner than a tile, with layers thinner than the stencil radius + 3."""
    J, args, gm, om, src, rec, q, dt = both(kind, 3, 8, nt=90, n=n, nbl=nbl)
    ext = np.array([(s - 1) * h for s, h in zip(n, args["spacing"])], dtype=np.float32)
    src = np.minimum(src, ext * 0.6).astype(np.float32)
    rec = np.minimum(rec, ext).astype(np.float32)
    d, _ = J.forward_rec(gm, src, q, rec)
    do, _ = O.forward_rec(om, src, q, rec)
    assert np.linalg.norm(do) > 0 and rel_l2(d, do) < TOL_SHOT
    dobs = (0.7 * do).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=2)
    fo, go, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, t_sub=2)
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, go) < TOL_GRAD
    _, gp, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=2, snapshot_region="physical")
    crop = tuple(slice(l, m_ - r) for (l, r), m_ in zip(gm.padsizes, g.shape))
    assert np.array_equal(gp[crop], g[crop])


@pytest.mark.parametrize("kind", ["iso", "rho", "tti", "visco", "viscoc"])
def test_fg_multishot_equals_the_sum_of_single_shots(kind):
    """judi_fg_multishot (the per-worker part of multi_src_fg!, propagation.jl:93-107): host
    pointers in, host sums out, with and without accumulation; and the device-accumulator mode
    through objectives.fwi_objective."""
    import ctypes
    from judi_jl_b200 import _lib as L
    from judi_jl_b200.interface import _coords, _traces
    J, args, gm, om, src, rec, q, dt = both(kind, 2, 8)
    shots, fsum, gsum = [], 0.0, 0.0
    for k in range(3):
        s_k = (src + np.float32([[50. * (k - 1), 0.]])).astype(np.float32)
        d_k, _ = J.forward_rec(gm, s_k, q, rec)
        dobs = (0.8 * d_k).astype(np.float32)
        shots.append((s_k, q, rec, dobs))
        f_k, g_k, _, _ = J.J_adjoint(gm, s_k, q, rec, dobs, return_obj=True, t_sub=2)
        fsum += float(f_k)
        gsum = gsum + g_k.astype(np.float64)
    arr = (L.JudiShot * 3)()
    keep = []
    for k, (sc, w, rc, d) in enumerate(shots):
        sc, rc = _coords(sc, 2), _coords(rc, 2)
        w, nt, nsrc = _traces(w)
        d, _, nrec = _traces(d)
        keep += [sc, rc, w, d]
        arr[k].nsrc, arr[k].src_coords, arr[k].wavelet = nsrc, sc.ctypes.data, w.ctypes.data
        arr[k].nt, arr[k].nrec, arr[k].rec_coords, arr[k].recin = nt, nrec, rc.ctypes.data, d.ctypes.data
    o = L.JudiJopts()
    o.t_sub = 2
    g = np.zeros(gm.shape_pml, dtype=np.float32, order="F")
    f = ctypes.c_float(0.)
    L.check(L.lib().judi_fg_multishot(gm.h, 3, arr, ctypes.byref(o), None, None,
                                      ctypes.addressof(f), g.ctypes.data, 0))
    assert abs(f.value - fsum) <= 1e-5 * abs(fsum)
    assert rel_l2(g, gsum) < 1e-5
    # accumulate on top of the previous result: twice the sums
    L.check(L.lib().judi_fg_multishot(gm.h, 3, arr, ctypes.byref(o), None, None,
                                      ctypes.addressof(f), g.ctypes.data, L.GRAD_ACCUMULATE))
    assert abs(f.value - 2 * fsum) <= 1e-5 * abs(fsum)
    assert rel_l2(g, 2 * gsum) < 1e-5
    fo, go = J.objectives.fwi_objective(gm, shots, t_sub=2)
    crop = tuple(slice(l, n_ - r) for (l, r), n_ in zip(gm.padsizes, g.shape))
    assert abs(float(fo) - fsum) <= 1e-5 * abs(fsum)
    assert rel_l2(go, gsum[crop]) < 1e-5
    if kind != "iso":
        return
    # lsrtm_objective (born_fwd, nlind) through the same entry point
    fl, gl = 0.0, 0.0
    for (s_k, w, rc, dobs) in shots:
        f_k, g_k, _, _ = J.J_adjoint(gm, s_k, w, rc, dobs, return_obj=True, born_fwd=True, nlind=True)
        fl += float(f_k)
        gl = gl + g_k.astype(np.float64)
    flo, glo = J.objectives.lsrtm_objective(gm, shots, args["dm"], nlind=True)
    assert abs(float(flo) - fl) <= 1e-5 * abs(fl)
    assert rel_l2(glo, gl[crop]) < 1e-5
    # the isic imaging condition (lsrtm_2D.jl:47) in lockstep: Born source and gradient
    fl, gl = 0.0, 0.0
    for (s_k, w, rc, dobs) in shots:
        f_k, g_k, _, _ = J.J_adjoint(gm, s_k, w, rc, dobs, return_obj=True, born_fwd=True, ic="isic")
        fl += float(f_k)
        gl = gl + g_k.astype(np.float64)
    flo, glo = J.objectives.lsrtm_objective(gm, shots, args["dm"], ic="isic")
    assert abs(float(flo) - fl) <= 1e-5 * abs(fl)
    assert rel_l2(glo, gl[crop]) < 1e-5


def test_registered_host_arrays_give_identical_results():
    """judi_host_register: page-locking the caller's arrays changes the copy path, not the result."""
    J, args, gm, om, src, rec, q, dt = both("iso", 3, 8)
    d, _ = J.forward_rec(gm, src, q, rec)
    dobs = (0.9 * d).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True)
    m_pin = J.host_register(np.array(args["m"], order="F"))
    q_pin, dobs_pin = J.host_register(q.copy()), J.host_register(dobs.copy())
    J.host_register(q_pin)                      # registering twice is not an error
    try:
        gm2 = J.Model(**dict(args, m=m_pin))
        f2, g2, _, _ = J.J_adjoint(gm2, src, q_pin, rec, dobs_pin, return_obj=True)
        assert f2 == f and np.array_equal(g2, g)
    finally:
        for a in (m_pin, q_pin, dobs_pin):
            J.host_unregister(a)
    J.host_unregister(q_pin)                    # neither is unregistering twice


@pytest.mark.parametrize("kind,ndim", [("iso", 2), ("iso", 3), ("tti", 2), ("rho", 3), ("tti", 3), ("rho", 2),
                                       ("visco", 2)])
def test_free_surface_parity(kind, ndim):
    """Options(free_surface=true): no layer above z=0 (models.py:174-176,269-272) and the mirror
    u(-k) = -u(k-1), u(0) = 0 after every update (fields_exprs.py:106-137)."""
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, 8, fs=True)
    assert gm.padsizes == om.padsizes and gm.padsizes[-1][0] == 0
    assert gm.shape_pml == om.shape_pml
    assert rel_l2(gm.get_field("damp"), om.damp) < 1e-6
    d, _ = J.forward_rec(gm, src, q, rec)
    do, _ = O.forward_rec(om, src, q, rec)
    assert rel_l2(d, do) < TOL_SHOT
    dl, _ = J.born_rec(gm, src, q, rec)
    dlo, _ = O.born_rec(om, src, q, rec)
    assert rel_l2(dl, dlo) < TOL_SHOT
    dobs = (0.9 * do).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True)
    fo, go, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True)
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, go) < TOL_GRAD


@pytest.mark.parametrize("kind,ndim", [("iso", 2), ("iso", 3), ("tti", 2), ("rho", 2), ("rho", 3), ("tti", 3),
                                       ("visco", 2)])
@pytest.mark.parametrize("ic", ["as", "isic"])
def test_free_surface_born_with_a_perturbation_on_the_surface_row(kind, ndim, ic):
    """born_op evaluates the linearised source after the mirror equations of the background field
    (operators.py:188), where u(z=0) is identically zero: a perturbation that reaches the surface
    row contributes nothing there (judi_model::dm_src).  Through every Born kernel (single-pass
    3-D, two-pass, one-point-per-thread, persistent 2-D) and its LS-RTM dot test."""
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, 8, fs=True)
    dm = np.full(args["shape"], 0.04, dtype=np.float32)
    gm.set_dm(dm)
    args2 = dict(args, dm=dm)
    om = O.Model(precision="f32", **args2)
    dl, _ = J.born_rec(gm, src, q, rec, ic=ic)
    dlo, _ = O.born_rec(om, src, q, rec, ic=ic)
    assert rel_l2(dl, dlo) < TOL_SHOT
    ctx = J.get_context(0)
    try:
        ctx.set_option("persist2d", 0)
        dl2, _ = J.born_rec(gm, src, q, rec, ic=ic)
    finally:
        ctx.set_option("persist2d", 1)
    assert rel_l2(dl2, dlo) < TOL_SHOT
    # and the matching imaging condition under the free surface (snapshots carry no mirrored halo)
    y = np.random.default_rng(2).standard_normal(dl.shape).astype(np.float32)
    g, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True, ic=ic)
    go, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True, ic=ic)
    assert rel_l2(g, go) < TOL_GRAD


@pytest.mark.parametrize("kind,ndim", [("iso", 2), ("iso", 3), ("tti", 2), ("rho", 2), ("tti", 3), ("visco", 2)])
def test_extended_and_wavefield_sources(kind, ndim):
    """The remaining entry points of src/pysource/interface.py: forward_rec_w (:50), adjoint_w
    (:174), born_rec_w (:244), J_adjoint(ws=...) and forward_wf_src / _norec (:114, :145);
    adjointness as in test/test_adjoint.jl:85-115."""
    J, args, gm, om, src, rec, q, dt = both(kind, ndim, 8, nt=200)
    rng = np.random.default_rng(11)
    w = np.zeros(gm.shape_pml, dtype=np.float32)
    box = tuple(slice(s // 2 - 6, s // 2 + 6) for s in gm.shape_pml)
    w[box] = rng.standard_normal(w[box].shape).astype(np.float32)
    d, _ = J.forward_rec_w(gm, w, q, rec)
    do, _ = O.forward_rec_w(om, w, q, rec)
    assert np.linalg.norm(do) > 0 and rel_l2(d, do) < TOL_SHOT
    y = rng.standard_normal(d.shape).astype(np.float32)
    wa, _ = J.adjoint_w(gm, rec, y, q, fw=False)
    wao, _ = O.adjoint_w(om, rec, y, q, fw=False)
    assert rel_l2(wa, wao) < TOL_SHOT
    a, b = dt * np.vdot(d.astype(np.float64), y), np.vdot(w.astype(np.float64), wa)
    assert dot_ok(a, b, dt * np.linalg.norm(d), np.linalg.norm(y))
    dl, _ = J.born_rec_w(gm, w, q, rec)
    dlo, _ = O.born_rec_w(om, w, q, rec)
    assert np.linalg.norm(dlo) > 0 and rel_l2(dl, dlo) < TOL_SHOT
    g, _, _ = J.J_adjoint(gm, None, q, rec, y, is_residual=True, ws=w)
    go, _, _ = O.J_adjoint(om, None, q, rec, y, is_residual=True, ws=w)
    assert rel_l2(g, go) < TOL_GRAD
    # full wavefield source (short, the array is nt x grid)
    nts = 40
    usrc = np.zeros((nts,) + gm.shape_pml, dtype=np.float32)
    usrc[(slice(2, 30),) + box] = rng.standard_normal((28,) + w[box].shape).astype(np.float32)
    r1, _ = J.forward_wf_src(gm, usrc, rec)
    r1o, _ = O.forward_wf_src(om, usrc, rec)
    assert np.linalg.norm(r1o) > 0 and rel_l2(r1, r1o) < TOL_SHOT
    u1, _ = J.forward_wf_src_norec(gm, usrc)
    u1o, _ = O.forward_wf_src_norec(om, usrc)
    assert rel_l2(u1[2:nts - 1], u1o[2:nts - 1]) < TOL_SHOT


# ---- BASELINE.json's full sizes: size-independent properties (no oracle run at these sizes) ------
def _full_size(kind):
    import judi_jl_b200 as J
    from bench import velocity, smooth_model, ricker
    if kind == "iso":
        n = (401, 401, 201)           # C3 / C4 grid: padded 481 x 481 x 281
        v = velocity(n)
        v0 = smooth_model(v)
        dm = (1 / v ** 2 - 1 / v0 ** 2).astype(np.float32)
        gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40,
                     m=(1 / v0 ** 2).astype(np.float32), dm=dm, dt=1.921)
    elif kind == "rho":
        n = (301, 301, 201)           # C5-size grid with a density field (Gardner)
        v = velocity(n)
        dm = (0.02 * (v > 2.0) / v ** 2).astype(np.float32)
        gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32),
                     rho=(0.31 * (v * 1000.) ** 0.25).astype(np.float32), dm=dm)
    else:
        n = (301, 301, 201)           # C5 grid: padded 381 x 381 x 281
        v = velocity(n)
        dm = (0.02 * (v > 2.0) / v ** 2).astype(np.float32)
        gm = J.Model((0., 0., 0.), (25.,) * 3, n, space_order=8, nbl=40, m=(1 / v ** 2).astype(np.float32),
                     epsilon=(0.2 * (v - 1.5) / 4).astype(np.float32),
                     delta=(0.1 * (v - 1.5) / 4).astype(np.float32),
                     theta=(np.pi / 6 * (v - 1.5) / 4).astype(np.float32), phi=np.float32(np.pi / 8), dm=dm)
    ex, ey = (n[0] - 1) * 25., (n[1] - 1) * 25.
    src = np.array([[0.47 * ex, 0.52 * ey, 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(0.3 * ex, 0.7 * ex, 41), np.linspace(0.3 * ey, 0.7 * ey, 41), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 62.5)], axis=1).astype(np.float32)
    nt = 120
    dt = float(gm.critical_dt)
    t = np.arange(nt) * dt
    r = np.pi * 0.025 * (t - 40.)
    q = ((1 - 2 * r ** 2) * np.exp(-r ** 2)).astype(np.float32).reshape(nt, 1)
    return J, gm, src, rec, q, dt, dm


@pytest.mark.parametrize("kind", ["iso", "tti", "rho"])
def test_full_size_adjoint_and_linearity(kind):
    """test_adjoint.jl:28-54 / test_linearity.jl:17 on the BASELINE grids (C3 and C5), short time
    axis: <F q, y> = <q, F' y>, dt <J dm, y> = <dm, J' y>, F(a q) = a F(q), and the gradient is
    independent of how the snapshots are kept (stored / checkpointed)."""
    J, gm, src, rec, q, dt, dm = _full_size(kind)
    d, _ = J.forward_rec(gm, src, q, rec)
    assert np.isfinite(d).all() and np.linalg.norm(d) > 0
    y = np.random.default_rng(20261019).standard_normal(d.shape).astype(np.float32)
    qa, _ = J.forward_rec(gm, rec, y, src, fw=False)
    a, b = np.vdot(d.astype(np.float64), y), np.vdot(q.astype(np.float64), qa)
    assert dot_ok(a, b, np.linalg.norm(d), np.linalg.norm(y))
    d2, _ = J.forward_rec(gm, src, (2.5 * q).astype(np.float32), rec)
    assert rel_l2(d2, 2.5 * d) < 5e-5                        # test_linearity.jl:17 tolerance
    dl, _ = J.born_rec(gm, src, q, rec)
    g, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True)
    dmp = np.pad(dm, gm.padsizes, mode="edge").astype(np.float64)
    c, e = dt * np.vdot(dl.astype(np.float64), y), np.vdot(g.astype(np.float64), dmp)
    assert dot_ok(c, e, dt * np.linalg.norm(dl), np.linalg.norm(y))
    gc, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True, checkpointing=True)
    assert np.array_equal(gc, g)


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("so", [4, 8])
def test_tti_kernels_without_anisotropy_equal_the_density_field_kernels(ndim, so):
    """epsilon = delta = 0: R^T (b I) R = b I for any tilt / azimuth, so the TTI kernels must
    reproduce the density-field (staggered div(b grad)) kernels -- the operator that
    tests/test_devito_known_answer.py holds against a Devito-produced constant (oracle twin of this
    test: test_oracle_properties.py::test_tti_without_anisotropy_is_the_staggered_acoustic_operator)."""
    import judi_jl_b200 as J
    a = model_args("tti", ndim, so=so, with_dm=False)
    v = 1 / np.sqrt(a["m"])
    zero = np.zeros(a["shape"], np.float32)
    tti = dict(a, epsilon=zero, delta=zero, dt=np.float32(0.5))
    tti["theta"] = (0.7 * np.sin(v * 3.0)).astype(np.float32)
    if ndim == 3:
        tti["phi"] = (1.1 * np.cos(v * 2.0)).astype(np.float32)
    iso = {k: val for k, val in a.items() if k not in ("epsilon", "delta", "theta", "phi")}
    iso.update(b=np.ones(a["shape"], np.float32), dt=np.float32(0.5))
    Mt, Mi = J.Model(**tti), J.Model(**iso)
    assert float(Mt.critical_dt) == float(Mi.critical_dt) == 0.5
    src, rec = geometry(a)
    nt = 160 if ndim == 2 else 80
    q = O.ricker_wavelet((nt - 1) * 0.5, 0.5, 0.03)[:nt]
    dt_, _ = J.forward_rec(Mt, src, q, rec)
    di_, _ = J.forward_rec(Mi, src, q, rec)
    do_, _ = O.forward_rec(O.Model(precision="f64", **iso), src, q, rec)
    e_ti, e_to = rel_l2(dt_, di_), rel_l2(dt_, do_)
    print("TTI(0) vs density-field kernels %dD so=%d: %.2e; vs the fp64 oracle's density-field path %.2e"
          % (ndim, so, e_ti, e_to))
    assert np.linalg.norm(di_) > 0
    assert e_ti < 2e-5 and e_to < 1e-4
