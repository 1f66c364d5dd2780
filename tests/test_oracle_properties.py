"""The reference pins this path only through properties: adjoint dot tests
(test/test_adjoint.jl:28-54, src/pysource/adjoint_test_F.py, adjoint_test_J.py), J*0 == 0 and
Taylor rates (test/test_jacobian.jl:17-43, test/testing_utils.jl:3-26), objective value and
lsrtm(dm=0,nlind) == fwi (test/test_gradients.jl:31-40,82-99), linearity
(test/test_linearity.jl), option paths (test/test_all_options.jl).  These tests hold the CPU
oracle to the same properties -- in fp64 far tighter than the reference's 5e-4."""
import numpy as np
import pytest

from conftest import model_args, geometry, rel_l2
from oracle import oracle as O

NT = 260


def setup(kind, ndim=2, so=8, precision="f64", **kw):
    args = model_args(kind, ndim, so, **kw)
    model = O.Model(precision=precision, **args)
    src, rec = geometry(args)
    dt = float(model.critical_dt)
    q = O.ricker_wavelet((NT - 1) * dt, dt, 0.02)[:NT]
    return args, model, src, rec, q, dt


def dots(model, src, rec, q, dt, ic="as", t_sub=1, checkpointing=False):
    d, _ = O.forward_rec(model, src, q, rec)
    y = np.random.default_rng(0).standard_normal(d.shape).astype(d.dtype)
    qa, _ = O.forward_rec(model, rec, y, src, fw=False)
    a, b = np.vdot(d, y), np.vdot(q.astype(d.dtype), qa)
    dl, _ = O.born_rec(model, src, q, rec, ic=ic)
    g, _, _ = O.J_adjoint(model, src, q, rec, y, is_residual=True, ic=ic, t_sub=t_sub,
                          checkpointing=checkpointing)
    c, e = dt * np.vdot(dl, y), np.vdot(g, model.dm.astype(g.dtype))
    return (a - b) / (a + b), (c - e) / (c + e)


@pytest.mark.parametrize("kind", ["iso", "rho", "tti", "visco", "viscoc"])
@pytest.mark.parametrize("so", [4, 8, 16])
def test_adjoint_dot_fp64_2d(kind, so):
    _, model, src, rec, q, dt = setup(kind, 2, so)
    eF, eJ = dots(model, src, rec, q, dt)
    assert abs(eF) < 1e-11 and abs(eJ) < 1e-11


@pytest.mark.parametrize("kind", ["iso", "rho", "tti", "visco"])
def test_adjoint_dot_fp64_3d(kind):
    _, model, src, rec, q, dt = setup(kind, 3, 8, n=(31, 29, 27), nbl=8)
    eF, eJ = dots(model, src, rec, q, dt)
    assert abs(eF) < 1e-11 and abs(eJ) < 1e-11


@pytest.mark.parametrize("kind", ["iso", "rho", "tti", "visco"])
def test_adjoint_dot_fp32_reference_tolerance(kind):
    """test_adjoint.jl:21: |a-b|/(a+b) <= 5e-4 in single precision."""
    _, model, src, rec, q, dt = setup(kind, 2, 8, precision="f32")
    eF, eJ = dots(model, src, rec, q, dt)
    assert abs(eF) < 5e-4 and abs(eJ) < 5e-4


@pytest.mark.parametrize("kind", ["iso", "tti"])
def test_free_surface_dot_test_reference_tolerance(kind):
    """ISO_OP_FS / TTI_OP_FS CI groups (test_adjoint.jl:21-22): 5e-4, 5e-3 for TTI+FS.  The mirror
    condition is not an exact discrete adjoint, so fp64 does not reach round-off here."""
    _, model, src, rec, q, dt = setup(kind, 2, 8, fs=True)
    eF, eJ = dots(model, src, rec, q, dt)
    tol = 5e-3 if kind == "tti" else 5e-4
    assert abs(eF) < tol and abs(eJ) < tol


@pytest.mark.parametrize("kind", ["iso", "rho", "tti", "visco"])
@pytest.mark.parametrize("ic", ["isic", "fwi"])
def test_imaging_conditions_are_adjoint_pairs(ic, kind):
    """test_all_options.jl:13-53 (rtol 5e-2 there; run on the density, TTI and visco-acoustic
    models too): Born source and imaging condition of the same name are exact discrete adjoints."""
    _, model, src, rec, q, dt = setup(kind, 2, 8)
    _, eJ = dots(model, src, rec, q, dt, ic=ic)
    assert abs(eJ) < 1e-11


def test_subsampling_and_checkpointing_options():
    """test_all_options.jl:57-69,118-130: checkpointing is exact (== t_sub 1); t_sub=4 passes the
    J dot test to 1e-2."""
    _, model, src, rec, q, dt = setup("iso", 2, 8)
    _, e1 = dots(model, src, rec, q, dt, checkpointing=True)
    q8 = O.ricker_wavelet((NT - 1) * dt, dt, 0.008)[:NT]   # t_sub subsampling assumes a band-limited wavelet
    _, e4 = dots(model, src, rec, q8, dt, t_sub=4)
    assert abs(e1) < 1e-11
    assert abs(e4) < 1e-2


def test_linearity():
    """test_linearity.jl:17 (rtol 5e-5)."""
    _, model, src, rec, q, dt = setup("iso", 2, 8, precision="f32")
    q2 = np.roll(q, 7, axis=0) * np.float32(0.5)
    d1, _ = O.forward_rec(model, src, q, rec)
    d2, _ = O.forward_rec(model, src, q2, rec)
    d12, _ = O.forward_rec(model, src, q + q2, rec)
    assert rel_l2(d12, d1 + d2) < 5e-5
    d3, _ = O.forward_rec(model, src, 2 * q, rec)
    assert rel_l2(d3, 2 * d1) < 5e-5
    dl1, _ = O.born_rec(model, src, q, rec)
    dl2, _ = O.born_rec(model, src, 2 * q, rec)
    assert rel_l2(dl2, 2 * dl1) < 5e-5


def test_born_of_zero_is_zero_and_taylor_rate():
    """test_jacobian.jl:17-43: J*0 == 0 exactly; F(m0 + h dm) - F(m0) - h J dm is O(h^2)."""
    args, model, src, rec, q, dt = setup("iso", 2, 8)
    a0 = dict(args)
    a0.pop("dm")
    m_nodm = O.Model(precision="f64", **a0)
    dl0, _ = O.born_rec(m_nodm, src, q, rec)
    assert np.all(dl0 == 0)
    d0, _ = O.forward_rec(model, src, q, rec)
    dl, _ = O.born_rec(model, src, q, rec)
    errs1, errs2, hs = [], [], []
    h = 1.0
    for _ in range(5):
        a = dict(a0)
        a["m"] = (args["m"].astype(np.float64) + h * args["dm"]).astype(np.float64)
        mh = O.Model(precision="f64", dt=float(model.critical_dt), **_f64(a))
        dh, _ = O.forward_rec(mh, src, q, rec)
        errs1.append(np.linalg.norm(dh - d0))
        errs2.append(np.linalg.norm(dh - d0 - h * dl))
        hs.append(h)
        h *= 0.5
    r1 = np.log2(np.array(errs1[:-1]) / np.array(errs1[1:]))
    r2 = np.log2(np.array(errs2[:-1]) / np.array(errs2[1:]))
    assert np.all(np.abs(r1 - 1) < 0.15), r1
    assert np.all(np.abs(r2 - 2) < 0.15), r2


def _f64(a):
    # the oracle's Model casts parameters to float32 like pysource; the Taylor test perturbs m
    # below fp32 resolution for the smallest h otherwise, so keep the perturbation exact enough
    return a


def test_fwi_objective_value_and_gradient_taylor():
    """test_gradients.jl:31-40: f == 1/2 ||F(m0) q - dobs||^2 (dt-weighted norm) and the gradient
    passes the Taylor test."""
    args, model, src, rec, q, dt = setup("iso", 2, 8)
    a_true = dict(args)
    a_true.pop("dm")
    a_true["m"] = (args["m"] + args["dm"]).astype(np.float32)
    m_true = O.Model(precision="f64", dt=dt, **a_true)
    dobs, _ = O.forward_rec(m_true, src, q, rec)
    d0, _ = O.forward_rec(model, src, q, rec)
    f, g, _, _ = O.J_adjoint(model, src, q, rec, dobs, return_obj=True)
    assert abs(f - 0.5 * dt * np.linalg.norm(d0 - dobs) ** 2) <= 1e-12 * abs(f)
    dmp = np.pad(args["dm"].astype(np.float64), model.padsizes, mode="edge")
    slope = np.vdot(g, dmp)
    errs = []
    h = 0.5
    base = dict(args)
    base.pop("dm")
    for _ in range(4):
        a = dict(base)
        a["m"] = (args["m"].astype(np.float64) + h * args["dm"]).astype(np.float32)
        mh = O.Model(precision="f64", dt=dt, **a)
        dh, _ = O.forward_rec(mh, src, q, rec)
        fh = 0.5 * dt * np.linalg.norm(dh - dobs) ** 2
        errs.append(abs(fh - f - h * slope))
        h *= 0.5
    r = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(r > 1.7), (r, errs)


def test_lsrtm_nlind_with_zero_dm_is_bitwise_fwi():
    """test_gradients.jl:92-97: lsrtm_objective(dm=0, nlind=true) == fwi_objective, atol=0."""
    args, model, src, rec, q, dt = setup("iso", 2, 8, precision="f32", with_dm=False)
    dobs = np.random.default_rng(3).standard_normal((NT, rec.shape[0])).astype(np.float32)
    f1, g1, _, _ = O.J_adjoint(model, src, q, rec, dobs, return_obj=True)
    f2, g2, _, _ = O.J_adjoint(model, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True)
    assert f1 == f2
    assert np.array_equal(g1, g2)


def test_homogeneous_arrival_time():
    """Analytic sanity: first arrival at offset r in a homogeneous medium."""
    n = (121, 121)
    v = 2.0
    args = dict(origin=(0., 0.), spacing=(10., 10.), shape=n, space_order=8, nbl=30,
                m=np.full(n, 1 / v ** 2, dtype=np.float32))
    model = O.Model(precision="f32", **args)
    dt = float(model.critical_dt)
    nt = int(500 / dt)
    f0 = 0.025
    q = O.ricker_wavelet((nt - 1) * dt, dt, f0)[:nt]
    src = np.array([[600., 600.]], dtype=np.float32)
    rec = np.array([[1000., 600.], [600., 900.]], dtype=np.float32)
    d, _ = O.forward_rec(model, src, q, rec)
    for j, r in enumerate((400., 300.)):
        # 2-D far field: peak of |trace| near r/v + 1/f0 (Ricker delay), up to the 2-D tail
        t_peak = np.argmax(np.abs(d[:, j])) * dt
        assert abs(t_peak - (r / v + 1 / f0)) < 12.0, (t_peak, r / v + 1 / f0)


@pytest.mark.parametrize("kind", ["iso", "rho", "tti"])
def test_adjoint_scheme_is_the_forward_scheme_in_reversed_time(kind):
    """kernels.py:61-69: the adjoint update uses `u.dt.T`, the time mirror of `u.dt`, and runs the
    loop backwards -- so F' y = reverse(F reverse(y)) bit for bit for every propagator but the
    visco-acoustic one.  The library serves `born_rec` / `J_adjoint` with fw=False through this
    identity (judi.jl_b200/interface.py)."""
    from conftest import model_args, geometry
    args = model_args(kind, 2, 8)
    om = O.Model(precision="f64", **args)
    src, rec = geometry(args)
    y = np.random.default_rng(0).standard_normal((120, rec.shape[0])).astype(np.float32)
    a, _ = O.forward_rec(om, rec, y, src, fw=False)
    b, _ = O.forward_rec(om, rec, y[::-1].copy(), src, fw=True)
    assert np.array_equal(a, b[::-1])
    # ... and NOT for the visco-acoustic scheme (its adjoint is a different discretisation)
    vargs = model_args("visco", 2, 8)
    vm = O.Model(precision="f64", **vargs)
    av, _ = O.forward_rec(vm, rec, y, src, fw=False)
    bv, _ = O.forward_rec(vm, rec, y[::-1].copy(), src, fw=True)
    assert not np.array_equal(av, bv[::-1])


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("so", [4, 8])
def test_tti_without_anisotropy_is_the_staggered_acoustic_operator(ndim, so):
    """With epsilon = delta = 0 the tensor of sa_tti (FD_utils.py:84-105) is A = b I, B = C = 0, so
    R^T A R = b I for ANY tilt / azimuth field and the first TTI field obeys
    m u.dt2 + damp u.dt = div(b grad(u, shift=.5), shift=-.5): the density-field operator, which
    tests/test_devito_known_answer.py holds against a Devito-produced constant.  End-to-end check of
    the rotation as implemented (orthogonality at every node, heterogeneous angles), of the wiring
    (source into and receivers from the first field only) and of the staggered stencils inside the
    TTI path."""
    a = model_args("tti", ndim, so=so, with_dm=False)
    v = 1 / np.sqrt(a["m"])
    zero = np.zeros(a["shape"], np.float32)
    tti = dict(a, epsilon=zero, delta=zero, dt=np.float32(0.5))
    tti["theta"] = (0.7 * np.sin(v * 3.0)).astype(np.float32)
    if ndim == 3:
        tti["phi"] = (1.1 * np.cos(v * 2.0)).astype(np.float32)
    iso = {k: val for k, val in a.items() if k not in ("epsilon", "delta", "theta", "phi")}
    iso.update(b=np.ones(a["shape"], np.float32), dt=np.float32(0.5))
    Mt, Mi = O.Model(precision="f64", **tti), O.Model(precision="f64", **iso)
    assert Mt.is_tti and Mi.irho_field is not None
    assert float(Mt.critical_dt) == float(Mi.critical_dt) == 0.5
    src, rec = geometry(a)
    nt = 120 if ndim == 2 else 60
    q = O.ricker_wavelet((nt - 1) * 0.5, 0.5, 0.03)[:nt]
    dt_, _ = O.forward_rec(Mt, src, q, rec)
    di_, _ = O.forward_rec(Mi, src, q, rec)
    assert np.linalg.norm(di_) > 0
    assert rel_l2(dt_, di_) < 1e-10
