"""On-the-fly DFT gradients (J_adjoint_freq, interface.py:348-408).

CPU: the oracle restatement against a literal evaluation of fields_exprs.py:140-169 /
sensitivity.py:72-104 and against the time-domain gradient (the weight -(2 pi f)^2 makes the
frequency sum a band-limited version of u * v.dt2).  GPU: parity of the CUDA path with the oracle,
dft subsampling, TTI field pairs, and the reference's own checks (test_all_options.jl:74-115:
J*0 == 0, finite results, <J dm, y> vs <dm, J' y>)."""
import numpy as np
import pytest

from conftest import model_args, geometry, rel_l2
from oracle import oracle as O

NT = 200
FREQS = [0.006, 0.011, 0.017, 0.024]


def _setup(kind, ndim, **kw):
    args = model_args(kind, ndim, 8, **kw)
    om = O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dt = float(om.critical_dt)
    q = O.ricker_wavelet((NT - 1) * dt, dt, 0.015)[:NT]
    return args, om, src, rec, q, dt


def test_oracle_dft_gradient_matches_literal_loops():
    args, om, src, rec, q, dt = _setup("iso", 2, with_dm=False, n=(41, 35), nbl=8)
    do, u, _ = O.forward(om, src, rec, q, save=True)
    y = (0.1 * do).astype(np.float32)
    g = O.J_adjoint_freq(om, src, q, rec, y, FREQS[:2], dft_sub=3, is_residual=True)
    _, v, _ = O.forward(om, rec, None, y, save=True, fw=False)
    ref = np.zeros(om.shape_pml)
    for f in FREQS[:2]:
        uf = np.zeros(om.shape_pml, dtype=complex)
        for t in range(1, NT - 1):
            if t % 3 == 0:
                uf += 3 * np.exp(-1j * 2 * np.pi * f * t * dt) * u[t, 0]
        w = -(2 * np.pi * f) ** 2 / (NT - 2)
        for t in range(NT - 2, 0, -1):      # every step: crosscorr_freq never sees dft_sub (sensitivity.py:50)
            ref -= (w * uf * np.exp(1j * 2 * np.pi * f * t * dt) * v[t, 0]).real
    assert rel_l2(g, ref) < 1e-5


def test_oracle_dft_gradient_approaches_time_domain_gradient():
    """With the whole band sampled the DFT gradient points along the time-domain one."""
    args, om, src, rec, q, dt = _setup("iso", 2, with_dm=False, n=(61, 51), nbl=10)
    do, _ = O.forward_rec(om, src, q, rec)
    y = (0.1 * do).astype(np.float32)
    gt, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True)
    freqs = np.arange(1, NT // 2) / (NT * dt)                  # all DFT bins up to Nyquist (kHz)
    gf = O.J_adjoint_freq(om, src, q, rec, y, freqs, is_residual=True)
    crop = tuple(slice(12, -12) for _ in range(2))
    a, b = gt[crop].ravel().astype(np.float64), gf[crop].ravel().astype(np.float64)
    assert np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)) > 0.95


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ndim,dft_sub", [("iso", 2, None), ("iso", 2, 4), ("iso", 3, 2),
                                               ("tti", 2, 3), ("tti", 3, None), ("rho", 3, 4),
                                               ("rho", 2, 2), ("visco", 2, None), ("viscoc", 3, 3)])
def test_dft_gradient_parity(kind, ndim, dft_sub):
    import judi_jl_b200 as J
    args, om, src, rec, q, dt = _setup(kind, ndim, with_dm=True)
    gm = J.Model(**args)
    do, _ = O.forward_rec(om, src, q, rec)
    dobs = (0.9 * do).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, freq_list=FREQS, dft_sub=dft_sub)
    fo, go = O.J_adjoint_freq(om, src, q, rec, dobs, FREQS, dft_sub=dft_sub, return_obj=True)
    assert np.isfinite(g).all() and np.linalg.norm(go) > 0
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, go) < 1e-3                                 # north-star gradient tolerance
    # lsrtm_objective with frequencies (test_gradients.jl:104-131): f and g as the algebra says
    fl, gl, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True,
                               freq_list=FREQS, dft_sub=dft_sub)
    dl, _ = J.born_rec(gm, src, q, rec)
    d0, _ = J.forward_rec(gm, src, q, rec)
    res = (d0 + dl - dobs).astype(np.float32)
    g2, _, _ = J.J_adjoint(gm, src, q, rec, res, is_residual=True, freq_list=FREQS, dft_sub=dft_sub)
    assert abs(float(fl) - 0.5 * dt * np.linalg.norm(res) ** 2) <= 5e-4 * abs(float(fl))
    assert rel_l2(gl, g2) < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ndim,dft_sub", [("iso", 2, None), ("iso", 2, 3), ("iso", 3, 2), ("tti", 2, 2),
                                               ("rho", 2, None), ("tti", 3, None), ("visco", 2, 2)])
@pytest.mark.parametrize("ic", ["isic", "fwi"])
def test_dft_gradient_isic_fwi_parity(kind, ndim, dft_sub, ic):
    """test_all_options.jl:133-172: IC="isic" / "fwi" together with frequencies -- isic_freq /
    fwi_freq (sensitivity.py:125-154), which subsample the adjoint correlation by dft_sub."""
    import judi_jl_b200 as J
    args, om, src, rec, q, dt = _setup(kind, ndim, with_dm=True)
    gm = J.Model(**args)
    do, _ = O.forward_rec(om, src, q, rec)
    dobs = (0.9 * do).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, freq_list=FREQS, dft_sub=dft_sub, ic=ic)
    fo, go = O.J_adjoint_freq(om, src, q, rec, dobs, FREQS, dft_sub=dft_sub, return_obj=True, ic=ic)
    assert np.isfinite(g).all() and np.linalg.norm(go) > 0
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, go) < 1e-3                                 # north-star gradient tolerance
    ga, _, _ = J.J_adjoint(gm, src, q, rec, dobs, freq_list=FREQS, dft_sub=dft_sub)
    assert rel_l2(g, ga) > 1e-2                                 # and it is not the "as" gradient


@pytest.mark.gpu
def test_checkpointing_takes_precedence_over_frequencies():
    """interface.py:331-345: J_adjoint dispatches on `checkpointing` first; a frequency list given
    together with it is ignored."""
    import judi_jl_b200 as J
    args, om, src, rec, q, dt = _setup("iso", 2, with_dm=False)
    gm = J.Model(**args)
    y = np.random.default_rng(1).standard_normal((NT, rec.shape[0])).astype(np.float32)
    g0, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True, checkpointing=True)
    g1, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True, checkpointing=True, freq_list=FREQS)
    assert np.linalg.norm(g0) > 0 and np.array_equal(g0, g1)


def test_numpy_inner_grad_is_the_c_loops_inner_grad():
    """The numpy inner_grad of the frequency-domain conditions against the C oracle's time-domain
    isic gradient, rebuilt from the wavefield histories: gradm -= dt irho (u v.dt2 m + inner_grad(u, v))."""
    args, om, src, rec, q, dt = _setup("iso", 2, with_dm=False, n=(31, 27), nbl=6)
    om = O.Model(precision="f64", **args)
    nt = 60
    q = q[:nt]
    do, u, _ = O.forward(om, src, rec, q, save=True)
    y = 0.1 * do
    g, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True, ic="isic")
    _, v, _ = O.forward(om, rec, None, y, save=True, fw=False)
    inner = O.inner_grad_np(om)
    m = om.m.astype(np.float64)
    ref = np.zeros(om.shape_pml)
    for t in range(1, nt - 1):
        vdt2 = (v[t + 1, 0] - 2 * v[t, 0] + v[t - 1, 0]) / dt ** 2
        ref -= dt * om.irho_const * (u[t, 0] * vdt2 * m + inner(u[t, 0], v[t, 0]))
    assert rel_l2(ref, g) < 1e-12
