"""The oracle (and, on a GPU box, the CUDA library) against vectors produced by EXECUTING the
reference's own Python -- src/pysource with a `devito` import stub, see
tests/golden/make_reference_golden.py, which also lists the reference file:line of every item.

What this pins to the reference's executed code: the CFL coefficient and critical dt, padsizes,
the rotated TTI flux algebra (R_mat / thomsen_mat / sa_tti), signs and factors of the "as" / "isic"
/ "fwi" imaging conditions and Born sources, Loss, the number of saved snapshots, the
visco-acoustic SLS update (modulo the time-stencil convention the stub states) and the
absorbing-layer field of both abc types (damp_op's Eq / Inc list, interpreted on numpy), which
field / time level the source goes into with which factor and what the receivers read
(geom_expr), and the Ricker wavelet / time axis of sources.py.  What it
cannot pin: the Devito-generated stencil loops themselves.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))["cases"]


def _model_kw(c):
    nd = c["ndim"]
    m = np.asarray(c["m"], dtype=np.float32)
    shape = (len(m),) + (5,) * (nd - 1)
    bc = m.reshape((-1,) + (1,) * (nd - 1)) * np.ones(shape, dtype=np.float32)
    kw = dict(origin=(0.,) * nd, spacing=(c["h"],) * nd, shape=shape, space_order=c["so"], nbl=6,
              m=bc.astype(np.float32))
    if c["tti"]:
        e = np.asarray(c["eps"], dtype=np.float32).reshape((-1,) + (1,) * (nd - 1))
        kw["epsilon"] = (e * np.ones(shape, dtype=np.float32)).astype(np.float32)
        kw["delta"] = (0.5 * kw["epsilon"]).astype(np.float32)
        kw["theta"] = np.zeros(shape, dtype=np.float32)
    if c["user_dt"] is not None:
        kw["dt"] = c["user_dt"]
    return kw


@pytest.mark.parametrize("i", range(0, len(GOLD["critical_dt"]), 3))
def test_cfl_and_critical_dt(i):
    """models.py:440-482 executed: dt = f32("%.3e" % (cfl h / (sqrt(1 + 2 max eps) vmax)))."""
    c = GOLD["critical_dt"][i]
    assert O.cfl_coeff(c["so"], c["ndim"]) == pytest.approx(c["cfl"], rel=1e-14)
    om = O.Model(precision="f32", **_model_kw(c))
    assert np.float32(om.critical_dt) == np.float32(c["dt"])


@pytest.mark.parametrize("c", GOLD["padsizes"], ids=lambda c: "%dd_fs%d" % (c["ndim"], c["fs"]))
def test_padsizes(c):
    nd = c["ndim"]
    om = O.Model((0.,) * nd, (10.,) * nd, (9,) * nd, space_order=8, nbl=c["nbl"],
                 m=np.full((9,) * nd, 0.25, dtype=np.float32), fs=c["fs"], precision="f32")
    assert [list(p) for p in om.padsizes] == c["padsizes"]


def _closed_form(ndim, b, e, d, theta, phi, g1, g2):
    """The form the CUDA kernels evaluate (flux3d.cu / tti3d.cu): through the symmetry axis r."""
    if ndim == 2:
        r = np.array([np.sin(theta), 0., np.cos(theta)])
    else:
        r = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
    A0, B0, C0 = b * d, b * np.sqrt((e - d) * d), b * (e - d)
    d1, d2 = r @ g1, r @ g2
    P = A0 * g1 + B0 * g2 + ((b - A0) * d1 - B0 * d2) * r
    M = B0 * g1 + C0 * g2 - (B0 * d1 + C0 * d2) * r
    return P, M


@pytest.mark.parametrize("i", range(len(GOLD["tti_flux"])))
def test_tti_flux_algebra(i):
    """FD_utils.sa_tti executed symbolically (grad -> symbols, div -> captured argument)."""
    c = GOLD["tti_flux"][i]
    v, nd = c["inputs"], c["ndim"]
    if nd == 3:
        g1 = np.array([v["gu_x"], v["gu_y"], v["gu_z"]]); g2 = np.array([v["gv_x"], v["gv_y"], v["gv_z"]])
        sel = [0, 1, 2]
    else:
        g1 = np.array([v["gu_x"], 0., v["gu_z"]]); g2 = np.array([v["gv_x"], 0., v["gv_z"]])
        sel = [0, 2]
    phi = v.get("phi", 0.)
    P, M = O.tti_flux_point(nd, v["b"], v["eps"], v["delta"], v["theta"], phi, g1, g2, precision="f64")
    assert np.allclose(P[sel], c["P"], rtol=1e-12, atol=1e-14)
    assert np.allclose(M[sel], c["M"], rtol=1e-12, atol=1e-14)
    P32, M32 = O.tti_flux_point(nd, v["b"], v["eps"], v["delta"], v["theta"], phi, g1, g2, precision="f32")
    assert np.allclose(P32[sel], c["P"], rtol=2e-5, atol=2e-6)
    assert np.allclose(M32[sel], c["M"], rtol=2e-5, atol=2e-6)
    Pc, Mc = _closed_form(nd, v["b"], v["eps"], v["delta"], v["theta"], phi, g1, g2)
    assert np.allclose(Pc[sel], c["P"], rtol=1e-12, atol=1e-14)
    assert np.allclose(Mc[sel], c["M"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("c", GOLD["imaging"], ids=lambda c: "%s_nf%d" % (c["ic"], c["nf"]))
def test_imaging_conditions_and_born_sources(c):
    """sensitivity.py ic_dict / ls_dict executed on symbolic fields: gradm -= w * sum_f ic(u_f, v_f)
    with w = dt_save * irho, and q_f = lin_src(u_f)."""
    v = c["inputs"]
    w = v["dt_save"] * v["irho"]
    tot = 0.
    for f in range(c["nf"]):
        ig = 0.
        if c["ic"] != "as":
            ig = sum(v["gu%d_%s" % (f, d)] * v["gv%d_%s" % (f, d)] for d in "xyz")
        tot += O.ic_point(c["ic"], v["u%d" % f], v["v%d_dt2" % f], v.get("m", 1.), ig)
    assert w * tot == pytest.approx(c["grad_expr"], rel=1e-12)
    for f in range(c["nf"]):
        lapw = v.get("div%d" % f, 0.)
        q = O.linsrc_point(c["ic"], v["dm"], v["irho"], v.get("m", 1.), v["u%d_dt2" % f], lapw)
        assert q == pytest.approx(c["lin_src"][f], rel=1e-12)
        if c["ic"] != "as":
            # laplacian(u, dm*irho) = div(dm * irho * grad(u, +1/2), -1/2): node-valued coefficient
            want = [v["dm"] * v["irho"] * v["gu%d_%s" % (f, d)] for d in "xyz"]
            assert np.allclose(c["div_args"]["div%d" % f], want, rtol=1e-12)


@pytest.mark.parametrize("r", GOLD["loss"]["results"], ids=lambda r: r["mode"])
def test_loss(r):
    L = GOLD["loss"]
    dsyn = np.asarray(L["dsyn"], dtype=np.float32)
    dlin = np.asarray(L["dlin"], dtype=np.float32)
    dobs = np.asarray(L["dobs"], dtype=np.float32)
    l1 = lambda a, b: (float(np.sum(np.abs(a - b))), np.sign(a - b).astype(np.float32))
    if r["mode"] == "nlind":
        f, res = O.Loss((dsyn, dlin), dobs, r["dt"])
    elif r["mode"] == "misfit":
        f, res = O.Loss(dsyn, dobs, r["dt"], misfit=l1)
    elif r["mode"] == "misfit_nlind":
        f, res = O.Loss((dsyn, dlin), dobs, r["dt"], misfit=l1)
    else:
        f, res = O.Loss(dsyn, dobs, r["dt"], is_residual=(r["mode"] == "is_residual"))
    assert float(f) == pytest.approx(r["f"], rel=1e-6)
    assert np.allclose(res, np.asarray(r["residual"]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("i", range(len(GOLD["visco"])))
def test_visco_acoustic_update(i):
    """kernels.SLS_2nd_order executed symbolically (forward and adjoint): relaxation parameters,
    signs, where b, m and the mask-type damp enter, and what the adjoint's Laplacians act on."""
    c = GOLD["visco"][i]
    v = c["inputs"]
    dl = 1.0 - v["damp"]                      # the reference's field is the mask 1 - layer profile
    if c["fw"]:
        rn, pn, _, _ = O.visco_point(0, v["dt"], v["f0"], v["qp"], v["b"], v["m"], dl, v["p"], v["p_b"],
                                     v["rp"], v["div0"], v["q"])
    else:
        L = v["div0"] + v["div1"]             # laplacian((1+tt) p, b) + laplacian((tt/(b t_s)) r-, b)
        rn, pn, wp, wr = O.visco_point(1, v["dt"], v["f0"], v["qp"], v["b"], v["m"], dl, v["p"], v["p_f"],
                                       v["rp"], L, v["q"])
        assert wp == pytest.approx(c["w_factor_p"], rel=1e-11)
        assert wr == pytest.approx(c["w_factor_r"], rel=1e-9)
    assert rn == pytest.approx(c["r_new"], rel=1e-11)
    assert pn == pytest.approx(c["p_new"], rel=1e-10)


@pytest.mark.parametrize("i", range(len(GOLD["update"])))
def test_acoustic_and_tti_time_update(i):
    """kernels.acoustic_kernel / tti_kernel executed symbolically, forward and adjoint:
    solve(m irho u.dt2 + damp u.dt[.T] - L - q, u+-) with L = irho * u.laplace for a scalar density,
    div(irho grad u) for a density field, (H0, H1) of sa_tti for TTI."""
    c = GOLD["update"][i]
    v = c["inputs"]
    other = "_b" if c["fw"] else "_f"
    for f, name in enumerate(["u1", "u2"][:len(c["new"])]):
        if c["kind"] == "acoustic":
            L = c["irho"] * v[name + "_lap"]
        else:
            L = v["div%d" % f]
        S = L + v["q%d" % (f + 1)]
        un = O.leapfrog_point(v["dt"], v["m"], c["irho"], v["damp"], v[name], v[name + other], S)
        assert un == pytest.approx(c["new"][f], rel=1e-10)


def test_number_of_saved_snapshots():
    for c in GOLD["nsave"]:
        if c["nsave"] is not None:
            assert O.nsave(c["nt"], c["t_sub"]) == c["nsave"], c


def _one_step_setup(c, ndim=2):
    """A model with varying m and density (and Thomsen parameters), one source exactly on a grid
    node: after the first time step from rest the wavefield holds nothing but the injection."""
    n = (9, 8) if ndim == 2 else (7, 6, 8)
    rng = np.random.default_rng(3)
    kw = dict(origin=(0.,) * ndim, spacing=(10.,) * ndim, shape=n, space_order=8, nbl=6,
              m=rng.uniform(0.1, 0.4, n).astype(np.float32), rho=rng.uniform(0.9, 2.2, n).astype(np.float32))
    if c["tti"]:
        kw.update(epsilon=rng.uniform(0.1, 0.2, n).astype(np.float32), delta=rng.uniform(0., 0.1, n).astype(np.float32),
                  theta=rng.uniform(-.5, .5, n).astype(np.float32))
    node = tuple(s // 2 for s in n)
    src = np.array([[10. * i for i in node]], dtype=np.float32)
    q = np.array([[0.3], [-1.7], [0.9], [0.2]], dtype=np.float32)
    return kw, node, src, q


@pytest.mark.parametrize("i", range(len(GOLD["geometry"])))
def test_injection_target_and_factor(i):
    """geom_utils.py:61-96 executed: the source goes into the NEW time level (u.forward, or
    u.backward for the adjoint) of the FIRST field only, as src * dt^2 / (m * irho); receivers
    read the first field at the current level."""
    c = GOLD["geometry"][i]
    assert c["inject_field"] == ("u1_f" if c["fw"] else "u1_b") and c["interpolate"] == "u1"
    v = c["inputs"]
    assert np.isclose(c["injected"], v[c["src_name"]] * v["dt"] ** 2 / (v["m"] * v["b"]), rtol=1e-12)
    kw, node, src, q = _one_step_setup(c)
    om = O.Model(precision="f64", **kw)
    dt = float(om.critical_dt)
    _, usave, _ = O.forward(om, src, None, q, save=True, fw=c["fw"])
    pn = tuple(i_ + l for i_, (l, r) in zip(node, om.padsizes))
    m, b = float(kw["m"][node]), 1.0 / float(kw["rho"][node])
    # forward: the first step is t = 1 (writes level 2 from src[1]); adjoint: t = nt - 2 (writes
    # level nt - 3 from src[nt - 2])
    lvl, row = (2, 1) if c["fw"] else (q.shape[0] - 3, q.shape[0] - 2)
    expect = float(q[row, 0]) * dt ** 2 / (m * b)
    got = usave[lvl, 0]
    assert np.isclose(got[pn], expect, rtol=1e-6)
    assert np.count_nonzero(got) == 1
    if c["tti"]:
        assert np.count_nonzero(usave[lvl, 1]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("tti", [False, True])
def test_library_injection_target_and_factor(tti, ndim):
    """The same known answer through the C-ABI (judi_forward_no_rec returns the wavefield history)."""
    import judi_jl_b200 as J
    c = dict(tti=tti, fw=True)
    kw, node, src, q = _one_step_setup(c, ndim)
    gm = J.Model(**kw)
    dt = float(gm.critical_dt)
    u, _ = J.forward_no_rec(gm, src, q)
    pn = tuple(i_ + l for i_, (l, r) in zip(node, gm.padsizes))
    expect = float(q[1, 0]) * dt ** 2 / (float(kw["m"][node]) / float(kw["rho"][node]))
    assert np.isclose(u[2][pn], expect, rtol=1e-5)
    assert np.count_nonzero(u[2]) == 1


@pytest.mark.parametrize("c", GOLD["ricker"])
def test_ricker_wavelet_and_time_axis(c):
    """sources.py:11-66 (TimeAxis) and :160-176 (RickerSource.wavelet) executed, against the
    oracle's restatement of the Julia helper the reference's scripts use (auxiliaryFunctions.jl:205-213)."""
    t0, dt, nt = O.time_axis(c["dt"], c["tmax"])
    assert nt == c["num"]
    if "wavelet" not in c:
        return
    w = O.ricker_wavelet(c["tmax"], c["dt"], c["f0"])[:160, 0]
    ref = np.asarray(c["wavelet"])[:len(w)]
    assert np.abs(w - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("i", range(len(GOLD["dft"])))
def test_on_the_fly_dft_and_frequency_imaging_condition(i):
    """fields_exprs.otf_dft + sensitivity.grad_expr -> crosscorr_freq / isic_freq / fwi_freq executed
    on the time series of one grid point: the forward DFT is subsampled by dft_sub (and scaled by
    it); the "as" correlation runs over `time`, never `tsave` -- every step, whatever dft_sub
    is -- while isic_freq / fwi_freq do read `factor` and run on the subsampled steps, with
    w2 = +-factor / time_M on the inner_grad term (a bilinear stand-in kappa * a * b in the vectors)."""
    c = GOLD["dft"][i]
    sub = bool(c["dft_sub"] and c["dft_sub"] > 1)
    assert ("tsave" in c["forward_symbols"]) == sub
    assert ("tsave" in c["gradient_symbols"]) == (sub and c["ic"] != "as")
    u = np.asarray(c["u"])[:, :, None]
    v = np.asarray(c["v"])[:, :, None]
    g = O.dft_gradient(u, v, c["dt"], c["freqs"], c["dft_sub"], ic=c["ic"], m=c["m"],
                       inner_grad=lambda a, b: c["kappa"] * a * b)
    assert np.isclose(g[0], c["gradm"], rtol=1e-10, atol=1e-14)


def _damp_model_kw(c):
    n = tuple(c["shape"])
    return dict(origin=(0.,) * c["ndim"], spacing=tuple(c["spacing"]), shape=n, space_order=8,
                nbl=c["nbl"], fs=c["fs"], m=np.full(n, 0.25, dtype=np.float32))


def _check_damp(c, damp):
    """`damp` (abc_type "damp", padded grid) against the entry's full array or its lines."""
    sign, base = (1., 0.) if c["abc"] == "damp" else (-1., 1.)    # mask = 1 - damp (models.py:73-93)
    got = base + sign * np.asarray(damp, dtype=np.float64)
    if "damp" in c:
        ref = np.asarray(c["damp"])
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 3e-7 * max(np.abs(ref).max(), 1.)
        return
    nd = c["ndim"]
    mid = tuple(s // 2 for s in got.shape)
    for ax in range(nd):
        line = got[tuple(slice(None) if a == ax else mid[a] for a in range(nd))]
        ref = np.asarray(c["lines"][ax])
        assert np.abs(line - ref).max() <= 3e-7 * max(np.abs(ref).max(), 1.)
    diag = np.array([got[(i,) * nd] for i in range(min(got.shape))])
    assert np.abs(diag - np.asarray(c["diag"])).max() <= 3e-7 * max(np.abs(diag).max(), 1.)


@pytest.mark.parametrize("i", range(len(GOLD["damp"])))
def test_absorbing_layer_field(i):
    """models.py:53-98 (damp_op) executed: layer width nbl - 3, coefficient 1.5 ln(1000) / (nbl - 3),
    the +1 in the position, division by the spacing of each dimension, no top layer with a free
    surface, and the mask form 1 - profile the visco-acoustic equations use."""
    c = GOLD["damp"][i]
    om = O.Model(precision="f32", **_damp_model_kw(c))
    assert [list(t) for t in om.padsizes] == c["padsizes"]
    _check_damp(c, om.damp)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(0, len(GOLD["damp"]), 2))
def test_library_absorbing_layer_matches_reference(i):
    """The separable 1-D profiles of the CUDA library, summed on the device, against the same vectors."""
    import judi_jl_b200 as J
    c = GOLD["damp"][i]
    gm = J.Model(**_damp_model_kw(c))
    _check_damp(c, gm.get_field("damp"))


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(0, len(GOLD["critical_dt"]), 7))
def test_library_critical_dt_matches_reference(i):
    """judi_model_critical_dt (C-ABI) against the reference's executed formula."""
    import judi_jl_b200 as J
    c = GOLD["critical_dt"][i]
    gm = J.Model(**_model_kw(c))
    assert np.float32(gm.critical_dt) == np.float32(c["dt"])


@pytest.mark.parametrize("so", [4, 8, 12, 16])
def test_finite_difference_weights_are_sympy_taylor_weights(so):
    """Devito builds its stencils from sympy.finite_diff_weights (the same function models.py:5,452
    imports for the CFL number): centred 2nd derivative on so+1 points, and the staggered first
    derivative evaluated half a cell off its so points (grad(., shift=.5) / div(., shift=-.5))."""
    from sympy import finite_diff_weights, Rational
    r = so // 2
    om = O.Model((0., 0.), (10., 10.), (9, 9), space_order=so, nbl=6,
                 m=np.full((9, 9), 0.25, dtype=np.float32), precision="f64")
    c2, w1 = O.fd_weights(om)
    ref2 = finite_diff_weights(2, list(range(-r, r + 1)), 0)[2][-1]
    assert np.allclose(c2, [float(ref2[r + k]) for k in range(r + 1)], rtol=1e-13, atol=0)
    pts = [Rational(2 * k - 1, 2) for k in range(-r + 1, r + 1)]          # -r+1/2 .. r-1/2
    ref1 = finite_diff_weights(1, pts, 0)[1][-1]
    assert np.allclose(w1, [float(ref1[r + k - 1]) for k in range(1, r + 1)], rtol=1e-13, atol=0)
