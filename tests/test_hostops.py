"""The steps around every propagator call: time_resample
(time_utilities.jl:32-48,84), remove_padding (auxiliaryFunctions.jl:79-89),
limit_model_to_receiver_area (auxiliaryFunctions.jl:129-173) and the per-shot driver
(time_modeling_serial.jl:9-51).  CPU tests pin the oracle restatements on known answers; GPU tests
compare the device implementations (through the C-ABI) with them."""
import numpy as np
import pytest

from conftest import model_args, geometry, rel_l2
from oracle import oracle as O


# ---- oracle restatements (CPU) -------------------------------------------------------------------
def test_time_axis_and_resample_cases():
    """time_utilities.jl:34-41: equal dt -> shifted view; integer ratio -> decimation."""
    assert O.time_axis(4.0, 1000.0) == (np.float32(0), np.float32(4), 251)
    assert O.time_axis(1.5, 10.0)[2] == 8                      # 0:1.5:(1.5*ceil(10/1.5))
    rng = np.random.default_rng(3)
    d = rng.standard_normal((101, 4)).astype(np.float32)
    assert np.array_equal(O.time_resample(d, (0., 2., 101), (0., 2., 101)), d)
    assert np.array_equal(O.time_resample(d, (0., 2., 101), (8., 2., 97)), d[4:])
    assert np.array_equal(O.time_resample(d, (0., 2., 101), (0., 4., 51)), d[::2])
    assert np.array_equal(O.time_resample(d, (0., 2., 101), (8., 4., 49)), d[::2][2:])


def test_sinc_resample_reproduces_band_limited_signal_and_nodes():
    """time_utilities.jl:84: the sinc matrix is the identity on coinciding samples and
    interpolates a band-limited signal (well below Nyquist) to ~1e-3 away from the ends."""
    dt_in, nt_in = np.float32(4.0), 501
    t_in = np.arange(nt_in) * 4.0
    f = 0.008                                                   # kHz, Nyquist is 0.125
    sig = (np.exp(-((t_in - 1000.) / 300.) ** 2) * np.sin(2 * np.pi * f * t_in)).astype(np.float32)
    dt_new, nt_new = np.float32(1.733), int(np.floor(2000. / 1.733)) + 1
    out = O.time_resample(sig, (0., dt_in, nt_in), (0., dt_new, nt_new))
    t_new = np.arange(nt_new) * np.float64(dt_new)
    ref = np.exp(-((t_new - 1000.) / 300.) ** 2) * np.sin(2 * np.pi * f * t_new)
    assert out.shape == (nt_new, 1)
    assert np.abs(out[:, 0] - ref).max() < 2e-3
    # down then up returns the original samples where the axes coincide (dt ratio 3/2)
    back = O.time_resample(sig, (0., 4., nt_in), (0., 6., 334))
    assert np.allclose(back[::2, 0], sig[::3][:167], atol=2e-6)


def test_limit_model_matches_julia_index_arithmetic():
    """auxiliaryFunctions.jl:143-157 on a 2-D and a 3-D model (1-based `÷` arithmetic)."""
    shape, spacing, origin = (101, 51), (10., 10.), (0., 0.)
    m = np.arange(101 * 51, dtype=np.float32).reshape(shape)
    src = np.array([[400., 20.]], dtype=np.float32)
    rec = np.stack([np.linspace(250., 655., 30), np.full(30, 20.)], axis=1).astype(np.float32)
    o, n, p, pert, inds = O.limit_model_to_receiver_area(src, rec, origin, spacing, shape, 100.,
                                                         {"m": m, "rho": 1.0}, pert=m.ravel())
    assert inds[0] == slice(15, 76) and inds[1] == slice(0, 51)   # (250-100)/10 .. (655+100)÷10
    assert n == (61, 51) and o == (np.float32(150.), np.float32(0.))
    assert np.array_equal(p["m"], m[15:76]) and p["rho"] == 1.0
    assert np.array_equal(pert, m[15:76])
    # the buffer is clipped at the model edges
    o, n, p, _, inds = O.limit_model_to_receiver_area(src, rec, origin, spacing, shape, 5000.,
                                                      {"m": m})
    assert n == shape and o == (np.float32(0.), np.float32(0.))
    shape3 = (41, 31, 21)
    m3 = np.zeros(shape3, np.float32)
    src3 = np.array([[200., 100., 10.]], dtype=np.float32)
    rec3 = np.array([[120., 60., 10.], [260., 180., 10.]], dtype=np.float32)
    o, n, p, _, inds = O.limit_model_to_receiver_area(src3, rec3, (0., 0., 0.), (10.,) * 3, shape3,
                                                      30., {"m": m3})
    assert inds == (slice(9, 30), slice(3, 22), slice(0, 21))
    assert o == (np.float32(90.), np.float32(30.), np.float32(0.))


# ---- device implementations (GPU) ----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["equal", "shift", "decimate", "sinc_up", "sinc_down", "long"])
def test_time_resample_device_matches_oracle(case):
    import judi_jl_b200 as J
    rng = np.random.default_rng(7)
    cfg = {"equal": ((0., 2., 301), (0., 2., 301), 5),
           "shift": ((0., 2., 301), (10., 2., 296), 5),
           "decimate": ((0., 1., 1201), (4., 4., 300), 3),
           "sinc_up": ((0., 4., 251), (0., 1.733, 578), 17),
           "sinc_down": ((0., 1.921, 1042), (0., 4., 501), 9),
           "long": ((0., 0.75, 2700), (0., 2., 1013), 2)}[case]
    t_in, t_new, ntr = cfg
    d = rng.standard_normal((t_in[2], ntr)).astype(np.float32)
    ref = O.time_resample(d, t_in, t_new)
    out = J.time_resample(d, t_in, t_new)
    assert out.shape == ref.shape
    if case in ("equal", "shift", "decimate"):
        assert np.array_equal(out, ref)                       # pure data movement: bit-exact
    else:
        assert rel_l2(out, ref) < 1e-5                        # fp32 sums in a different order


@pytest.mark.gpu
def test_time_resample_device_tensor_path():
    import torch
    import judi_jl_b200 as J
    rng = np.random.default_rng(8)
    d = rng.standard_normal((400, 33)).astype(np.float32)
    t_in, t_new = (0., 3., 400), (0., 1.3, 921)
    ref = J.time_resample(d, t_in, t_new)
    dd = torch.from_numpy(np.ascontiguousarray(d.T)).cuda()
    out = J.time_resample(dd, t_in, t_new)
    assert out.is_cuda and tuple(out.shape) == (33, 921)
    assert np.array_equal(out.cpu().numpy().T, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("fold", [False, True])
def test_remove_padding_device_matches_oracle(ndim, fold):
    import torch
    import judi_jl_b200 as J
    args = model_args("iso", ndim, 8, with_dm=False, fs=(ndim == 2 and fold))
    gm = J.Model(**args)
    rng = np.random.default_rng(5)
    g = np.asfortranarray(rng.standard_normal(gm.shape_pml).astype(np.float32))
    ref = O.remove_padding(g.astype(np.float64), gm.padsizes, true_adjoint=fold)
    out = J.time_modeling.remove_padding(g, gm, true_adjoint=fold)
    assert out.shape == args["shape"]
    if not fold:
        assert np.array_equal(out, ref.astype(np.float32))    # crop: bit-exact
    else:
        assert rel_l2(out, ref) < 1e-6
        # adjoint of edge padding: <pad(x), g> == <x, remove_padding(g, true_adjoint)>
        x = rng.standard_normal(args["shape"])
        a = np.vdot(np.pad(x, gm.padsizes, mode="edge"), g.astype(np.float64))
        assert abs(a - np.vdot(x, out.astype(np.float64))) < 1e-5 * abs(a)
    gd = torch.from_numpy(np.ascontiguousarray(g.transpose(*range(ndim - 1, -1, -1)))).cuda()
    od = torch.empty(args["shape"][::-1], dtype=torch.float32, device="cuda")
    J.time_modeling.remove_padding(gd, gm, true_adjoint=fold, out=od)
    assert np.array_equal(od.cpu().numpy().transpose(*range(ndim - 1, -1, -1)), out)


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
def test_time_modeling_driver_with_resampling_and_limit_m(ndim):
    """_time_modeling (time_modeling_serial.jl:9-51): limit_m crop, wavelet resampled to dt_comp,
    shot record resampled to the receiver sampling, gradient cropped and embedded."""
    import judi_jl_b200 as J
    n = (81, 61) if ndim == 2 else (41, 37, 29)
    args = model_args("iso", ndim, 8, with_dm=True, n=n, nbl=10)
    spacing, shape = args["spacing"], args["shape"]
    ex = (shape[0] - 1) * spacing[0]
    if ndim == 2:
        sg = J.Geometry([0.5 * ex + 3.], None, [22.], dt=2.0, t=600.)
        rg = J.Geometry(np.linspace(0.3 * ex, 0.7 * ex, 21), None, np.full(21, 31.7), dt=3.0, t=600.)
    else:
        ey = (shape[1] - 1) * spacing[1]
        sg = J.Geometry([0.5 * ex + 3.], [0.5 * ey], [22.], dt=2.0, t=300.)
        xr, yr = np.meshgrid(np.linspace(0.3 * ex, 0.7 * ex, 5), np.linspace(0.4 * ey, 0.6 * ey, 3),
                             indexing="ij")
        rg = J.Geometry(xr.ravel(), yr.ravel(), np.full(xr.size, 31.7), dt=3.0, t=300.)
    q = O.ricker_wavelet(sg.t, sg.dt, 0.02)
    opt = J.Options(space_order=8, nbl=10, limit_m=True, buffer_size=60.)
    params = {"m": args["m"]}
    d = J.time_modeling.time_modeling(args["origin"], spacing, shape, params, sg, q, rg, None, None,
                                      "forward", opt)
    # the same composition with the oracle pieces
    sc, rc = sg.coords(ndim), rg.coords(ndim)
    o2, n2, p2, dm2, inds = O.limit_model_to_receiver_area(sc, rc, args["origin"], spacing, shape, 60.,
                                                           params, pert=args["dm"])
    assert n2[0] < shape[0]
    om = O.Model(o2, spacing, n2, space_order=8, nbl=10, m=p2["m"], dm=dm2, precision="f32")
    dtc = om.critical_dt
    nq = int(np.floor(np.float32(sg.t) / dtc)) + 1
    qi = O.time_resample(q, sg.taxis, (0., dtc, nq))
    do, _ = O.forward_rec(om, sc, qi, rc)
    dref = O.time_resample(do, (0., dtc, do.shape[0]), rg.taxis)
    assert d.shape == (rg.nt, rc.shape[0]) == dref.shape
    assert rel_l2(d, dref) < 1e-4
    # Born through the driver
    dl = J.time_modeling.time_modeling(args["origin"], spacing, shape, params, sg, q, rg, None,
                                       args["dm"], "born", opt)
    dlo, _ = O.born_rec(om, sc, qi, rc)
    assert rel_l2(dl, O.time_resample(dlo, (0., dtc, dlo.shape[0]), rg.taxis)) < 1e-4
    assert not J.time_modeling.time_modeling(args["origin"], spacing, shape, params, sg, q, rg, None,
                                             np.zeros(shape, np.float32), "born", opt).any()
    # adjoint Jacobian: gradient on the full grid, zero outside the limited area
    g = J.time_modeling.time_modeling(args["origin"], spacing, shape, params, sg, q, rg, dref, None,
                                      "adjoint_born", opt)
    di = O.time_resample(dref, rg.taxis, (0., dtc, nq))
    nt = min(qi.shape[0], di.shape[0])
    go, _, _ = O.J_adjoint(om, sc, qi[:nt], rc, di[:nt], is_residual=True)
    gref = np.zeros(shape, np.float32)
    gref[inds] = O.remove_padding(go, om.padsizes)
    assert g.shape == shape
    assert rel_l2(g, gref) < 1e-3


@pytest.mark.gpu
def test_fwi_objective_hbm_accumulation_matches_oracle_sum():
    """multi_src_fg! (propagation.jl:93-107): shots summed in HBM, cropped on the device."""
    import judi_jl_b200 as J
    args = model_args("iso", 2, 8, with_dm=False, n=(61, 51), nbl=12)
    gm = J.Model(**args)
    om = O.Model(precision="f32", **args)
    src0, rec = geometry(args)
    dt = float(om.critical_dt)
    q = O.ricker_wavelet(119 * dt, dt, 0.02)[:120]
    shots, fs, gs = [], 0.0, 0.0
    for i in range(3):
        src = src0.copy()
        src[0, 0] += 40.0 * (i - 1)
        d, _ = O.forward_rec(om, src, q, rec)
        dobs = (0.9 * d).astype(np.float32)
        shots.append((src, q, rec, dobs))
        f, g, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True)
        fs += float(f)
        gs = gs + g
    for fold in (False, True):
        f, g = J.objectives.fwi_objective(gm, shots, sum_padding=fold)
        gref = O.remove_padding(gs, om.padsizes, true_adjoint=fold)
        assert g.shape == args["shape"]
        assert abs(float(f) - fs) <= 1e-4 * abs(fs)
        assert rel_l2(g, gref) < 1e-3


@pytest.mark.gpu
def test_time_modeling_direction_comes_from_fw_alone():
    """_time_modeling: `op` selects the post-processing, the direction is `fw` (propagation.jl:9,
    52-54 -> python_interface.jl:42).  F' is called as op=:adjoint, fw=false."""
    import judi_jl_b200 as J
    args = model_args("iso", 2, 8, with_dm=False, n=(61, 51), nbl=12)
    om = O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dtc = float(om.critical_dt)
    nt = 150
    T = np.float32((nt - 1) * dtc)
    # y lives on the receiver geometry; F' injects it there and records at the source position
    rg = J.Geometry(rec[:, 0], None, rec[:, 1], dt=dtc, t=T)
    sg = J.Geometry(src[:, 0], None, src[:, 1], dt=dtc, t=T)
    y = np.random.default_rng(4).standard_normal((rg.nt, rec.shape[0])).astype(np.float32)
    opt = J.Options(space_order=8, nbl=12)
    params = {"m": args["m"]}
    qa = J.time_modeling.time_modeling(args["origin"], args["spacing"], args["shape"], params, rg, y, sg,
                                       None, None, "adjoint", opt, fw=False)
    qao, _ = O.forward_rec(om, rec, y[:rg.nt], src, fw=False)
    assert rel_l2(qa, qao[:qa.shape[0]]) < 1e-4
    qf = J.time_modeling.time_modeling(args["origin"], args["spacing"], args["shape"], params, rg, y, sg,
                                       None, None, "adjoint", opt, fw=True)
    assert rel_l2(qf, qao[:qf.shape[0]]) > 0.5       # the forward propagator is something else
    # the Jacobian of F' goes through (time mirror of the forward problem), it is NOT J of F
    dm = np.zeros(args["shape"], np.float32)
    dm[20:40, 15:35] = 0.02
    ql = O.ricker_wavelet(T, dtc, 0.02)
    dla = J.time_modeling.time_modeling(args["origin"], args["spacing"], args["shape"], params, sg, ql, rg,
                                        None, dm, "born", opt, fw=False)
    dlf = J.time_modeling.time_modeling(args["origin"], args["spacing"], args["shape"], params, sg, ql, rg,
                                        None, dm, "born", opt, fw=True)
    assert np.linalg.norm(dla) > 0 and rel_l2(dla, dlf) > 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ndim", [("iso", 2), ("tti", 2), ("iso", 3)])
def test_jacobian_of_the_adjoint_propagator(kind, ndim):
    """born_rec / J_adjoint with fw=False (`born(..., fw=False)`, `gradient(..., fw=False)`,
    propagators.py:98-230): J(F') dm against a finite difference of the tested F' path, and the dot
    test of the pair."""
    import judi_jl_b200 as J
    args = model_args(kind, ndim, 8)
    # one user time step below the critical one of all three models (models.py:476-481: the smaller wins)
    args["dt"] = np.float32(0.9 * float(J.Model(**args).critical_dt))
    gm = J.Model(**args)
    src, rec = geometry(args)
    dt = float(gm.critical_dt)
    assert dt <= float(args["dt"]) + 1e-6
    nt = 160
    # the adjoint propagator runs from the end of the time axis: a wavelet that fires late
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.02)[:nt][::-1].copy()
    dm = args["dm"]
    dl, _ = J.born_rec(gm, src, q, rec, fw=False)
    eps = 1e-2
    a0 = dict(args)
    a0.pop("dm")
    mp, mn = (J.Model(**dict(a0, m=(args["m"] + s_ * eps * dm).astype(np.float32))) for s_ in (1, -1))
    assert float(mp.critical_dt) == float(mn.critical_dt) == float(gm.critical_dt)
    dp, _ = J.forward_rec(mp, src, q, rec, fw=False)
    dmn, _ = J.forward_rec(mn, src, q, rec, fw=False)
    fd = (dp.astype(np.float64) - dmn) / (2 * eps)
    assert np.linalg.norm(fd) > 0
    assert rel_l2(dl, fd) < 2e-2                      # O(eps^2) + fp32 cancellation of the difference
    y = np.random.default_rng(3).standard_normal(dl.shape).astype(np.float32)
    g, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True, fw=False)
    dmp = np.pad(dm, gm.padsizes, mode="edge").astype(np.float64)
    c, e = dt * np.vdot(dl.astype(np.float64), y), np.vdot(g.astype(np.float64), dmp)
    assert abs((c - e) / (c + e)) < 1e-4
    with pytest.raises(NotImplementedError):
        J.J_adjoint(gm, src, q, rec, y, is_residual=True, fw=False, t_sub=7)


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
def test_lsrtm_objective_without_a_perturbation(ndim):
    """born_fwd with dm = None (propagators.py:181-194): the linearised record stays zero, so
    nlind=False gives residual = -dobs, f = dt/2 |dobs|^2, and nlind=True is the fwi objective."""
    import judi_jl_b200 as J
    args = model_args("iso", ndim, 8, with_dm=False)
    gm, om = J.Model(**args), O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dt = float(om.critical_dt)
    q = O.ricker_wavelet(119 * dt, dt, 0.02)[:120]
    d, _ = O.forward_rec(om, src, q, rec)
    dobs = (0.8 * d).astype(np.float32)
    f, g, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=False)
    fo, go, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=False)
    assert abs(float(fo) - 0.5 * dt * np.linalg.norm(dobs) ** 2) <= 1e-5 * abs(float(fo))
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, go) < 1e-3
    f2, g2, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True)
    f3, g3, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True)
    assert f2 == f3 and np.array_equal(g2, g3)
    # the multi-shot entry point (2-D: lockstep batch) follows the same rule
    fm, gmm = J.objectives.lsrtm_objective(gm, [(src, q, rec, dobs), (src, q, rec, dobs)], dm=None)
    assert abs(float(fm) - 2 * float(fo)) <= 1e-4 * abs(2 * float(fo))
    assert rel_l2(gmm, 2 * O.remove_padding(go, om.padsizes)) < 1e-3


@pytest.mark.gpu
def test_time_resample_rejects_an_axis_that_starts_before_the_data():
    import judi_jl_b200 as J
    a = np.ones((50, 3), np.float32)
    with pytest.raises(J.JudiB200Error, match="starts before"):
        J.time_modeling.time_resample(a, (8., 2., 50), (0., 2., 50))


@pytest.mark.gpu
def test_single_rank_communicator_and_trim():
    """judi_comm_init with one rank needs no NCCL; judi_fg_reduce then returns the local sums;
    judi_ctx_trim gives cached memory back and the context keeps working."""
    import ctypes
    import judi_jl_b200 as J
    from judi_jl_b200 import _lib as L
    ctx = J.get_context()
    assert J.objectives.init_comm(ctx) == 1
    assert L.lib().judi_comm_size(ctx.h) == 1 and L.lib().judi_comm_rank(ctx.h) == 0
    L.check(L.lib().judi_ctx_trim(ctx.h))
    args = model_args("iso", 2, 8, with_dm=False, n=(61, 51), nbl=12)
    gm, om = J.Model(**args), O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dt = float(om.critical_dt)
    q = O.ricker_wavelet(99 * dt, dt, 0.02)[:100]
    d, _ = O.forward_rec(om, src, q, rec)
    f, g = J.objectives.fwi_objective(gm, [(src, q, rec, (0.5 * d).astype(np.float32))])
    fo, go, _, _ = O.J_adjoint(om, src, q, rec, (0.5 * d).astype(np.float32), return_obj=True)
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, O.remove_padding(go, om.padsizes)) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
def test_fg_multishot_reduced_single_rank(ndim):
    """judi_fg_multishot_reduced (multi_src_fg! + reduce! + remove_padding in one call, the entry
    point of julia/B200Backend.jl): host arrays in, physical-grid gradient out; also with no shot
    at all on this rank (it still takes part in the reduction and returns zeros)."""
    import judi_jl_b200 as J
    args = model_args("iso", ndim, 8, with_dm=False)
    gm, om = J.Model(**args), O.Model(precision="f32", **args)
    J.objectives.init_comm(gm.ctx)
    src0, rec = geometry(args)
    dt = float(om.critical_dt)
    q = O.ricker_wavelet(99 * dt, dt, 0.02)[:100]
    shots, fs, gs = [], 0.0, 0.0
    for i in range(3):
        src = src0.copy()
        src[0, 0] += 30.0 * (i - 1)
        d, _ = O.forward_rec(om, src, q, rec)
        dobs = (0.9 * d).astype(np.float32)
        shots.append((src, q, rec, dobs))
        f, g, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True)
        fs += float(f)
        gs = gs + g
    for fold in (False, True):
        f, g = J.interface.fg_multishot_reduced(gm, shots, sum_padding=fold)
        assert g.shape == args["shape"]
        assert abs(float(f) - fs) <= 1e-4 * abs(fs)
        assert rel_l2(g, O.remove_padding(gs, om.padsizes, true_adjoint=fold)) < 1e-3
    f0, g0 = J.interface.fg_multishot_reduced(gm, [])
    assert float(f0) == 0.0 and not g0.any()


def test_maybe_pad_t0_cases():
    """_maybe_pad_t0 (time_utilities.jl:92-139): equal axes untouched, a pure shift pads both by
    the shift, different lengths pad to min(t0) .. max(t)."""
    import judi_jl_b200 as J
    TM = J.time_modeling
    G = lambda t0, t: TM.Geometry([0.], None, [0.], dt=2., t=t, t0=t0)
    q, d = np.ones((51, 1), np.float32), np.ones((51, 3), np.float32)
    a, b = TM.maybe_pad_t0(q, G(0, 100), d, G(0, 100), 2.)
    assert a.shape == (51, 1) and b.shape == (51, 3)
    a, b = TM.maybe_pad_t0(q, G(0, 100), d, G(20, 120), 2.)          # data recorded 20 ms late
    assert a.shape == (61, 1) and b.shape == (61, 3) and not b[:10].any() and not a[-10:].any()
    a, b = TM.maybe_pad_t0(q, G(20, 120), d, G(0, 100), 2.)          # source fires 20 ms late
    assert not a[:10].any() and not b[-10:].any() and a[10:].all() and b[:51].all()
    a, b = TM.maybe_pad_t0(q, G(0, 100), np.ones((31, 3), np.float32), G(20, 80), 2.)
    assert a.shape == (51, 1) and b.shape == (51, 3)
    assert not b[:10].any() and not b[41:].any() and b[10:41].all()
    assert TM.maybe_pad_t0(q, G(0, 100), None, G(0, 140), 2.).shape == (71, 1)


@pytest.mark.gpu
def test_multi_src_fg_driver_with_resampling_and_a_recording_delay():
    """_multi_src_fg (misfit_fg.jl:8-91) on geometries with their own time axes: 4 ms data that
    start 8 ms after the source does (time_resample + _maybe_pad_t0 + J_adjoint + remove_padding),
    against the same composition of the oracle's pieces."""
    import judi_jl_b200 as J
    TM = J.time_modeling
    args = model_args("iso", 2, 8, with_dm=False)
    om = O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dtc = np.float32(om.critical_dt)
    sg = TM.Geometry(src[:, 0], None, src[:, 1], dt=2.0, t=400.)
    rg = TM.Geometry(rec[:, 0], None, rec[:, 1], dt=4.0, t=408., t0=8.)
    q = O.ricker_wavelet(sg.t, sg.dt, 0.02)
    dobs = (0.1 * np.random.default_rng(8).standard_normal((rg.nt, rec.shape[0]))).astype(np.float32)
    opt = TM.Options(space_order=8, nbl=args["nbl"], sum_padding=True)
    f, g = TM.multi_src_fg(args["origin"], args["spacing"], args["shape"], {"m": args["m"]}, sg, q, rg, dobs,
                           options=opt)
    # the oracle's pieces
    nq = int(np.floor(np.float32(sg.t) / dtc)) + 1
    qi = O.time_resample(q, sg.taxis, (0., dtc, nq))
    nd = int(np.floor(np.float32(rg.t - rg.t0) / dtc)) + 1
    di = O.time_resample(dobs, rg.taxis, (rg.t0, dtc, nd))
    qi, di = TM.maybe_pad_t0(qi, sg, di, rg, dtc)
    assert qi.shape[0] == di.shape[0] and not di[0].any()
    fo, go, _, _ = O.J_adjoint(om, src, qi, rec, di, return_obj=True)
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo))
    assert rel_l2(g, O.remove_padding(go, om.padsizes, true_adjoint=True)) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
def test_time_modeling_with_extended_sources(ndim):
    """The judiWeights / judiLRWF methods of devito_interface (python_interface.jl:124-198): weights
    on the physical grid are zero-padded, the wavelet lives on the receiver geometry's time axis."""
    import judi_jl_b200 as J
    TM = J.time_modeling
    args = model_args("iso", ndim, 8)
    om = O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dtc = np.float32(om.critical_dt)
    nt = 130
    T = np.float32((nt - 1) * dtc)
    rg = TM.Geometry(rec[:, 0], rec[:, 1] if ndim == 3 else None, rec[:, -1], dt=dtc, t=T)
    assert rg.nt == nt
    q = O.ricker_wavelet(T, dtc, 0.02)[:nt]
    rng = np.random.default_rng(5)
    w = np.zeros(args["shape"], np.float32)
    box = tuple(slice(s // 2 - 4, s // 2 + 4) for s in args["shape"])
    w[box] = rng.standard_normal(w[box].shape).astype(np.float32)
    wpad = np.pad(w, om.padsizes, mode="constant")
    opt = TM.Options(space_order=8, nbl=args["nbl"])
    params = {"m": args["m"]}
    a = (args["origin"], args["spacing"], args["shape"], params)
    d = TM.time_modeling_w(*a, w, q, rg, None, None, "forward", opt)
    do, _ = O.forward_rec_w(om, wpad, q, rec)
    assert rel_l2(d, do) < 1e-4
    y = rng.standard_normal(d.shape).astype(np.float32)
    wa = TM.time_modeling_w(*a, None, q, rg, y, None, "adjoint", opt, fw=False)
    wao, _ = O.adjoint_w(om, rec, y, q, fw=False)
    assert wa.shape == args["shape"] and rel_l2(wa, O.remove_padding(wao, om.padsizes)) < 1e-4
    dl = TM.time_modeling_w(*a, w, q, rg, None, args["dm"], "born", opt)
    dlo, _ = O.born_rec_w(om, wpad, q, rec)
    assert rel_l2(dl, dlo) < 1e-4
    g = TM.time_modeling_w(*a, w, q, rg, y, None, "adjoint_born", opt)
    go, _, _ = O.J_adjoint(om, None, q, rec, y, is_residual=True, ws=wpad)
    assert rel_l2(g, O.remove_padding(go, om.padsizes)) < 1e-3


@pytest.mark.gpu
def test_time_modeling_with_wavefield_sources():
    """python_interface.jl:44-81: F Ps' q -> wavefield, Pr F u -> record, F u_in -> wavefield, and
    their consistency: sampling F Ps' q at the receivers reproduces Pr F Ps' q."""
    import judi_jl_b200 as J
    TM = J.time_modeling
    args = model_args("iso", 2, 8, with_dm=False)
    om = O.Model(precision="f32", **args)
    src, rec = geometry(args)
    dtc = np.float32(om.critical_dt)
    nt = 90
    T = np.float32((nt - 1) * dtc)
    sg = TM.Geometry(src[:, 0], None, src[:, 1], dt=dtc, t=T)
    rg = TM.Geometry(rec[:, 0], None, rec[:, 1], dt=dtc, t=T)
    q = O.ricker_wavelet(T, dtc, 0.02)[:nt]
    a = (args["origin"], args["spacing"], args["shape"], {"m": args["m"]})
    opt = TM.Options(space_order=8, nbl=args["nbl"])
    u, dt_u = TM.time_modeling_wf(*a, sg, q, None, opt)
    assert u.shape == (nt,) + om.shape_pml and dt_u == dtc
    _, uo, _ = O.forward(om, src, None, q, save=True)
    assert rel_l2(u[2:nt - 1], uo[2:nt - 1, 0]) < 1e-4
    usrc = np.zeros((nt,) + om.shape_pml, np.float32)
    usrc[:, 50:54, 40:44] = np.random.default_rng(2).standard_normal((nt, 4, 4)).astype(np.float32)
    d = TM.time_modeling_wf(*a, None, usrc, rg, opt)
    do, _ = O.forward_wf_src(om, usrc, rec)
    assert d.shape == (nt, rec.shape[0]) and rel_l2(d, do) < 1e-4
    u2, _ = TM.time_modeling_wf(*a, None, usrc, None, opt)
    u2o, _ = O.forward_wf_src_norec(om, usrc)
    assert rel_l2(u2[2:nt - 1], np.asarray(u2o)[2:nt - 1].reshape(u2[2:nt - 1].shape)) < 1e-4
