"""Edge cases of the path, through the C-ABI against the oracle: the shortest legal time axis,
sources / receivers outside the padded grid (Devito skips the out-of-grid corners),
no receivers at all, several sources per shot (simultaneous sources), and ragged multi-shot lists
(different receiver counts and record lengths per shot: judi_fg_multishot falls back to one call
per shot)."""
import numpy as np
import pytest

from conftest import model_args, geometry, rel_l2
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _pair(kind="iso", ndim=2, **kw):
    import judi_jl_b200 as J
    args = model_args(kind, ndim, 8, **kw)
    return J, args, J.Model(**args), O.Model(precision="f32", **args)


@pytest.mark.parametrize("ndim", [2, 3])
def test_shortest_time_axis_and_errors(ndim):
    J, args, gm, om = _pair(ndim=ndim)
    src, rec = geometry(args)
    q3 = np.array([[1.0], [0.5], [0.25]], np.float32)          # nt = 3: exactly one time step (t = 1)
    d, _ = J.forward_rec(gm, src, q3, rec)
    do, _ = O.forward_rec(om, src, q3, rec)
    assert d.shape == (3, rec.shape[0]) and np.array_equal(d, do.astype(np.float32)) or rel_l2(d, do) < 1e-6
    assert not d[0].any() and not d[2].any()
    with pytest.raises(J.JudiB200Error, match="at least 3"):
        J.forward_rec(gm, src, q3[:2], rec)
    with pytest.raises(J.JudiB200Error):
        J.J_adjoint(gm, src, q3, rec, np.zeros((3, rec.shape[0]), np.float32), t_sub=0)


@pytest.mark.parametrize("ndim", [2, 3])
def test_traces_outside_the_grid_and_simultaneous_sources(ndim):
    """Corners outside the padded grid are skipped (half-in receivers keep their inside corners),
    fully outside traces stay zero; two sources in one shot inject both wavelets."""
    J, args, gm, om = _pair(ndim=ndim)
    src, rec = geometry(args)
    dt = float(gm.critical_dt)
    nt = 140
    ext = np.array([(s - 1) * h for s, h in zip(args["shape"], args["spacing"])], np.float32)
    pad = np.float32(args["nbl"] * args["spacing"][0])
    far = rec[:3].copy()
    far[0, 0] = -pad - 55.                 # fully outside
    far[1, 0] = -pad - 4.                  # straddles the left edge of the padded grid
    far[2, 0] = ext[0] + pad + 3.          # straddles the right edge
    rec2 = np.concatenate([rec, far]).astype(np.float32)
    src2 = np.concatenate([src, src + np.float32(47.)]).astype(np.float32)
    q = np.concatenate([O.ricker_wavelet((nt - 1) * dt, dt, 0.02)[:nt],
                        -0.5 * np.roll(O.ricker_wavelet((nt - 1) * dt, dt, 0.03)[:nt], 7, axis=0)], axis=1)
    d, _ = J.forward_rec(gm, src2, q, rec2)
    do, _ = O.forward_rec(om, src2, q, rec2)
    assert rel_l2(d, do) < 1e-4
    assert not d[:, rec.shape[0]].any()                      # the fully outside trace
    y = np.random.default_rng(1).standard_normal(d.shape).astype(np.float32)
    qa, _ = J.forward_rec(gm, rec2, y, src2, fw=False)
    qao, _ = O.forward_rec(om, rec2, y, src2, fw=False)
    assert qa.shape == (nt, 2) and rel_l2(qa, qao) < 1e-4
    f, g, _, _ = J.J_adjoint(gm, src2, q, rec2, (0.5 * do).astype(np.float32), return_obj=True)
    fo, go, _, _ = O.J_adjoint(om, src2, q, rec2, (0.5 * do).astype(np.float32), return_obj=True)
    assert abs(float(f) - float(fo)) <= 1e-4 * abs(float(fo)) and rel_l2(g, go) < 1e-3


def test_forward_without_receivers_returns_the_wavefield_only():
    J, args, gm, om = _pair(ndim=2)
    src, rec = geometry(args)
    dt = float(gm.critical_dt)
    q = O.ricker_wavelet(59 * dt, dt, 0.02)[:60]
    d, I = J.forward_rec(gm, src, q, rec[:0], illum=True)
    assert d.shape == (60, 0)
    _, Io = O.forward_rec(om, src, q, rec, illum=True)
    assert rel_l2(I, Io) < 1e-4


def test_ragged_multishot_list_matches_the_sum_of_single_shots():
    """Shots with different receiver counts and record lengths in one fwi_objective call."""
    J, args, gm, om = _pair(ndim=2, with_dm=False)
    src, rec = geometry(args)
    dt = float(gm.critical_dt)
    shots, fs, gs = [], 0.0, 0.0
    for i, (nt, nr) in enumerate([(120, rec.shape[0]), (90, 7), (150, 19)]):
        q = O.ricker_wavelet((nt - 1) * dt, dt, 0.02)[:nt]
        s_i = src + np.float32(30. * i)
        d, _ = O.forward_rec(om, s_i, q, rec[:nr])
        dobs = (0.7 * d).astype(np.float32)
        shots.append((s_i, q, rec[:nr], dobs))
        f, g, _, _ = O.J_adjoint(om, s_i, q, rec[:nr], dobs, return_obj=True)
        fs += float(f)
        gs = gs + g
    f, g = J.objectives.fwi_objective(gm, shots)
    assert abs(float(f) - fs) <= 1e-4 * abs(fs)
    assert rel_l2(g, O.remove_padding(gs, om.padsizes)) < 1e-3


def test_so16_y_role_variant_is_bit_identical():
    """acoustic3d_tma at so = 16 with the y part of the Laplacian computed by the Y warps (ctx option
    tile = 34, profiles/r02_so16_yrole.md): same operations in the same order, so shot record and
    gradient must equal the default instantiation bit for bit (odd grid shape: partial tiles)."""
    import judi_jl_b200 as J
    args = model_args("iso", 3, 16, n=(75, 41, 38), nbl=12)
    src, rec = geometry(args)
    ctx = J.get_context(0)
    out = []
    try:
        for tile in (8, 34):
            ctx.set_option("tile", tile)
            gm = J.Model(**args)
            dt = float(gm.critical_dt)
            q = O.ricker_wavelet(89 * dt, dt, 0.03)[:90]
            d, _ = J.forward_rec(gm, src, q, rec)
            f, g, _, _ = J.J_adjoint(gm, src, q, rec, (0.5 * d).astype(np.float32), return_obj=True, t_sub=2)
            out.append((d, np.array(g), float(f)))
    finally:
        ctx.set_option("tile", 8)
    assert np.linalg.norm(out[0][0]) > 0
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]
    om = O.Model(precision="f32", **args)
    do, _ = O.forward_rec(om, src, q, rec)
    assert rel_l2(out[1][0], do) < 1e-4
