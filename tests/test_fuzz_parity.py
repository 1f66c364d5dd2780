"""Randomised parity sweep: the CUDA path through the C-ABI against the CPU oracle on seeded random
configurations the fixed-shape parity matrix (tests/test_gpu_parity.py) does not visit -- grid
shapes that are not multiples of any tile size (the TMA kernels' partial tiles and z-chunks, the
persistent 2-D kernels' partial rows), a different spacing on every axis (per-axis stencil weights
and damping profiles: models.py:74-94 divides by each dimension's spacing), a non-zero origin
(sparse positions, models.py:172), every layer width down to the stencil radius, random source /
receiver positions anywhere in the physical box, random t_sub, free surface, and all three space
orders for every propagator family of the north star.

Each case checks the four operators of the path at the north-star tolerances: shot record and
Born record <= 1e-4, adjoint source <= 1e-4, gradient <= 1e-3 (relative L2), and the objective
value <= 1e-4 relative.  The case list is a pure function of the seed, so a failure reproduces."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import oracle as O

NCASES_3D, NCASES_2D = 12, 12


def make_case(seed, ndim):
    rng = np.random.default_rng(1000 * ndim + seed)
    kind = ["iso", "rho", "tti", "ttirho"][seed % 4]
    so = [8, 4, 16][(seed // 4) % 3]
    r = so // 2
    if ndim == 3:
        n = tuple(int(v) for v in rng.integers(19, 58, size=3))
        nbl = int(rng.integers(r + 4, 15))
        nt = int(rng.integers(150, 210))
    else:
        n = tuple(int(v) for v in rng.integers(37, 150, size=2))
        nbl = int(rng.integers(r + 4, 30))
        nt = int(rng.integers(170, 280))
    d = tuple(float(np.float32(v)) for v in rng.uniform(6., 14., size=ndim))
    origin = tuple(float(np.float32(v)) for v in rng.uniform(-500., 500., size=ndim))
    fs = bool(rng.integers(0, 4) == 0)
    # smooth random velocity: a gradient in depth + two random Gaussian blobs
    ax = [np.arange(k, dtype=np.float32) / np.float32(k) for k in n]
    X = np.meshgrid(*ax, indexing="ij")
    v = 1.5 + 1.2 * X[-1]
    for _ in range(2):
        c = rng.uniform(0.2, 0.8, size=ndim)
        w = rng.uniform(0.08, 0.3)
        v = v + rng.uniform(-0.3, 0.4) * np.exp(-sum((x - ci) ** 2 for x, ci in zip(X, c)) / w ** 2)
    v = v.astype(np.float32)
    kw = {}
    if kind in ("rho", "ttirho"):
        kw["rho"] = (0.31 * (1000. * v) ** 0.25).astype(np.float32)        # Gardner
    if kind in ("tti", "ttirho"):
        # epsilon >= delta everywhere: sqrt((eps~ - delta~) delta~) of thomsen_mat (FD_utils.py:62-81)
        kw["epsilon"] = (0.04 + 0.18 * X[0] * X[-1]).astype(np.float32)
        kw["delta"] = (0.45 * kw["epsilon"] * (0.5 + 0.5 * X[-1])).astype(np.float32)
        kw["theta"] = (0.6 * (X[0] - 0.5)).astype(np.float32)
        if ndim == 3:
            kw["phi"] = (0.8 * (X[1] - 0.3)).astype(np.float32)
    dm = np.zeros(n, np.float32)
    lo = [int(k * 0.25) for k in n]
    hi = [int(k * 0.8) for k in n]
    dm[tuple(slice(a, b) for a, b in zip(lo, hi))] = 0.02
    dm += (0.01 * np.sin(7 * X[0]) * X[-1]).astype(np.float32)
    args = dict(origin=origin, spacing=d, shape=n, space_order=so, nbl=nbl,
                m=(1 / v ** 2).astype(np.float32), dm=dm, **kw)
    if fs:
        args["fs"] = True
    ext = np.array([(k - 1) * h for k, h in zip(n, d)], np.float32)
    org = np.array(origin, np.float32)
    nsrc = int(rng.integers(1, 3))
    nrec = int(rng.integers(5, 40))
    src = (org + rng.uniform(0.05, 0.95, size=(nsrc, ndim)).astype(np.float32) * ext).astype(np.float32)
    src[:, -1] = org[-1] + np.float32(rng.uniform(0.5, 3.5)) * np.float32(d[-1])
    # receivers within 14 (3-D) / 30 (2-D) cells of the first source, so that the record holds the
    # direct wave and not only its numerical precursor, clipped to the physical box
    reach = (14. if ndim == 3 else 30.) * np.array(d, np.float32)
    rec = src[0] + rng.uniform(-1.0, 1.0, size=(nrec, ndim)).astype(np.float32) * reach
    rec = np.minimum(np.maximum(rec, org), org + ext).astype(np.float32)
    rec[:, -1] = org[-1] + rng.uniform(0.5, 6.0, size=nrec).astype(np.float32) * np.float32(d[-1])
    t_sub = int(rng.integers(1, 5))
    return dict(kind=kind, so=so, args=args, src=src, rec=rec, nt=nt, t_sub=t_sub, nsrc=nsrc,
                ic=["as", "as", "isic", "fwi"][int(rng.integers(0, 4))])


def oracle_side(case):
    om = O.Model(precision="f32", **case["args"])
    dt = float(om.critical_dt)
    nt = case["nt"]
    q1 = O.ricker_wavelet((nt - 1) * dt, dt, 0.04)[:nt]
    q = np.concatenate([q1 * (1 - 0.4 * k) for k in range(case["nsrc"])], axis=1).astype(np.float32)
    d, _ = O.forward_rec(om, case["src"], q, case["rec"])
    dl, _ = O.born_rec(om, case["src"], q, case["rec"], ic=case["ic"])
    y = np.random.default_rng(5).standard_normal(d.shape).astype(np.float32)
    qa, _ = O.forward_rec(om, case["rec"], y, case["src"], fw=False)
    dobs = (0.8 * d + 0.05 * np.abs(d).max() * y).astype(np.float32)
    f, g, _, _ = O.J_adjoint(om, case["src"], q, case["rec"], dobs, return_obj=True,
                             t_sub=case["t_sub"], ic=case["ic"])
    return om, q, y, dobs, dict(d=d, dl=dl, qa=qa, f=float(f), g=g)


IDS_3D = ["3d-%02d" % s for s in range(NCASES_3D)]
IDS_2D = ["2d-%02d" % s for s in range(NCASES_2D)]


def test_case_generator_is_deterministic_and_valid():
    """CPU: the list is reproducible, covers every family and order, and every case is a legal
    model for the oracle (positive fields, sources inside the physical box)."""
    seen = set()
    for ndim, ncase in ((3, NCASES_3D), (2, NCASES_2D)):
        for s in range(ncase):
            a, b = make_case(s, ndim), make_case(s, ndim)
            assert a["args"]["shape"] == b["args"]["shape"] and np.array_equal(a["src"], b["src"])
            seen.add((ndim, a["kind"], a["so"]))
            assert a["args"]["m"].min() > 0 and a["args"]["nbl"] >= a["so"] // 2 + 4
    for ndim in (2, 3):
        for k in ("iso", "rho", "tti", "ttirho"):
            assert any(x[0] == ndim and x[1] == k for x in seen)
        assert {x[2] for x in seen if x[0] == ndim} == {4, 8, 16}
    # one small case end to end on the oracle: finite, non-trivial outputs
    om, q, y, dobs, ref = oracle_side(make_case(1, 2))
    assert np.isfinite(ref["g"]).all() and np.linalg.norm(ref["d"]) > 0 and ref["f"] > 0


def _run(case):
    import judi_jl_b200 as J
    om, q, y, dobs, ref = oracle_side(case)
    gm = J.Model(**case["args"])
    assert gm.critical_dt == om.critical_dt and gm.shape_pml == om.shape_pml
    ijk, w, valid = gm.sparse_locate(case["rec"])
    oijk, ow, ovalid = O.sparse_locate(om, case["rec"])
    assert np.array_equal(ijk, oijk) and np.array_equal(w.view(np.int32), ow.view(np.int32))
    d, _ = J.forward_rec(gm, case["src"], q, case["rec"])
    assert rel_l2(d, ref["d"]) < 1e-4, ("forward", rel_l2(d, ref["d"]))
    dl, _ = J.born_rec(gm, case["src"], q, case["rec"], ic=case["ic"])
    assert rel_l2(dl, ref["dl"]) < 1e-4, ("born", rel_l2(dl, ref["dl"]))
    qa, _ = J.forward_rec(gm, case["rec"], y, case["src"], fw=False)
    assert rel_l2(qa, ref["qa"]) < 1e-4, ("adjoint", rel_l2(qa, ref["qa"]))
    f, g, _, _ = J.J_adjoint(gm, case["src"], q, case["rec"], dobs, return_obj=True,
                             t_sub=case["t_sub"], ic=case["ic"])
    assert abs(float(f) - ref["f"]) <= 1e-4 * abs(ref["f"]), ("objective", float(f), ref["f"])
    assert rel_l2(g, ref["g"]) < 1e-3, ("gradient", rel_l2(g, ref["g"]))
    # the physical-domain snapshot region gives the same gradient inside the physical box
    if case["ic"] == "as":
        f2, g2, _, _ = J.J_adjoint(gm, case["src"], q, case["rec"], dobs, return_obj=True,
                                   t_sub=case["t_sub"], snapshot_region="physical")
        crop = tuple(slice(a, s - b) for (a, b), s in zip(gm.padsizes, gm.shape_pml))
        assert rel_l2(np.asarray(g2)[crop], np.asarray(ref["g"])[crop]) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(NCASES_3D), ids=IDS_3D)
def test_fuzz_3d(seed):
    _run(make_case(seed, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(NCASES_2D), ids=IDS_2D)
def test_fuzz_2d(seed):
    _run(make_case(seed, 2))
