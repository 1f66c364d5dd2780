import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import oracle as O
    O.build()


# ---- shared model / geometry recipes (value ranges follow the reference tests:
# test/seismic_utils.jl:23-63, src/pysource/adjoint_test_F.py:28-54) ---------------------------
def model_args(kind="iso", ndim=2, so=8, with_dm=True, n=None, nbl=None, fs=False):
    if ndim == 2:
        n = n or (101, 81)
        nbl = 20 if nbl is None else nbl
        d = (10., 10.)
        v = np.ones(n, dtype=np.float32) * 1.5
        v[:, n[1] // 3:] = 2.0
        v[:, 2 * n[1] // 3:] = 2.8
    else:
        n = n or (41, 37, 33)
        nbl = 10 if nbl is None else nbl
        d = (10., 10., 10.)
        v = np.ones(n, dtype=np.float32) * 1.5
        v[:, :, n[2] // 2:] = 2.5
        v += 0.1 * (1 + np.sin(np.arange(n[0]) / 5.0)).astype(np.float32).reshape((-1, 1, 1))
    m = (1 / v ** 2).astype(np.float32)
    kw = {}
    if kind in ("rho", "visco", "ttirho"):
        kw["rho"] = ((v + .5) / 2).astype(np.float32)
    if kind in ("visco", "viscoc"):
        # test/seismic_utils.jl:50-53: qp0 = 3.516 (1000 v)^2.2 1e-6, Model(n, d, o, m, rho0, qp0);
        # "viscoc" is the same with a constant density (the b * p.laplace branch of FD_utils.py:18)
        kw["qp"] = (3.516 * ((v * 1000.) ** 2.2) * 1e-6).astype(np.float32)
        kw["f0"] = 0.02
    if kind in ("tti", "ttirho"):
        kw.update(epsilon=((v - 1.5) / 12).astype(np.float32),
                  delta=((v - 1.5) / 14).astype(np.float32),
                  theta=((v - 1.5) / 4).astype(np.float32))
        if ndim == 3:
            kw["phi"] = ((v - 1.5) / 3).astype(np.float32)
    if with_dm:
        dm = np.zeros(n, dtype=np.float32)
        if ndim == 2:
            dm[30:70, 25:60] = 0.03
            dm[45:55, 10:20] = -0.02
        else:
            dm[10:30, 8:28, 12:25] = 0.03
        kw["dm"] = dm
    if fs:
        kw["fs"] = True
    return dict(origin=(0.,) * ndim, spacing=d, shape=n, space_order=so, nbl=nbl, m=m, **kw)


def geometry(args):
    nd = len(args["shape"])
    ext = [(s - 1) * h for s, h in zip(args["shape"], args["spacing"])]
    if nd == 2:
        src = np.array([[ext[0] * 0.5 + 3.3, 22.]], dtype=np.float32)
        xr = np.linspace(15., ext[0] - 15., 31)
        rec = np.stack([xr, np.full_like(xr, 31.7)], axis=1).astype(np.float32)
    else:
        src = np.array([[ext[0] * 0.5 + 3.3, ext[1] * 0.45, 22.]], dtype=np.float32)
        xr, yr = np.meshgrid(np.linspace(15., ext[0] - 15., 7),
                             np.linspace(12., ext[1] - 12., 5), indexing="ij")
        rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 31.7)], axis=1).astype(np.float32)
    return src, rec


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
