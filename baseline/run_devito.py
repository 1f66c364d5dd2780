"""Run the REFERENCE's own Devito path (src/pysource, no Julia) on the deterministic inputs of
tests/conftest.py and bench.py, to (a) pin the CPU oracle to Devito's numbers and (b) time the
reference CPU path for bench.py's workload.  It cannot run in the build container (Devito is not
installable offline); it is what SURVEY.md section 8(c)/(d) asks to ship for a
box that has `pip install devito`:

    export DEVITO_LANGUAGE=openmp OMP_NUM_THREADS=$(nproc)
    python baseline/run_devito.py --pysource /path/to/JUDI.jl/src/pysource parity  --out devito_golden.npz
    python baseline/run_devito.py --pysource ... bench --tn 200          # Gpt/s of the reference CPU path
    python baseline/run_devito.py --pysource ... devito-examples         # Devito's own example constants

`parity` writes, for every physics x dimension of the test matrix, the shot record, the Born data
and the gradient of src/pysource/interface.py on the SAME arrays the GPU tests use; if the oracle is
importable (this repository on PYTHONPATH) it also prints the relative L2 residuals oracle vs Devito
against the north-star tolerances (shot records 1e-4, gradients 1e-3), which is the number that
turns "stencil parity unpinned" into a measured statement.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ref_model(pm, args):
    """conftest.model_args(...) -> pysource.models.Model (same keywords by construction)."""
    kw = dict(args)
    kw.pop("f0", None)
    return pm.Model(dtype=np.float32, **kw)


def parity(ns):
    import models as pm            # the reference's modules (from --pysource)
    import interface as ri
    from conftest import model_args, geometry
    try:
        from oracle import oracle as O
    except Exception:              # the script also runs without this repository's oracle
        O = None
    out = {}
    nt = 260
    for kind in ("iso", "rho", "tti", "visco"):
        for ndim in (2, 3):
            args = model_args(kind, ndim, 8)
            f0 = args.get("f0", 0.015)
            model = ref_model(pm, args)
            src, rec = geometry(args)
            dt = float(model.critical_dt)
            t = np.arange(nt) * dt
            r = np.pi * 0.02 * (t - 1 / 0.02)
            q = ((1 - 2 * r ** 2) * np.exp(-r ** 2)).astype(np.float32).reshape(nt, 1)
            d, _ = ri.forward_rec(model, src, q, rec, f0=f0)
            dl, _ = ri.born_rec(model, src, q, rec, f0=f0)
            y = np.random.default_rng(0).standard_normal(d.shape).astype(np.float32)
            g = ri.J_adjoint(model, src, q, rec, y, is_residual=True, f0=f0)[0]
            key = "%s%dd" % (kind, ndim)
            out[key + "_dt"] = np.float32(dt)
            out[key + "_rec"] = np.array(d)
            out[key + "_born"] = np.array(dl)
            out[key + "_grad"] = np.array(g)
            if O is not None:
                om = O.Model(precision="f32", **args)
                qo = O.ricker_wavelet((nt - 1) * dt, dt, 0.02)[:nt]
                assert np.allclose(qo, q, atol=1e-6), "wavelets differ"
                rl2 = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
                do, _ = O.forward_rec(om, src, q, rec)
                dlo, _ = O.born_rec(om, src, q, rec)
                go, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True)
                print("%-8s dt %.4f/%.4f  oracle vs Devito: rec %.2e (1e-4)  born %.2e (1e-4)  grad %.2e (1e-3)"
                      % (key, float(om.critical_dt), dt, rl2(do, np.array(d)), rl2(dlo, np.array(dl)),
                         rl2(go, np.array(g))), flush=True)
    np.savez_compressed(ns.out, **out)
    print("wrote", ns.out)


def bench(ns):
    import models as pm
    import interface as ri
    import bench as B              # this repository's bench.py: the same synthetic workload
    shape = tuple(ns.shape)
    v0 = B.smooth_model(B.velocity(shape))
    model = pm.Model(origin=(0., 0., 0.), spacing=B.SPACING, shape=shape, space_order=B.SO, nbl=B.NBL,
                     m=(1 / v0 ** 2).astype(np.float32), dt=B.DT, dtype=np.float32)
    nt = int(np.floor(ns.tn / B.DT)) + 1
    src, rec = B.geometry(shape, 0)
    q = B.ricker(nt, B.DT, B.F0)
    dobs = np.zeros((nt, rec.shape[0]), dtype=np.float32)
    ri.J_adjoint(model, src, q, rec, dobs, t_sub=B.T_SUB, return_obj=True)     # JIT compile, excluded
    t0 = time.perf_counter()
    ri.J_adjoint(model, src, q, rec, dobs, t_sub=B.T_SUB, return_obj=True)
    wall = time.perf_counter() - t0
    npts = float(np.prod(np.array(shape) + 2 * B.NBL))
    print("Devito %s: FWI gradient shot %dx%dx%d nt=%d in %.2f s = %.2f Gpt/s on %s threads"
          % (os.environ.get("DEVITO_LANGUAGE", "C"), shape[0], shape[1], shape[2], nt, wall,
             2 * (nt - 2) * npts / wall / 1e9, os.environ.get("OMP_NUM_THREADS", "?")))


def devito_examples(ns):
    """Re-produce, with Devito itself, the four constants tests/test_devito_known_answer.py chains the
    oracle and the CUDA library to.  They are quoted from Devito's own assertions
    (examples/seismic/acoustic/acoustic_example.py::test_isoacoustic,
    examples/seismic/viscoacoustic/viscoacoustic_example.py::test_viscoacoustic) from memory -- this
    command is how a box that has Devito (with its `examples` package importable) confirms them."""
    from devito import norm
    from examples.seismic.acoustic.acoustic_example import run as run_acoustic
    from examples.seismic.viscoacoustic.viscoacoustic_example import run as run_visco
    want = {"acoustic": 459.1678, "acoustic, free surface": 369.955,
            "visco-acoustic sls": 684.385, "visco-acoustic kv": 677.673}
    got = {}
    got["acoustic"] = float(norm(run_acoustic(fs=False, dtype=np.float32)[3][0]))
    got["acoustic, free surface"] = float(norm(run_acoustic(fs=True, dtype=np.float32)[3][0]))
    got["visco-acoustic sls"] = float(norm(run_visco(kernel="sls", time_order=2, dtype=np.float32)[3][0]))
    got["visco-acoustic kv"] = float(norm(run_visco(kernel="kv", time_order=2, dtype=np.float32)[3][0]))
    for k in want:
        print("%-24s Devito now %.4f, quoted in this repository %.4f (%+.1e)"
              % (k, got[k], want[k], got[k] / want[k] - 1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--pysource", required=True, help="path of the reference's src/pysource")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p1 = sub.add_parser("parity")
    p1.add_argument("--out", default="devito_golden.npz")
    p2 = sub.add_parser("bench")
    p2.add_argument("--shape", type=int, nargs=3, default=[401, 401, 201])
    p2.add_argument("--tn", type=float, default=200.)
    sub.add_parser("devito-examples")
    ns = ap.parse_args()
    sys.path.insert(0, ns.pysource)
    try:
        import devito  # noqa: F401
    except ImportError:
        sys.exit("devito is not importable here: run this script on a box with `pip install devito`")
    {"parity": parity, "bench": bench, "devito-examples": devito_examples}[ns.cmd](ns)
