"""Adjoint dot tests, case by case: the library (GPU) next to the fp32 and fp64 oracle.

Two recipes:
  ref    the reference's own (test/test_adjoint.jl:67-76, src/pysource/adjoint_test_F.py:75-88):
         y = F q, q_rand = (.5 + rand) .* q;  <F q_rand, y> vs <q_rand, F' y>;  J: <J dm, y> vs <dm, J' y>
  noise  y = white noise (what tests/test_gpu_parity.py used in round 1): the inner product
         <F q, y> cancels almost completely, kappa = |Fq||y| / |<Fq,y>| amplifies every rounding
         error of the fields -- err / kappa is the comparable number.
Prints one line per case; writes gpurun_out/dot_report.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import model_args, geometry
from oracle import oracle as O

CASES = [(k, nd, so) for k in ("iso", "rho", "tti") for nd in (2, 3) for so in (4, 8, 16)] + \
        [("visco", 2, 8), ("viscoc", 3, 8)]
NT = 260


def rerr(a, b):
    return abs((a - b) / (a + b))


def dots(F, Ft, Jf, Jt, q, dm_pad, dt, rng):
    """returns dict of the four dot errors + kappa of the noise recipe"""
    out = {}
    d = F(q)
    qr = ((0.5 + rng.random(q.shape, dtype=np.float32)) * q).astype(np.float32)
    y = d
    a, b = np.vdot(F(qr).astype(np.float64), y), np.vdot(qr.astype(np.float64), Ft(y))
    out["F_ref"] = rerr(a, b)
    dl = Jf(qr)
    c, e = dt * np.vdot(dl.astype(np.float64), y), np.vdot(Jt(qr, y).astype(np.float64), dm_pad)
    out["J_ref"] = rerr(c, e)
    yn = np.random.default_rng(0).standard_normal(d.shape).astype(np.float32)
    a, b = np.vdot(d.astype(np.float64), yn), np.vdot(q.astype(np.float64), Ft(yn))
    out["F_noise"] = rerr(a, b)
    out["F_kappa"] = float(np.linalg.norm(d) * np.linalg.norm(yn) / abs(a))
    dl = Jf(q)
    c, e = dt * np.vdot(dl.astype(np.float64), yn), np.vdot(Jt(q, yn).astype(np.float64), dm_pad)
    out["J_noise"] = rerr(c, e)
    out["J_kappa"] = float(dt * np.linalg.norm(dl) * np.linalg.norm(yn) / abs(c))
    return out


def api(mod, model, src, rec):
    F = lambda q: mod.forward_rec(model, src, q, rec)[0]
    Ft = lambda y: mod.forward_rec(model, rec, y, src, fw=False)[0]
    Jf = lambda q: mod.born_rec(model, src, q, rec)[0]
    Jt = lambda q, y: mod.J_adjoint(model, src, q, rec, y, is_residual=True)[0]
    return F, Ft, Jf, Jt


def main():
    import judi_jl_b200 as J
    O.set_threads(os.cpu_count() or 1)
    res = []
    print("%-12s %-5s | %-9s %-9s | %-9s %-9s | %-9s (k=%s) %-9s (k)" %
          ("case", "impl", "F ref", "J ref", "", "", "F noise", "", "J noise"))
    for kind, nd, so in CASES:
        args = model_args(kind, nd, so)
        src, rec = geometry(args)
        row = {"case": "%s-%dd-so%d" % (kind, nd, so)}
        for impl in ("gpu", "f32", "f64"):
            if impl == "gpu":
                model, mod = J.Model(**args), J
            else:
                model, mod = O.Model(precision=impl, **args), O
            dt = float(model.critical_dt)
            q = O.ricker_wavelet((NT - 1) * dt, dt, 0.02)[:NT]
            dm_pad = np.pad(args["dm"], model.padsizes, mode="edge").astype(np.float64)
            r = dots(*api(mod, model, src, rec), q, dm_pad, dt, np.random.default_rng(11))
            row[impl] = r
            print("%-12s %-5s | %.2e  %.2e | %.2e (k=%6.0f)  %.2e (k=%6.0f)"
                  % (row["case"], impl, r["F_ref"], r["J_ref"], r["F_noise"], r["F_kappa"],
                     r["J_noise"], r["J_kappa"]), flush=True)
        res.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dot_report.json"), "w"), indent=1)
    worst = {k: max(r["gpu"][k] for r in res) for k in ("F_ref", "J_ref", "F_noise", "J_noise")}
    print("worst GPU:", worst)


if __name__ == "__main__":
    main()
