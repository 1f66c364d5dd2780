"""First-contact GPU check: runs every path against the oracle and prints a table (no asserts)."""
import sys, os, time, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import judi_jl_b200 as J
from oracle import oracle as O

RES = []
def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

def report(name, **kw):
    RES.append(dict(name=name, **kw))
    print("%-44s" % name, " ".join("%s=%s" % (k, ("%.3e" % v) if isinstance(v, float) else v) for k, v in kw.items()), flush=True)

def case(fn):
    def run(*a, **k):
        try:
            t0 = time.time(); fn(*a, **k); 
        except Exception as e:
            traceback.print_exc(); report(fn.__name__ + str(a), error=repr(e)[:200])
    return run

def mk2d(kind, so=8, n=(101, 81), nbl=20, with_dm=True):
    d = (10., 10.)
    v = np.ones(n, dtype=np.float32) * 1.5
    v[:, n[1] // 3:] = 2.0; v[:, 2 * n[1] // 3:] = 2.8
    m = (1 / v ** 2).astype(np.float32)
    kw = {}
    if kind == "rho": kw['rho'] = ((v + .5) / 2).astype(np.float32)
    if kind == "tti":
        kw.update(epsilon=((v - 1.5) / 12).astype(np.float32), delta=((v - 1.5) / 14).astype(np.float32), theta=((v - 1.5) / 4).astype(np.float32))
    if with_dm:
        dmv = np.zeros(n, dtype=np.float32); dmv[30:70, 25:60] = 0.03; dmv[45:55, 10:20] = -0.02
        kw['dm'] = dmv
    return dict(origin=(0., 0.), spacing=d, shape=n, space_order=so, nbl=nbl, m=m, **kw)

def mk3d(kind, so=8, n=(41, 37, 33), nbl=10, with_dm=True):
    d = (10., 10., 10.)
    v = np.ones(n, dtype=np.float32) * 1.5
    v[:, :, n[2] // 2:] = 2.5
    v += 0.1 * (1 + np.sin(np.arange(n[0]) / 5.0))[:, None, None].astype(np.float32)
    m = (1 / v ** 2).astype(np.float32)
    kw = {}
    if kind == "rho": kw['rho'] = ((v + .5) / 2).astype(np.float32)
    if kind == "tti":
        kw.update(epsilon=((v - 1.5) / 12).astype(np.float32), delta=((v - 1.5) / 14).astype(np.float32), theta=((v - 1.5) / 4).astype(np.float32), phi=((v - 1.5) / 3).astype(np.float32))
    if with_dm:
        dmv = np.zeros(n, dtype=np.float32); dmv[10:30, 8:28, 12:25] = 0.03
        kw['dm'] = dmv
    return dict(origin=(0., 0., 0.), spacing=d, shape=n, space_order=so, nbl=nbl, m=m, **kw)

def geom(args, nt):
    nd = len(args['shape'])
    ext = [(s - 1) * h for s, h in zip(args['shape'], args['spacing'])]
    if nd == 2:
        src = np.array([[ext[0] * 0.5 + 3.3, 22.]], dtype=np.float32)
        xr = np.linspace(15., ext[0] - 15., 31)
        rec = np.stack([xr, np.full_like(xr, 31.7)], axis=1).astype(np.float32)
    else:
        src = np.array([[ext[0] * 0.5 + 3.3, ext[1] * 0.45, 22.]], dtype=np.float32)
        xr, yr = np.meshgrid(np.linspace(15., ext[0] - 15., 7), np.linspace(12., ext[1] - 12., 5), indexing="ij")
        rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 31.7)], axis=1).astype(np.float32)
    return src, rec

@case
def check(kind, nd, so, nt=300, opt=None):
    args = mk2d(kind, so) if nd == 2 else mk3d(kind, so)
    ctx = J.get_context(0)
    for k, v in (opt or {}).items(): ctx.set_option(k, v)
    tag = "%dD-%s-so%d%s" % (nd, kind, so, "" if not opt else "-" + ",".join("%s%d" % kv for kv in opt.items()))
    gm = J.Model(**args)
    om = O.Model(precision="f32", **args); om64 = O.Model(precision="f64", **args)
    dt = float(gm.critical_dt)
    report(tag + " dt", gpu=dt, oracle=float(om.critical_dt), equal=bool(gm.critical_dt == om.critical_dt))
    report(tag + " damp", rel=rel(gm.get_field("damp"), om.damp), m=rel(gm.get_field("m"), om.m))
    src, rec = geom(args, nt)
    ijk, w, valid = gm.sparse_locate(rec); oijk, ow, ovalid = O.sparse_locate(om, rec)
    report(tag + " locate", ijk_equal=bool((ijk == oijk).all()), w_bitequal=bool((w.view(np.int32) == ow.view(np.int32)).all()), valid_equal=bool((valid == ovalid).all()))
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.02)[:nt]
    # forward
    t0 = time.time(); d, I = J.forward_rec(gm, src, q, rec, illum=True); tg = time.time() - t0
    do, Io = O.forward_rec(om, src, q, rec, illum=True); d64, _ = O.forward_rec(om64, src, q, rec)
    report(tag + " forward_rec", vs_f32=rel(d, do), vs_f64=rel(d, d64), oracle32_vs_f64=rel(do, d64), illum=rel(I, Io), norm=float(np.linalg.norm(d)), sec=tg)
    # adjoint + dot test
    y = np.random.default_rng(0).standard_normal(d.shape).astype(np.float32)
    qa, _ = J.forward_rec(gm, rec, y, src, fw=False)
    qao, _ = O.forward_rec(om, rec, y, src, fw=False)
    a = float(np.dot(d.ravel().astype(np.float64), y.ravel())); b = float(np.dot(q.ravel().astype(np.float64), qa.ravel()))
    report(tag + " adjoint", vs_f32=rel(qa, qao), dot=(a - b) / (a + b))
    # born
    dl, _ = J.born_rec(gm, src, q, rec); dlo, _ = O.born_rec(om, src, q, rec); dl64, _ = O.born_rec(om64, src, q, rec)
    report(tag + " born_rec", vs_f32=rel(dl, dlo), vs_f64=rel(dl, dl64), oracle32_vs_f64=rel(dlo, dl64), norm=float(np.linalg.norm(dl)))
    # J' (is_residual) + dot test
    g, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True)
    go, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True); g64, _, _ = O.J_adjoint(om64, src, q, rec, y, is_residual=True)
    c = dt * float(np.dot(dl.ravel().astype(np.float64), y.ravel())); e = float(np.dot(g.ravel().astype(np.float64), np.pad(args['dm'], gm.padsizes, mode='edge').ravel()))
    report(tag + " J_adjoint", vs_f32=rel(g, go), vs_f64=rel(g, g64), oracle32_vs_f64=rel(go, g64), dot=(c - e) / (c + e))
    # fwi objective: dobs from a perturbed model via the oracle data 'do' scaled
    dobs = (0.9 * do).astype(np.float32)
    f, g2, Iu, Iv = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, illum=True)
    fo, g2o, Iuo, Ivo = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, illum=True)
    report(tag + " fwi f,g", f=float(f), fo=float(fo), frel=abs(float(f) - float(fo)) / abs(float(fo)), g=rel(g2, g2o), Iu=rel(Iu, Iuo), Iv=rel(Iv, Ivo))
    for ts in (4,):
        f3, g3, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, t_sub=ts)
        f3o, g3o, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, t_sub=ts)
        report(tag + " fwi t_sub=%d" % ts, g=rel(g3, g3o), vs_tsub1=rel(g3, g2))
    f4, g4, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, checkpointing=True)
    report(tag + " checkpointing", g_vs_standard=rel(g4, g2), bitwise=bool((g4 == g2).all()), f_equal=bool(f4 == f))
    f5, g5, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, snapshot_region="physical")
    pads = gm.padsizes
    crop = tuple(slice(l, n - r) for (l, r), n in zip(pads, g5.shape))
    report(tag + " snap physical", g_crop=rel(g5[crop], g2[crop]), outside_zero=bool(np.abs(g5).sum() == np.abs(g5[crop]).sum()))
    # lsrtm: born_fwd, nlind
    f6, g6, _, _ = J.J_adjoint(gm, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True)
    f6o, g6o, _, _ = O.J_adjoint(om, src, q, rec, dobs, return_obj=True, born_fwd=True, nlind=True)
    report(tag + " lsrtm nlind", frel=abs(float(f6) - float(f6o)) / abs(float(f6o)), g=rel(g6, g6o))
    if kind == "iso":
        for ic in ("isic", "fwi"):
            dli, _ = J.born_rec(gm, src, q, rec, ic=ic); dlio, _ = O.born_rec(om, src, q, rec, ic=ic)
            gi, _, _ = J.J_adjoint(gm, src, q, rec, y, is_residual=True, ic=ic); gio, _, _ = O.J_adjoint(om, src, q, rec, y, is_residual=True, ic=ic)
            c = dt * float(np.dot(dli.ravel().astype(np.float64), y.ravel())); e = float(np.dot(gi.ravel().astype(np.float64), np.pad(args['dm'], gm.padsizes, mode='edge').ravel()))
            report(tag + " ic=" + ic, born=rel(dli, dlio), grad=rel(gi, gio), dot=(c - e) / (c + e))

@case
def speed3d(n=(401, 401, 201), nbl=40, nt=40):
    ctx = J.get_context(0)
    v = np.full(n, 2.0, dtype=np.float32); v[:, :, n[2] // 2:] = 3.0
    m = (1 / v ** 2).astype(np.float32)
    gm = J.Model((0., 0., 0.), (25., 25., 25.), n, space_order=8, nbl=nbl, m=m)
    dt = float(gm.critical_dt)
    src = np.array([[5000., 5000., 50.]], dtype=np.float32)
    xr, yr = np.meshgrid(np.linspace(100., 9900., 51), np.linspace(100., 9900., 51), indexing="ij")
    rec = np.stack([xr.ravel(), yr.ravel(), np.full(xr.size, 50.)], axis=1).astype(np.float32)
    q = O.ricker_wavelet((nt - 1) * dt, dt, 0.008)[:nt]
    N = np.prod(gm.shape_pml)
    ref = None
    for kern, tile, zc in [(0, 0, 0)] + [(1, t, 0) for t in range(8)] + [(1, t, z) for t in (0, 4, 7) for z in (3, 5, 8, 11)]:
        ctx.set_option("acoustic3d_kernel", kern); ctx.set_option("tile", tile); ctx.set_option("zchunks", zc); ctx.set_option("timing", 1)
        J.forward_rec(gm, src, q, rec)  # warm
        ctx.reset_stats()
        t0 = time.time(); d, _ = J.forward_rec(gm, src, q, rec); wall = time.time() - t0
        st = ctx.stencil_stats()
        ms = st["ms"] / st["launches"]
        if ref is None: ref = d
        report("3D C3 fwd kern=%d tile=%d zc=%d" % (kern, tile, zc), ms_per_step=ms, gpts=N / ms / 1e6, gbs=16 * N / ms / 1e6, frac=16 * N / ms / 1e6 / 6535.7, vs_generic=rel(d, ref), bitwise=bool((d == ref).all()), wall=wall)
    ctx.set_option("timing", 0); ctx.set_option("tile", 0); ctx.set_option("zchunks", 0)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "parity"):
        for nd in (2, 3):
            for kind in ("iso", "rho", "tti"):
                check(kind, nd, 8)
        check("iso", 2, 4); check("iso", 2, 16); check("iso", 3, 4, 200); check("iso", 3, 16, 200)
        check("iso", 3, 8, 300, {"acoustic3d_kernel": 0})
        J.get_context(0).set_option("acoustic3d_kernel", 1)
    if which in ("all", "speed", "quick"):
        if which == "quick":
            check("iso", 3, 8, 200); check("iso", 3, 4, 200); check("iso", 3, 16, 200); check("tti", 3, 8, 200)
        speed3d()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(RES, open("gpurun_out/gpu_check_%s.json" % which, "w"), indent=1)
