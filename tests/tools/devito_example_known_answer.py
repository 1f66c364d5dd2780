"""A known answer produced by Devito itself, restated in plain numpy with the conventions the
oracle ASSUMES for Devito's lowering (the `[dv]` items).

Devito (the third-party dependency the reference's arithmetic lives in; absent from this image)
pins its own acoustic example in its test-suite: `examples/seismic/acoustic/acoustic_example.py`
`run()` with the defaults shape = (50, 50, 50), spacing = 20 m, tn = 1000 ms, space_order = 4,
nbl = 40, preset 'layers-isotropic' (3 layers, 1.5 ... 3.5 km/s), kernel OT2, and
its `test_isoacoustic` (same file) asserts `norm(rec) == 459.1678` (rtol 1e-3) -- recalled from
Devito's public repository, NOT present under /root/reference, so this script is evidence about
the `[dv]` assumptions, not a parity test of the reference's own files.

What the experiment exercises is exactly the list of Devito behaviours the oracle could not read
in the reference's Python: the centred `.laplace` weights, the one-sided `u.dt` in
`solve(m u.dt2 - u.laplace + damp u.dt, u.forward)`, the time-loop bounds 1 .. nt-2, injection of
`src dt^2 / m` into the new level with (bi/tri)linear weights, interpolation of the current level,
and the damping profile Devito's own example Model builds (layer of nbl cells).

Result: 459.339 (+0.037 %) and, with the free surface, 370.217 (+0.071 %).  The residual is inside
the spread of the fp32 accumulation of Devito's `norm` builtin (it sums rec^2 in the record's own
dtype): a sequential fp32 sum of the SAME numpy record gives 459.063 / 369.738, two partial sums
459.268 / 370.107 -- Devito's constants lie between the two in both cases.

    python tests/tools/devito_example_known_answer.py
"""
import itertools
import numpy as np

DEVITO_NORM_REC = 459.1678


def cfl_sympy(space_order, ndim):
    from sympy import finite_diff_weights as fd_w
    c = fd_w(2, range(-space_order, space_order + 1), 0)[-1][-1]
    return float(np.sqrt(4.0 / (ndim * sum(abs(float(x)) for x in c))))


def run(dt_rule="sympy", udt="forward", tmax_rule="nt-2", shape=(50, 50, 50), h=20.0, tn=1000.0, so=4,
        nbl=40, f0=0.010, dtype=np.float64, rec_level="current", fs=False, layer="devito", dt=None,
        full=False, mirror="devito"):
    """-> (norm(rec), dt, nt); with full=True also the record (nt, nx*ny), the wavelet, the velocity
    model and the source / receiver coordinates in metres (for running the same problem elsewhere)."""
    from sympy import finite_diff_weights as fd_w
    r = so // 2
    w = [float(x) for x in fd_w(2, list(range(-r, r + 1)), 0)[2][-1]]   # centred 2nd derivative
    nl = 3
    v = np.full(shape, 1.5)
    vi = np.linspace(1.5, 3.5, nl)
    for i in range(1, nl):
        v[..., i * int(shape[-1] / nl):] = vi[i]
    ztop = 0 if fs else nbl
    vp = np.pad(v, [(nbl, nbl), (nbl, nbl), (ztop, nbl)], mode="edge")
    m = (1.0 / vp ** 2).astype(dtype)
    coeff = {"0.38": 0.38, "sympy": cfl_sympy(so, 3)}[dt_rule]
    if dt is None:
        dt = float(np.float32("%.3e" % (coeff * h / vp.max())))
    nt = int(np.ceil((tn + dt) / dt))
    t = np.linspace(0.0, dt * (nt - 1), nt)
    rr = np.pi * f0 * (t - 1.0 / f0)
    q = (1 - 2.0 * rr ** 2) * np.exp(-rr ** 2)
    # damping layer of Devito's example Model: nbl cells, coefficient 1.5 ln(1000) / nbl, / spacing
    # (layer="judi": the reference's own profile, models.py:74-94: nbl - 3 cells)
    n = shape[0] + 2 * nbl
    nz = shape[2] + nbl + ztop
    wl = nbl if layer == "devito" else nbl - 3
    i = np.arange(wl, dtype=np.float64)
    pos = np.abs((wl - i + 1) / float(wl))
    prof = 1.5 * np.log(1000.0) / wl * (pos - np.sin(2 * np.pi * pos) / (2 * np.pi)) / h
    damp = np.zeros((n, n, nz))
    for ax in range(3):
        shp = [1, 1, 1]
        shp[ax] = wl
        sl = [slice(None)] * 3
        if not (fs and ax == 2):
            sl[ax] = slice(0, wl)
            damp[tuple(sl)] += prof.reshape(shp)
        sl[ax] = slice(damp.shape[ax] - wl, damp.shape[ax])
        damp[tuple(sl)] += prof[::-1].reshape(shp)
    damp = damp.astype(dtype)
    # geometry: source at the centre in x, y and one cell deep; receivers on every node, two cells deep
    dom = h * (shape[0] - 1)
    sx = (0.5 * dom) / h + nbl
    ix = int(np.floor(sx))
    ax_ = sx - ix
    sz = 1 + ztop
    src_cells = [(ix + a, ix + b, sz, (ax_ if a else 1 - ax_) * (ax_ if b else 1 - ax_))
                 for a, b in itertools.product((0, 1), (0, 1))]
    rz = 2 + ztop
    u = np.zeros((n, n, nz), dtype)
    up = np.zeros_like(u)
    rec = np.zeros((nt, shape[0], shape[1]), dtype)
    a_, b_ = m / dt ** 2, damp / dt
    tM = {"nt-2": nt - 2, "nt-1": nt - 1}[tmax_rule]

    # free surface.  mirror="devito": Devito's example, antisymmetric about the surface row,
    # u(-k) = -u(k), u(0) = 0.  mirror="judi": the reference's own equations (fields_exprs.py:106-137,
    # placed after the injection, operators.py:131): u(-k) = -u(k-1) for k = 1 .. so/2 evaluated on
    # the freshly updated level, THEN u(0) = 0 -- the rows above the surface are stored state.
    ghost = np.zeros((n, n, r), dtype)       # ghost[..., k-1] = u(z = -k), mirror="judi" only

    def lap(f):
        p = np.pad(f, r)
        if fs and mirror == "devito":
            p[:, :, :r] = -p[:, :, 2 * r:r:-1]
        elif fs:
            p[r:-r, r:-r, :r] = ghost[:, :, ::-1]
        c = p[r:-r, r:-r, r:-r]
        L = 3 * w[r] * c
        for k in range(1, r + 1):
            L = L + w[r + k] * (p[r + k:n + r + k, r:-r, r:-r] + p[r - k:n + r - k, r:-r, r:-r] +
                                p[r:-r, r + k:n + r + k, r:-r] + p[r:-r, r - k:n + r - k, r:-r] +
                                p[r:-r, r:-r, r + k:nz + r + k] + p[r:-r, r:-r, r - k:nz + r - k])
        return L / (h * h)
    for it in range(1, tM + 1):
        L = lap(u)
        if udt == "forward":
            un = (a_ * (2 * u - up) + b_ * u + L) / (a_ + b_)
        else:
            un = (a_ * (2 * u - up) + 0.5 * b_ * up + L) / (a_ + 0.5 * b_)
        if fs and mirror == "devito":
            un[:, :, 0] = 0
        for (cx, cy, cz, ww) in src_cells:
            un[cx, cy, cz] += ww * q[it] * dt ** 2 / m[cx, cy, cz]
        if fs and mirror == "judi":
            ghost = -un[:, :, :r]
            un[:, :, 0] = 0
        rec[it] = (u if rec_level == "current" else un)[nbl:nbl + shape[0], nbl:nbl + shape[1], rz]
        up, u = u, un
    nrm = float(np.sqrt(np.sum(rec.astype(np.float64) ** 2)))
    if not full:
        return nrm, dt, nt
    xs = h * np.arange(shape[0])
    rec_xyz = np.array([[x, y, 2 * h] for x in xs for y in h * np.arange(shape[1])], np.float32)
    src_xyz = np.array([[0.5 * dom, 0.5 * dom, h]], np.float32)
    return nrm, dt, nt, rec.reshape(nt, -1), q, v, src_xyz, rec_xyz


DEVITO_NORM_REC_FS = 369.955


if __name__ == "__main__":
    import sys
    variants = [dict(), dict(dtype=np.float32), dict(tmax_rule="nt-1"), dict(rec_level="new"),
                dict(udt="centred"), dict(dt_rule="0.38"), dict(fs=True), dict(fs=True, dtype=np.float32)]
    if len(sys.argv) > 1:
        variants = [variants[int(k)] for k in sys.argv[1:]]
    for kw in variants:
        nr, dt, nt = run(**kw)[:3]
        ref = DEVITO_NORM_REC_FS if kw.get("fs") else DEVITO_NORM_REC
        print("%-40s dt = %.3f nt = %d: norm(rec) = %.4f  (%+.3f %% vs Devito's %.4f)"
              % (", ".join("%s=%s" % (k, getattr(v, "__name__", v)) for k, v in kw.items()) or "oracle conventions",
                 dt, nt, nr, 100 * (nr / ref - 1), ref), flush=True)
