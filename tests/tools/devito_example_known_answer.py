"""A known answer produced by Devito itself, restated in plain numpy with the conventions the
oracle ASSUMES for Devito's lowering (the `[dv]` items).

Devito (the third-party dependency the reference's arithmetic lives in; absent from this image)
pins its own acoustic example in its test-suite: `examples/seismic/acoustic/acoustic_example.py`
`run()` with the defaults shape = (50, 50, 50), spacing = 20 m, tn = 1000 ms, space_order = 4,
nbl = 40, preset 'layers-isotropic' (3 layers, 1.5 ... 3.5 km/s), kernel OT2, and
`tests/test_..::test_isoacoustic` asserts `norm(rec) == 459.1678` (rtol 1e-3) -- recalled from
Devito's public repository, NOT present under /root/reference, so this script is evidence about
the `[dv]` assumptions, not a parity test of the reference's own files.

What the experiment exercises is exactly the list of Devito behaviours the oracle could not read
in the reference's Python: the centred `.laplace` weights, the one-sided `u.dt` in
`solve(m u.dt2 - u.laplace + damp u.dt, u.forward)`, the time-loop bounds 1 .. nt-2, injection of
`src dt^2 / m` into the new level with (bi/tri)linear weights, interpolation of the current level,
and the damping profile Devito's own example Model builds (layer of nbl cells).

    python tests/tools/devito_example_known_answer.py
"""
import itertools
import numpy as np

DEVITO_NORM_REC = 459.1678


def cfl_sympy(space_order, ndim):
    from sympy import finite_diff_weights as fd_w
    c = fd_w(2, range(-space_order, space_order + 1), 0)[-1][-1]
    return float(np.sqrt(4.0 / (ndim * sum(abs(float(x)) for x in c))))


def run(dt_rule="0.38", udt="forward", tmax_rule="nt-2", shape=(50, 50, 50), h=20.0, tn=1000.0, so=4,
        nbl=40, f0=0.010, dtype=np.float64):
    from sympy import finite_diff_weights as fd_w
    r = so // 2
    w = [float(x) for x in fd_w(2, list(range(-r, r + 1)), 0)[2][-1]]   # centred 2nd derivative
    nl = 3
    v = np.full(shape, 1.5)
    vi = np.linspace(1.5, 3.5, nl)
    for i in range(1, nl):
        v[..., i * int(shape[-1] / nl):] = vi[i]
    vp = np.pad(v, nbl, mode="edge")
    m = (1.0 / vp ** 2).astype(dtype)
    coeff = {"0.38": 0.38, "sympy": cfl_sympy(so, 3)}[dt_rule]
    dt = float(np.float32("%.3e" % (coeff * h / vp.max())))
    nt = int(np.ceil((tn + dt) / dt))
    t = np.linspace(0.0, dt * (nt - 1), nt)
    rr = np.pi * f0 * (t - 1.0 / f0)
    q = (1 - 2.0 * rr ** 2) * np.exp(-rr ** 2)
    # damping layer of Devito's example Model: nbl cells, coefficient 1.5 ln(1000) / nbl, / spacing
    n = shape[0] + 2 * nbl
    i = np.arange(nbl, dtype=np.float64)
    pos = np.abs((nbl - i + 1) / float(nbl))
    prof = 1.5 * np.log(1000.0) / nbl * (pos - np.sin(2 * np.pi * pos) / (2 * np.pi)) / h
    damp = np.zeros((n, n, n))
    for ax in range(3):
        shp = [1, 1, 1]
        shp[ax] = nbl
        sl = [slice(None)] * 3
        sl[ax] = slice(0, nbl)
        damp[tuple(sl)] += prof.reshape(shp)
        sl[ax] = slice(n - nbl, n)
        damp[tuple(sl)] += prof[::-1].reshape(shp)
    damp = damp.astype(dtype)
    # geometry: source at the centre in x, y and one cell deep; receivers on every node, two cells deep
    dom = h * (shape[0] - 1)
    sx = (0.5 * dom) / h + nbl
    ix = int(np.floor(sx))
    ax_ = sx - ix
    sz = 1 + nbl
    src_cells = [(ix + a, ix + b, sz, (ax_ if a else 1 - ax_) * (ax_ if b else 1 - ax_))
                 for a, b in itertools.product((0, 1), (0, 1))]
    rz = 2 + nbl
    u = np.zeros((n, n, n), dtype)
    up = np.zeros_like(u)
    rec = np.zeros((nt, shape[0], shape[1]), dtype)
    a_, b_ = m / dt ** 2, damp / dt
    tM = {"nt-2": nt - 2, "nt-1": nt - 1}[tmax_rule]

    def lap(f):
        p = np.pad(f, r)
        c = p[r:-r, r:-r, r:-r]
        L = 3 * w[r] * c
        for k in range(1, r + 1):
            L = L + w[r + k] * (p[r + k:n + r + k, r:-r, r:-r] + p[r - k:n + r - k, r:-r, r:-r] +
                                p[r:-r, r + k:n + r + k, r:-r] + p[r:-r, r - k:n + r - k, r:-r] +
                                p[r:-r, r:-r, r + k:n + r + k] + p[r:-r, r:-r, r - k:n + r - k])
        return L / (h * h)
    for it in range(1, tM + 1):
        L = lap(u)
        if udt == "forward":
            un = (a_ * (2 * u - up) + b_ * u + L) / (a_ + b_)
        else:
            un = (a_ * (2 * u - up) + 0.5 * b_ * up + L) / (a_ + 0.5 * b_)
        for (cx, cy, cz, ww) in src_cells:
            un[cx, cy, cz] += ww * q[it] * dt ** 2 / m[cx, cy, cz]
        rec[it] = u[nbl:nbl + shape[0], nbl:nbl + shape[1], rz]
        up, u = u, un
    return float(np.sqrt(np.sum(rec.astype(np.float64) ** 2))), dt, nt


if __name__ == "__main__":
    for dt_rule, udt, tmax in itertools.product(("0.38", "sympy"), ("forward", "centred"), ("nt-2",)):
        nr, dt, nt = run(dt_rule, udt, tmax)
        print("dt rule %-5s (dt = %.3f, nt = %d), u.dt %-7s, time_M %s: norm(rec) = %.4f  (%+.3f %% vs Devito's %.4f)"
              % (dt_rule, dt, nt, udt, tmax, nr, 100 * (nr / DEVITO_NORM_REC - 1), DEVITO_NORM_REC), flush=True)
