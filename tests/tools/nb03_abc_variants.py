"""Notebook 03 (`/root/reference/examples/notebooks/03_constrained_fwi.ipynb`) in plain numpy, with
the absorbing-boundary formulation as a parameter: which historical form of the reference's
boundary term reproduces the numbers the notebook printed (scale = 0.018069575143305386, cell 35)?

Independent of the oracle's C code on purpose (a second restatement of the same experiment: so = 8
centred Laplacian, bilinear injection / interpolation, sinc resampling both ways); the variant
"damp417" must reproduce the oracle's own value 0.0160153, which checks this script.

    python tests/tools/nb03_abc_variants.py            # table of variants
"""
import sys
import os
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle as O   # only ricker_wavelet / time_resample / cfl_coeff (host helpers)

NB_SCALE = 0.018069575143305386
N, H, NBL = 251, 15.0, 40
TN, DT, F0 = 4000.0, np.float32(2.0), 0.005
NSRC, NREC = 8, 251
W8 = np.array([-205 / 72, 8 / 5, -1 / 5, 8 / 315, -1 / 560])   # centred 2nd derivative, so = 8


def laplacian(u):
    """u: (..., nx, nz) with a zero halo of 4 implied (Dirichlet)."""
    p = np.pad(u, [(0, 0)] * (u.ndim - 2) + [(4, 4), (4, 4)])
    nx, nz = u.shape[-2:]
    c = p[..., 4:4 + nx, 4:4 + nz]
    L = 2 * W8[0] * c
    for k in range(1, 5):
        L = L + W8[k] * (p[..., 4 + k:4 + k + nx, 4:4 + nz] + p[..., 4 - k:4 - k + nx, 4:4 + nz] +
                         p[..., 4:4 + nx, 4 + k:4 + k + nz] + p[..., 4:4 + nx, 4 - k:4 - k + nz])
    return L / (H * H)


def profile(width, coeff_den, over_h=True):
    """One side of the layer, outermost cell first: coeff (pos - sin(2 pi pos) / (2 pi)) [/ h]."""
    i = np.arange(width, dtype=np.float64)
    pos = np.abs((width - i + 1) / float(width))
    val = 1.5 * np.log(1000.0) / coeff_den * (pos - np.sin(2 * np.pi * pos) / (2 * np.pi))
    return val / H if over_h else val


def damp_field(width, coeff_den, over_h=True):
    n = N + 2 * NBL
    d = np.zeros((n, n))
    p = profile(width, coeff_den, over_h)
    for ax in range(2):
        shp = [1, 1]
        shp[ax] = width
        sl = [slice(None)] * 2
        sl[ax] = slice(0, width)
        d[tuple(sl)] += p.reshape(shp)
        sl[ax] = slice(n - width, n)
        d[tuple(sl)] += p[::-1].reshape(shp)
    return d


def bilinear(x, z):
    """grid index / weights of a point at physical (x, z)."""
    fx, fz = x / H + NBL, z / H + NBL
    ix, iz = int(np.floor(fx)), int(np.floor(fz))
    ax, az = fx - ix, fz - iz
    return [(ix, iz, (1 - ax) * (1 - az)), (ix + 1, iz, ax * (1 - az)),
            (ix, iz + 1, (1 - ax) * az), (ix + 1, iz + 1, ax * az)]


def run(m_phys, abc="damp417", dt=None, width=NBL - 3, coeff_den=None, over_h=True):
    """All 8 shots at once -> list of 8 shot records on the 2 ms axis."""
    coeff_den = coeff_den if coeff_den is not None else width
    m = np.pad(m_phys.astype(np.float64), NBL, mode="edge")
    vmax = np.sqrt(1.0 / m.min())
    if dt is None:
        dt = np.float32("%.3e" % (O.cfl_coeff(8, 2) * H / vmax))
    dt = float(np.float32(dt))
    q = O.ricker_wavelet(TN, DT, F0)
    nt = q.shape[0]
    ntc = int(np.floor(np.float32(TN) / np.float32(dt))) + 1
    qc = O.time_resample(q, (0, DT, nt), (0, np.float32(dt), ntc)).astype(np.float64)[:, 0]
    damp = damp_field(width, coeff_den, over_h)
    zsrc = np.linspace(0.0, (N - 1) * H, NSRC)
    srcs = [bilinear(H, z) for z in zsrc]
    ixr = int(round((N - 2) * H / H)) + NBL
    u = np.zeros((NSRC,) + m.shape)
    up = np.zeros_like(u)
    rec = np.zeros((ntc, NSRC, NREC))
    a, b = m / dt ** 2, damp / dt
    mask = 1.0 - damp
    for t in range(1, ntc - 1):
        L = laplacian(u)
        if abc == "damp417":      # solve(m u.dt2 + damp u.dt - L, u.forward), u.dt one-sided
            un = (a * (2 * u - up) + b * u + L) / (a + b)
        elif abc == "centred":    # u.dt = (u+ - u-) / (2 dt)
            un = (a * (2 * u - up) + 0.5 * b * up + L) / (a + 0.5 * b)
        elif abc == "mask":       # u+ = damp (2 u - damp u- + dt^2 / m L), damp = 1 - profile
            un = mask * (2 * u - mask * up + dt ** 2 / m * L)
        else:
            raise ValueError(abc)
        for s, corners in enumerate(srcs):
            for ix, iz, w in corners:
                un[s, ix, iz] += w * qc[t] * dt ** 2 / m[ix, iz]
        rec[t] = u[:, ixr, NBL:NBL + N]       # receivers sit on nodes (x = 3735, z = 15 k)
        up, u = u, un
    out = []
    for s in range(NSRC):
        out.append(O.time_resample(rec[:, s, :].astype(np.float32), (0, np.float32(dt), ntc), (0, DT, nt)))
    return out


def scale_of(records):
    nd = sum(float(DT) * float(np.sum(r.astype(np.float64) ** 2)) for r in records)
    return 10.0 / np.sqrt(nd / NSRC)


def main():
    m = np.float32(1.5) ** -2 * np.ones((N, N), np.float32)
    m[100:150, 100:150] = np.float32(1.7) ** -2
    variants = [
        ("v4.1.7: damp u.dt, layer nbl-3, cfl dt", dict(abc="damp417")),
        ("damp u.dt, layer nbl, dt = 0.42 h / vmax", dict(abc="damp417", width=NBL, dt=np.float32("%.3e" % (0.42 * H / 1.7)))),
        ("mask, layer nbl, coeff / h, dt 0.42", dict(abc="mask", width=NBL, dt=np.float32("%.3e" % (0.42 * H / 1.7)))),
        ("mask, layer nbl, coeff / h, cfl dt", dict(abc="mask", width=NBL)),
        ("mask, layer nbl, coeff not / h, dt 0.42", dict(abc="mask", width=NBL, over_h=False, dt=np.float32("%.3e" % (0.42 * H / 1.7)))),
    ]
    for name, kw in variants:
        s = scale_of(run(m, **kw))
        print("%-48s scale = %.7f  (%+.2f %% vs notebook)" % (name, s, 100 * (s / NB_SCALE - 1)), flush=True)


if __name__ == "__main__":
    main()
