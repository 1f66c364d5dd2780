import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import oracle as O

def setup(tti=False, rho=False, prec="f32", so=8, n=(101, 81), nbl=20, dm=False):
    d = (10., 10.)
    v = np.ones(n, dtype=np.float32) * 1.5
    v[:, n[1]//2:] = 2.5
    m = (1/v**2).astype(np.float32)
    kw = {}
    if rho: kw['rho'] = ((v + .5)/2).astype(np.float32)
    if tti:
        kw.update(epsilon=((v-1.5)/12).astype(np.float32), delta=((v-1.5)/14).astype(np.float32), theta=((v-1.5)/4).astype(np.float32))
    if dm:
        dmv = np.zeros(n, dtype=np.float32); dmv[30:70, 30:60] = 0.05
        kw['dm'] = dmv
    return O.Model((0., 0.), d, n, space_order=so, nbl=nbl, m=m, precision=prec, **kw)

for prec in ("f32", "f64"):
  for tti, rho in ((False, False), (False, True), (True, False)):
    model = setup(tti, rho, prec, dm=True)
    dt = model.critical_dt
    nt = 300
    q = O.ricker_wavelet((nt-1)*dt, dt, 0.015)[:nt]
    src = np.array([[500., 20.]], dtype=np.float32)
    rec = np.stack([np.linspace(50, 950, 51), np.full(51, 30.)], axis=1).astype(np.float32)
    t0 = time.time()
    d, _ = O.forward_rec(model, src, q, rec)
    y = np.random.default_rng(0).standard_normal(d.shape).astype(d.dtype)
    qh, _ = O.forward_rec(model, rec, y, src, fw=False)
    a = np.dot(d.ravel().astype(np.float64), y.ravel()); b = np.dot(q[:,0].astype(np.float64), qh[:,0])
    dl, _ = O.born_rec(model, src, q, rec)
    g, _, _ = O.J_adjoint(model, src, q, rec, y, is_residual=True)
    c = float(dt)*np.dot(dl.ravel().astype(np.float64), y.ravel()); e = np.dot(g.ravel().astype(np.float64), model.dm.ravel())
    print(prec, "tti" if tti else ("rho" if rho else "iso"), "dt=%.4f" % dt, "F: %.3e" % ((a-b)/(a+b)), "J: %.3e" % ((c-e)/(c+e)), "|d|=%.3e |dl|=%.3e" % (np.linalg.norm(d), np.linalg.norm(dl)), "%.1fs" % (time.time()-t0))
