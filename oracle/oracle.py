"""
CPU ORACLE (test infrastructure, NOT product code) -- numpy/ctypes front end of judi_oracle.c.

STENCIL PARITY UNPINNED, host logic and point-wise algebra pinned to the executed reference: see
the header of judi_oracle.c.  Nothing under judi.jl_b200/ may import this.

The classes/functions below restate, with the same names and argument meaning, the part of
/root/reference/src/pysource that sits above the Devito operators:

  Model                models.py:123-527  (padded grid :157-180, damp :54-120, params :301-326,
                                           critical_dt :466-482, _cfl_coeff :440-456)
  forward / born / gradient   propagators.py:15-230 (what they allocate, bind and return)
  forward_rec / born_rec / J_adjoint / J_adjoint_standard / J_adjoint_checkpointing
                              interface.py:17-47, 208-241, 279-345, 411-470, 473-562
  Loss                 sensitivity.py:236-276
  J_adjoint_freq       interface.py:348-408, fields_exprs.py:140-169, sensitivity.py:72-104
  remove_padding / pad_array  src/TimeModeling/Utils/auxiliaryFunctions.jl:52-89
  ricker_wavelet       src/TimeModeling/Utils/auxiliaryFunctions.jl:205-213
  time_resample        src/TimeModeling/Utils/time_utilities.jl:32-48, 84
  limit_model_to_receiver_area  src/TimeModeling/Utils/auxiliaryFunctions.jl:129-173

Array conventions are pysource's: model arrays have shape (nx[,ny],nz); coordinates are
(npoint, ndim) in metres; wavelets / shot records are (nt, ntrace).
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compile the two oracle libraries with the committed Makefile."""
    need = force or not all(os.path.exists(os.path.join(_HERE, "libjudi_oracle_%s.so" % p))
                            for p in ("f32", "f64"))
    src_m = os.path.getmtime(os.path.join(_HERE, "judi_oracle.c"))
    for p in ("f32", "f64"):
        f = os.path.join(_HERE, "libjudi_oracle_%s.so" % p)
        if os.path.exists(f) and os.path.getmtime(f) < src_m:
            need = True
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)


def _lib(precision):
    if precision in _LIBS:
        return _LIBS[precision]
    path = os.path.join(_HERE, "libjudi_oracle_%s.so" % precision)
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    real = ctypes.c_float if precision == "f32" else ctypes.c_double
    P = ctypes.c_void_p
    lib.oracle_create.restype = P
    lib.oracle_create.argtypes = [ctypes.c_int, P, ctypes.c_int, P, P, real, P, P, P, real,
                                  ctypes.c_int, P, real, P, real, P, real, P, real, P]
    lib.oracle_destroy.argtypes = [P]
    lib.oracle_sparse_create.restype = P
    lib.oracle_sparse_create.argtypes = [P, ctypes.c_int, P]
    lib.oracle_sparse_destroy.argtypes = [P]
    lib.oracle_sparse_export.argtypes = [P, P, P, P]
    lib.oracle_weights.argtypes = [P, P, P]
    lib.oracle_forward.argtypes = [P, ctypes.c_int, ctypes.c_int, P, P, P, P, ctypes.c_int,
                                   ctypes.c_int, P, P]
    lib.oracle_born.argtypes = [P, ctypes.c_int, P, P, P, P, P, ctypes.c_int, ctypes.c_int,
                                ctypes.c_int, P, P]
    lib.oracle_gradient.argtypes = [P, ctypes.c_int, P, P, ctypes.c_int, ctypes.c_int, P, P, P]
    lib.oracle_time_steps.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
    lib.oracle_set_fast.argtypes = [ctypes.c_int]
    lib.oracle_set_fs.argtypes = [P, ctypes.c_int]
    lib.oracle_set_visco.argtypes = [P, P, real]
    lib.oracle_set_threads.argtypes = [ctypes.c_int]
    lib.oracle_set_threads.restype = ctypes.c_int
    lib.oracle_set_ext.argtypes = [P, P, P, P, P, P]
    lib.oracle_tti_flux_point.argtypes = [ctypes.c_int, real, real, real, real, real, P, P, P, P]
    lib.oracle_ic_point.argtypes = [ctypes.c_int, real, real, real, real]
    lib.oracle_ic_point.restype = real
    lib.oracle_linsrc_point.argtypes = [ctypes.c_int, real, real, real, real, real]
    lib.oracle_linsrc_point.restype = real
    lib.oracle_leapfrog_point.argtypes = [real] * 7
    lib.oracle_leapfrog_point.restype = real
    lib.oracle_visco_point.argtypes = [ctypes.c_int] + [real] * 11 + [P, P, P, P]
    lib._real = real
    lib._dtype = np.float32 if precision == "f32" else np.float64
    _LIBS[precision] = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


IC = {"as": 0, "isic": 1, "fwi": 2}


# --------------------------------------------------------------------------------------------
# Host helpers (Julia side of the boundary)
def ricker_wavelet(tmax, dt, f0, t0=None):
    """auxiliaryFunctions.jl:205-213 (f0 in kHz, times in ms) -> (nt, 1) float32."""
    tmax, dt, f0 = np.float32(tmax), np.float32(dt), np.float32(f0)
    if t0 is None:
        t0 = np.float32(0)
    else:
        tmax = np.float32(tmax - t0)
    nt = int(np.floor(tmax / dt)) + 1
    t = np.linspace(t0, tmax, nt, dtype=np.float32)
    r = (np.float32(np.pi) * f0 * (t - np.float32(1) / f0)).astype(np.float32)
    q = ((1 - 2 * r ** 2) * np.exp(-r ** 2)).astype(np.float32)
    return q.reshape(nt, 1)


def pad_array(m, nb, mode="border"):
    """auxiliaryFunctions.jl:52-62."""
    return np.pad(m, nb, mode="edge" if mode == "border" else "constant")


def remove_padding(g, nb, true_adjoint=False):
    """auxiliaryFunctions.jl:79-89: crop; true_adjoint first folds the pad into the edge."""
    g = np.array(g, copy=True)
    if true_adjoint:
        for dim, (nbl, nbr) in enumerate(nb):
            N = g.shape[dim]
            sl = [slice(None)] * g.ndim

            def sel(i):
                s = list(sl)
                s[dim] = i
                return tuple(s)
            g[sel(nbl)] += g[sel(slice(0, nbl))].sum(axis=dim)
            g[sel(N - nbr - 1)] += g[sel(slice(N - nbr, N))].sum(axis=dim)
    crop = tuple(slice(nbl, n - nbr) for (nbl, nbr), n in zip(nb, g.shape))
    return g[crop]


def _steprange(t0, dt, n):
    """Elements of a Julia Float32 StepRangeLen (computed in double, rounded once)."""
    return (np.float64(np.float32(t0)) + np.arange(n, dtype=np.float64) *
            np.float64(np.float32(dt))).astype(np.float32)


def time_axis(dt, t):
    """`0:dt:(dt*ceil(t/dt))` (time_utilities.jl:50-51) as (t0, dt, nt)."""
    dt = np.float32(dt)
    return (np.float32(0), dt, int(np.ceil(np.float32(t) / dt)) + 1)


def time_resample(data, t_in, t_new):
    """time_utilities.jl:32-48: `data` (nt_in, ntrace) sampled on the range t_in = (t0, dt, nt)
    -> samples on t_new.  Equal steps: shifted view; integer ratio: decimation; else
    `sinc.((Up .- S') ./ dt) * Y` in Float32 (:84)."""
    data = np.asarray(data, dtype=np.float32)
    if data.ndim == 1:
        data = data.reshape(-1, 1)
    t0i, dti, nti = np.float32(t_in[0]), np.float32(t_in[1]), int(t_in[2])
    t0n, dtn, ntn = np.float32(t_new[0]), np.float32(t_new[1]), int(t_new[2])
    assert data.shape[0] == nti
    if dtn == dti:
        idx0 = int(np.trunc(np.float32(t0n - t0i) / dtn))
        return data[idx0:, :]
    if np.fmod(dtn, dti) == 0:
        rate = int(np.trunc(dtn / dti))
        idx0 = int(np.trunc(np.float32(t0n - t0i) / dtn))
        return data[::rate, :][idx0:, :]
    S = _steprange(t0i, dti, nti)
    Up = _steprange(t0n, dtn, ntn)
    x = ((Up[:, None] - S[None, :]) / dti).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        # Julia: sinc(x) = sinpi(x) / (pi * x); sinpi is exact in the argument reduction
        red = (x.astype(np.float64) - 2.0 * np.round(x.astype(np.float64) / 2.0))
        sp = np.sin(np.pi * red).astype(np.float32)
        k = np.where(x == 0, np.float32(1), sp / (np.float32(np.pi) * x)).astype(np.float32)
    return (k @ data).astype(np.float32)


def limit_model_to_receiver_area(src_coords, rec_coords, origin, spacing, shape, buffer, params,
                                 pert=None):
    """auxiliaryFunctions.jl:129-173: restrict the model arrays in `params` (dict of arrays or
    scalars, shape `shape`) to the x[,y] range covered by sources and receivers plus `buffer`
    metres.  Returns (new_origin, new_shape, new_params, new_pert, index slices)."""
    ndim = len(shape)
    src = np.asarray(src_coords, dtype=np.float32).reshape(-1, ndim)
    rec = np.asarray(rec_coords, dtype=np.float32).reshape(-1, ndim)
    inds = []
    for d in range(ndim - 1):
        o, h, n = np.float32(origin[d]), np.float32(spacing[d]), shape[d]
        lo = min(rec[:, d].min(), src[:, d].min())
        hi = max(rec[:, d].max(), src[:, d].max())
        lo = max(o, np.float32(lo - np.float32(buffer)))
        hi = min(np.float32(o + h * np.float32(n - 1)), np.float32(hi + np.float32(buffer)))
        i0 = int(np.floor(np.float32(lo - o) / h)) + 1       # Julia 1-based, `÷` on floats
        i1 = int(np.floor(np.float32(hi - o) / h)) + 1
        inds.append(slice(max(1, i0) - 1, i1))
    inds.append(slice(0, shape[-1]))
    inds = tuple(inds)
    newp = {k: (np.asarray(v)[inds] if isinstance(v, np.ndarray) and v.ndim == ndim else v)
            for k, v in params.items()}
    new_shape = tuple(s.stop - s.start for s in inds)
    new_origin = tuple(np.float32(origin[d]) + np.float32(spacing[d]) * np.float32(inds[d].start)
                       for d in range(ndim))
    newpert = None if pert is None else np.asarray(pert).reshape(shape)[inds]
    return new_origin, new_shape, newp, newpert, inds


# --------------------------------------------------------------------------------------------
def _damp_profile(npad, h):
    """One side of models.py:74-94: values for the outer (npad-3) cells, outermost first."""
    nb = npad - 3
    if nb <= 0:
        return np.zeros(0, dtype=np.float32)
    coeff = np.float32(1.5 * np.log(1.0 / 0.001) / nb)
    i = np.arange(nb, dtype=np.float32)
    pos = np.abs((np.float32(nb) - i + np.float32(1)) / np.float32(nb)).astype(np.float32)
    val = coeff * (pos - np.sin(np.float32(2 * np.pi) * pos) / np.float32(2 * np.pi))
    return (val.astype(np.float32) / np.float32(h)).astype(np.float32)


def damp_field(shape_pml, padsizes, spacing, fs=False):
    """initialize_damp (models.py:99-120) with abc_type='damp': separable sum of 1-D profiles,
    accumulated dimension by dimension in fp32 (left then right, like the Inc sequence)."""
    damp = np.zeros(shape_pml, dtype=np.float32)
    nd = len(shape_pml)
    for d, ((nbl, nbr), h) in enumerate(zip(padsizes, spacing)):
        shp = [1] * nd
        if not fs or d != nd - 1:
            prof = _damp_profile(nbl, h)
            n = prof.size
            if n:
                shp[d] = n
                sl = [slice(None)] * nd
                sl[d] = slice(0, n)
                damp[tuple(sl)] += prof.reshape(shp)
        prof = _damp_profile(nbr, h)
        n = prof.size
        if n:
            shp = [1] * nd
            shp[d] = n
            sl = [slice(None)] * nd
            sl[d] = slice(shape_pml[d] - n, shape_pml[d])
            damp[tuple(sl)] += prof[::-1].reshape(shp)
    return damp


def cfl_coeff(space_order, ndim):
    """models.py:440-456 (acoustic / TTI branch), literally, with sympy."""
    from sympy import finite_diff_weights as fd_w
    so = max(space_order // 2, 4)
    coeffs = fd_w(2, range(-so, so), 0)[-1][-1]
    return .9 * np.sqrt(4 / float(ndim * sum(np.abs(np.array(coeffs, dtype=np.float64)))))


class Model(object):
    """Restatement of pysource models.Model for the acoustic / TTI families."""

    def __init__(self, origin, spacing, shape, space_order=8, nbl=40, m=None, epsilon=None,
                 delta=None, theta=None, phi=None, rho=None, b=None, dm=None, fs=False,
                 dt=None, precision="f32", qp=None, f0=0.015):
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.nbl = int(nbl)
        self.fs = fs
        self.space_order = int(space_order)
        self.spacing = tuple(np.float32(s) for s in spacing)
        self.origin = tuple(np.float32(o) for o in origin)
        # models.py:172-177
        op = [np.float32(float(o) - float(s) * nbl) for o, s in zip(origin, spacing)]
        sp = [int(s + 2 * self.nbl) for s in self.shape]
        if fs:   # models.py:174-176
            op[-1] = np.float32(origin[-1])
            sp[-1] -= self.nbl
        self.origin_pml = tuple(op)
        self.shape_pml = tuple(sp)
        self.precision = precision
        self._dt = dt
        self.damp = damp_field(self.shape_pml, self.padsizes, self.spacing, fs)
        self.m = self._param(m)
        # density (models.py:250-266): scalar -> Constant irho; array -> field irho = 1/rho
        self.irho_field = None
        self.irho_const = 1.0
        if rho is not None:
            if np.ndim(rho) == 0:
                self.irho_const = float(np.reciprocal(np.float32(rho)))
            else:
                self.irho_field = self._param(np.reciprocal(np.asarray(rho, dtype=np.float32)))
        elif b is not None:
            if np.ndim(b) == 0:
                self.irho_const = float(np.float32(b))
            else:
                self.irho_field = self._param(b)
        self.dm = None if dm is None or (np.ndim(dm) == 0 and dm == 0) else self._param(dm)
        # visco-acoustic (models.py:208-214): quality factor field; f0 is the peak frequency the
        # reference passes with every call (propagators.py:46), in kHz
        self.is_viscoacoustic = qp is not None
        self.qp = None if qp is None else self._param(np.broadcast_to(np.asarray(qp, dtype=np.float32), self.shape))
        self.f0 = float(f0)
        self.is_tti = any(p is not None for p in (epsilon, delta, theta, phi))
        self.tti = {}
        self.scale = 1.0
        if self.is_tti:
            # models.py:216-225
            eps = 1 if epsilon is None else 1 + 2 * np.asarray(epsilon, dtype=np.float32)
            dlt = 1 if delta is None else 1 + 2 * np.asarray(delta, dtype=np.float32)
            self._eps_max = float(np.max(eps))
            for name, val, default in (("epsilon", eps, 1.0), ("delta", dlt, 1.0),
                                       ("theta", theta, 0.0), ("phi", phi, 0.0)):
                if name == "phi" and self.dim == 2:
                    val = None
                if val is None:
                    self.tti[name] = (None, default)
                elif np.ndim(val) == 0:
                    self.tti[name] = (None, float(np.float32(val)))
                else:
                    self.tti[name] = (self._param(val), 0.0)
        self._handle = None

    def _param(self, field):
        if field is None:
            return None
        field = np.asarray(field, dtype=np.float32)
        if field.shape == self.shape_pml:
            return np.ascontiguousarray(field)
        assert field.shape == self.shape, (field.shape, self.shape)
        return np.ascontiguousarray(np.pad(field, self.padsizes, mode="edge"))

    @property
    def padsizes(self):
        p = [(self.nbl, self.nbl) for _ in range(self.dim - 1)]
        p.append((0 if self.fs else self.nbl, self.nbl))
        return tuple(p)

    @property
    def critical_dt(self):
        """models.py:466-482."""
        vmax = np.sqrt(1. / np.min(self.m))
        scale = np.sqrt(1 + 2 * self._eps_max) if self.is_tti else 1
        dt = cfl_coeff(self.space_order, self.dim) * np.min(self.spacing) / (scale * vmax)
        dt = np.float32("%.3e" % dt)
        if self._dt:
            if self._dt <= dt:
                return np.float32("%.3e" % self._dt)
        return dt

    # ---- C handle -------------------------------------------------------------------------
    def handle(self):
        if self._handle is not None:
            return self._handle
        lib = _lib(self.precision)
        dt_ = lib._dtype
        self._keep = []

        def conv(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt_)
            self._keep.append(a)
            return a
        N = np.array(self.shape_pml, dtype=np.int32)
        h = np.array(self.spacing, dtype=dt_)
        o = np.array(self.origin_pml, dtype=np.float32)
        t = self.tti
        g = lambda k, d: t.get(k, (None, d))
        self._handle = lib.oracle_create(
            self.dim, _ptr(N), self.space_order, _ptr(h), _ptr(o), lib._real(self.critical_dt),
            _ptr(conv(self.m)), _ptr(conv(self.damp)), _ptr(conv(self.irho_field)),
            lib._real(self.irho_const), int(self.is_tti),
            _ptr(conv(g("epsilon", 1.0)[0])), lib._real(g("epsilon", 1.0)[1]),
            _ptr(conv(g("delta", 1.0)[0])), lib._real(g("delta", 1.0)[1]),
            _ptr(conv(g("theta", 0.0)[0])), lib._real(g("theta", 0.0)[1]),
            _ptr(conv(g("phi", 0.0)[0])), lib._real(g("phi", 0.0)[1]),
            _ptr(conv(self.dm)))
        lib.oracle_set_fs(self._handle, int(bool(self.fs)))
        if self.is_viscoacoustic:
            lib.oracle_set_visco(self._handle, _ptr(conv(self.qp)), lib._real(self.f0))
        return self._handle

    def set_f0(self, f0):
        """Peak frequency (kHz) of the next calls (the reference's f0 keyword)."""
        self.f0 = float(f0)
        if self._handle is not None and self.is_viscoacoustic:
            lib = _lib(self.precision)
            lib.oracle_set_visco(self._handle, _ptr(np.ascontiguousarray(self.qp, dtype=lib._dtype)),
                                 lib._real(self.f0))

    def __del__(self):
        try:
            if self._handle is not None:
                _lib(self.precision).oracle_destroy(self._handle)
        except Exception:
            pass

    @property
    def lib(self):
        return _lib(self.precision)

    @property
    def nfields(self):
        return 2 if self.is_tti else 1


class _Sparse(object):
    def __init__(self, model, coords):
        self.model = model
        c = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, model.dim)
        self.np = c.shape[0]
        self.h = model.lib.oracle_sparse_create(model.handle(), self.np, _ptr(c))

    def export(self):
        nc = 2 ** self.model.dim
        ijk = np.zeros((self.np, nc, 3), dtype=np.int32)
        w = np.zeros((self.np, nc), dtype=np.float32)
        valid = np.zeros((self.np, nc), dtype=np.int32)
        self.model.lib.oracle_sparse_export(self.h, _ptr(ijk), _ptr(w), _ptr(valid))
        return ijk, w, valid

    def __del__(self):
        try:
            self.model.lib.oracle_sparse_destroy(self.h)
        except Exception:
            pass


def sparse_locate(model, coords):
    """(ijk, weights, valid) of every interpolation corner: the bit-exact indexing contract."""
    return _Sparse(model, coords).export()


def nsave(nt, t_sub):
    """fields.py:141-144."""
    return nt if t_sub == 1 else int(np.ceil((nt + t_sub) / t_sub))


# --------------------------------------------------------------------------------------------
# propagators.py
class _Ext(object):
    """Binds extended source (ws, wavelet), extended receiver (wr wavelet) or a full wavefield
    source for the duration of one propagation (fields_exprs.py:58-103, fields.py:103-120)."""

    def __init__(self, model, ws=None, wavelet=None, wr=None, qwf=None):
        self.model, dt_ = model, model.lib._dtype
        self.ws = None if ws is None else np.ascontiguousarray(ws, dtype=dt_)
        self.wt = None if ws is None else np.ascontiguousarray(np.asarray(wavelet)[:, 0], dtype=dt_)
        self.wrt = None if wr is None else np.ascontiguousarray(np.asarray(wr)[:, 0], dtype=dt_)
        self.wr = None if wr is None else np.zeros(model.shape_pml, dtype=dt_)
        self.qwf = None if qwf is None else np.ascontiguousarray(qwf, dtype=dt_)
        if self.ws is not None:
            assert self.ws.shape == model.shape_pml, "weights must be padded-grid sized"

    def __enter__(self):
        self.model.lib.oracle_set_ext(self.model.handle(), _ptr(self.ws), _ptr(self.wt),
                                      _ptr(self.wr), _ptr(self.wrt), _ptr(self.qwf))
        return self

    def __exit__(self, *a):
        self.model.lib.oracle_set_ext(self.model.handle(), None, None, None, None, None)


def forward(model, src_coords, rcv_coords, wavelet, save=False, t_sub=1, illum=False, fw=True,
            ws=None, wr=None, qwf=None):
    """propagators.py:15-89 -> (rec, usave, I), or (wr, usave, I) with an extended receiver."""
    if ws is not None or wr is not None or qwf is not None:
        nt = wavelet.shape[0] if wavelet is not None else qwf.shape[0]
        with _Ext(model, ws=ws, wavelet=wavelet, wr=wr, qwf=qwf) as ext:
            wav = wavelet if wavelet is not None else np.zeros((nt, 1), dtype=np.float32)
            rec, usave, I = forward(model, src_coords, rcv_coords, wav, save=save, t_sub=t_sub,
                                    illum=illum, fw=fw)
        return (ext.wr if wr is not None else rec), usave, I
    lib, dt_ = model.lib, model.lib._dtype
    nt = wavelet.shape[0]
    src = _Sparse(model, src_coords) if src_coords is not None else None
    rec = _Sparse(model, rcv_coords) if rcv_coords is not None else None
    wav = np.ascontiguousarray(wavelet, dtype=dt_)
    rec_out = np.zeros((nt, rec.np), dtype=dt_) if rec is not None else None
    usave = None
    if save:
        usave = np.zeros((nsave(nt, t_sub), model.nfields) + model.shape_pml, dtype=dt_)
    I = np.zeros(model.shape_pml, dtype=dt_) if illum else None
    lib.oracle_forward(model.handle(), int(fw), nt, src.h if src else None, _ptr(wav),
                       rec.h if rec else None, _ptr(rec_out), int(bool(save)), int(t_sub),
                       _ptr(usave), _ptr(I))
    return rec_out, usave, I


def born(model, src_coords, rcv_coords, wavelet, save=False, ic="as", t_sub=1, nlind=False,
         illum=False, ws=None):
    """propagators.py:155-230 -> ((recl, recnl) if nlind else recl, usave, I)."""
    if ws is not None:
        with _Ext(model, ws=ws, wavelet=wavelet):
            return born(model, src_coords, rcv_coords, wavelet, save=save, ic=ic, t_sub=t_sub,
                        nlind=nlind, illum=illum)
    lib, dt_ = model.lib, model.lib._dtype
    nt = wavelet.shape[0]
    if model.dm is None:
        # propagators.py:181-194: zero perturbation -> plain forward, linearised data stay zero
        rnl, usave, I = forward(model, src_coords, rcv_coords, wavelet, save=save, t_sub=t_sub,
                                illum=illum)
        recl = np.zeros_like(rnl)
        return ((recl, rnl) if nlind else recl), usave, I
    src = _Sparse(model, src_coords) if src_coords is not None else None
    rec = _Sparse(model, rcv_coords)
    wav = np.ascontiguousarray(wavelet, dtype=dt_)
    recl = np.zeros((nt, rec.np), dtype=dt_)
    recnl = np.zeros((nt, rec.np), dtype=dt_) if nlind else None
    usave = None
    if save:
        usave = np.zeros((nsave(nt, t_sub), model.nfields) + model.shape_pml, dtype=dt_)
    I = np.zeros(model.shape_pml, dtype=dt_) if illum else None
    rc = lib.oracle_born(model.handle(), nt, src.h if src else None, _ptr(wav), rec.h, _ptr(recl), _ptr(recnl),
                         IC[ic], int(bool(save)), int(t_sub), _ptr(usave), _ptr(I))
    assert rc == 0
    return ((recl, recnl) if nlind else recl), usave, I


def gradient(model, residual, rcv_coords, usave, t_sub=1, ic="as", illum=False):
    """propagators.py:98-152 -> (gradm on the padded grid, Iv)."""
    lib, dt_ = model.lib, model.lib._dtype
    nt = residual.shape[0]
    rec = _Sparse(model, rcv_coords)
    res = np.ascontiguousarray(residual, dtype=dt_)
    g = np.zeros(model.shape_pml, dtype=dt_)
    I = np.zeros(model.shape_pml, dtype=dt_) if illum else None
    us = np.ascontiguousarray(usave, dtype=dt_)
    lib.oracle_gradient(model.handle(), nt, rec.h, _ptr(res), IC[ic], int(t_sub), _ptr(us),
                        _ptr(g), _ptr(I))
    return g, I


# --------------------------------------------------------------------------------------------
# point-wise algebra of the C loops, exposed so that tests can hold it against the vectors the
# reference's own Python produces (tests/golden/make_reference_golden.py)
_IC = {"as": 0, "isic": 1, "fwi": 2}


def tti_flux_point(ndim, b, eps_t, delta_t, theta, phi, g1, g2, precision="f64"):
    """(P, M) of FD_utils.sa_tti at one node from the staggered gradients g1, g2 (x, y, z)."""
    lib = _lib(precision)
    a1 = np.ascontiguousarray(g1, dtype=lib._dtype)
    a2 = np.ascontiguousarray(g2, dtype=lib._dtype)
    P, M = np.zeros(3, dtype=lib._dtype), np.zeros(3, dtype=lib._dtype)
    lib.oracle_tti_flux_point(ndim, b, eps_t, delta_t, theta, phi, _ptr(a1), _ptr(a2), _ptr(P), _ptr(M))
    return P, M


def ic_point(ic, us, dt2v, m, inner_grad, precision="f64"):
    """Imaging condition of one field at one point, without the weight t_sub*dt*irho."""
    return float(_lib(precision).oracle_ic_point(_IC[ic], us, dt2v, m, inner_grad))


def linsrc_point(ic, dm, irho, m, dt2u, lapw, precision="f64"):
    """Linearised (Born) source of one field at one point."""
    return float(_lib(precision).oracle_linsrc_point(_IC[ic], dm, irho, m, dt2u, lapw))


def fd_weights(model):
    """(centred 2nd-derivative weights c2[0..r], staggered 1st-derivative weights w1[1..r]) of the
    C loops, in units of 1/h^2 and 1/h."""
    lib = model.lib
    c2 = np.zeros(9, dtype=lib._dtype); w1 = np.zeros(9, dtype=lib._dtype)
    lib.oracle_weights(model.handle(), _ptr(c2), _ptr(w1))
    r = model.space_order // 2
    return c2[:r + 1].copy(), w1[1:r + 1].copy()


def leapfrog_point(dt, m, irho, damp, uc, uo, S, precision="f64"):
    """New time level of one point of the acoustic / TTI update from S = L(u) + q."""
    return float(_lib(precision).oracle_leapfrog_point(dt, m, irho, damp, uc, uo, S))


def visco_point(adj, dt, f0, qp, b, m, dl, pc, po, r, L, q=0.0, precision="f64"):
    """One point of the visco-acoustic SLS step: (r_new, p_new, w factor of p, w factor of r-)."""
    lib = _lib(precision)
    out = [np.zeros(1, dtype=lib._dtype) for _ in range(4)]
    lib.oracle_visco_point(int(adj), dt, f0, qp, b, m, dl, pc, po, r, L, q, *[_ptr(o) for o in out])
    return tuple(float(o[0]) for o in out)


# --------------------------------------------------------------------------------------------
# sensitivity.py:236-276
def Loss(dsyn, dobs, dt, is_residual=False, misfit=None):
    if misfit is not None:
        if isinstance(dsyn, tuple):
            f, r = misfit(dsyn[0], dobs - dsyn[1])
            return dt * f, r
        f, r = misfit(dsyn, dobs)
        return dt * f, r
    if not is_residual:
        if isinstance(dsyn, tuple):
            r = dsyn[0] - (dobs - dsyn[1])
            return .5 * dt * np.linalg.norm(r) ** 2, r
        r = dsyn - dobs
    else:
        r = np.array(dobs, dtype=dsyn[0].dtype if isinstance(dsyn, tuple) else dsyn.dtype)
    return .5 * dt * np.linalg.norm(r) ** 2, r


# --------------------------------------------------------------------------------------------
# interface.py
def forward_rec(model, src_coords, wavelet, rec_coords, illum=False, fw=True):
    rec, _, I = forward(model, src_coords, rec_coords, wavelet, save=False, illum=illum, fw=fw)
    return rec, I


def born_rec(model, src_coords, wavelet, rec_coords, ic="as", illum=False):
    rec, _, I = born(model, src_coords, rec_coords, wavelet, save=False, ic=ic, illum=illum)
    return rec, I


def forward_rec_w(model, weight, wavelet, rec_coords, illum=False, fw=True):
    """interface.py:50-80."""
    rec, _, I = forward(model, None, rec_coords, wavelet, ws=weight, illum=illum, fw=fw)
    return rec, I


def adjoint_w(model, rec_coords, data, wavelet, illum=False, fw=True):
    """interface.py:174-204: data injected at the receivers, wavefield correlated with the wavelet."""
    w, _, I = forward(model, rec_coords, None, data, wr=wavelet, illum=illum, fw=fw)
    return w, I


def born_rec_w(model, weight, wavelet, rec_coords, ic="as", illum=False):
    """interface.py:244-276."""
    rec, _, I = born(model, None, rec_coords, wavelet, ws=weight, ic=ic, illum=illum)
    return rec, I


def forward_wf_src(model, u, rec_coords, illum=False, fw=True):
    """interface.py:114-142: u is (nt, padded grid)."""
    rec, _, I = forward(model, None, rec_coords, None, qwf=u, illum=illum, fw=fw)
    return rec, I


def forward_wf_src_norec(model, u, illum=False, fw=True):
    """interface.py:145-171: returns the wavefield history (nt, padded grid)."""
    _, usave, I = forward(model, None, None, None, qwf=u, save=True, illum=illum, fw=fw)
    return usave[:, 0], I


def J_adjoint(model, src_coords, wavelet, rec_coords, recin, is_residual=False,
              checkpointing=False, n_checkpoints=None, t_sub=1, return_obj=False, ic="as",
              illum=False, born_fwd=False, nlind=False, misfit=None, ws=None):
    """interface.py:279-345 + :411-470.  Checkpointing (interface.py:473-562) recomputes the
    forward wavefield instead of storing it, so its result equals the t_sub=1 standard path;
    the oracle evaluates it that way."""
    if checkpointing:
        t_sub = 1
    if born_fwd:
        rec, u, Iu = born(model, src_coords, rec_coords, wavelet, save=True, nlind=nlind,
                          illum=illum, ic=ic, t_sub=t_sub, ws=ws)
    else:
        rec, u, Iu = forward(model, src_coords, rec_coords, wavelet, save=True, illum=illum,
                             t_sub=t_sub, ws=ws)
    f, residual = Loss(rec, recin, float(model.critical_dt), is_residual=is_residual,
                       misfit=misfit)
    g, Iv = gradient(model, residual, rec_coords, u, t_sub=t_sub, ic=ic, illum=illum)
    if return_obj:
        return f, g, Iu, Iv
    return g, Iu, Iv


def inner_grad_np(model):
    """sensitivity.py:212-223 on padded-grid arrays: sum_d D+_d a * D+_d b, D+ the staggered first
    derivative at +1/2 (zero outside the padded grid: Devito's halo), combined point by point."""
    _, w1 = fd_weights(model)
    r = len(w1)

    def dplus(a, ax):
        pad = [(0, 0)] * a.ndim
        pad[ax] = (r, r)
        ap = np.pad(a, pad)
        n = a.shape[ax]
        out = np.zeros_like(a, dtype=np.float64)
        for k in range(1, r + 1):
            hi = [slice(None)] * a.ndim
            lo = [slice(None)] * a.ndim
            hi[ax] = slice(r + k, r + k + n)
            lo[ax] = slice(r - (k - 1), r - (k - 1) + n)
            out += w1[k - 1] * (ap[tuple(hi)] - ap[tuple(lo)])
        return out / float(model.spacing[ax])

    def inner(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return sum(dplus(a, ax) * dplus(b, ax) for ax in range(a.ndim))
    return inner


def dft_gradient(u, v, dt, freq_list, factor=1, ic="as", m=None, inner_grad=None):
    """On-the-fly DFT gradient from wavefield histories u, v of shape (nt, nfields, ...):
      forward  fields_exprs.py:140-169:  uf[f] = sum_{t=1, t % factor == 0}^{nt-2}
                                                 factor * exp(-i 2 pi f t dt) * u[t]
      "as"     sensitivity.py:72-104:    gradm -= w_f * Re(uf[f] exp(i 2 pi f t dt) v[t]) for EVERY
               t = nt-2 .. 1, w_f = -(2 pi f)^2 / time_M, time_M = nt - 2, summed over frequencies
               and over the field pairs (u_i, v_i) of TTI.  grad_expr hands dft_sub to the imaging
               condition as `factor=` (sensitivity.py:50), which crosscorr_freq's signature does
               not take: its own dft_sub stays None and the correlation is never subsampled.
      "isic" / "fwi"  sensitivity.py:125-154:  gradm -= Re( w_f U v m - w2 inner_grad(U, v) ),
               U = uf[f] exp(i 2 pi f t dt), w2 = +-factor / time_M; isic_freq does read `factor`:
               these run on the steps with t % factor == 0 only.
      (tests/golden/make_reference_golden.py section 10 executes exactly these functions.)
    The complex product is stored into the real `gradm`, i.e. its real part is kept.
    m: padded-grid m (isic / fwi); inner_grad(a, b): the bilinear form of sensitivity.py:212-223."""
    nt = u.shape[0]
    factor = int(factor) if factor else 1
    tf = np.array([t for t in range(1, nt - 1) if t % factor == 0], dtype=np.int64)
    ta = np.arange(1, nt - 1, dtype=np.int64) if ic == "as" else tf
    ics = {"as": 0.0, "isic": 1.0, "fwi": -1.0}[ic]
    g = np.zeros(u.shape[2:], dtype=np.float64)
    for fk in np.asarray(freq_list, dtype=np.float64):
        phf = np.exp(-1j * 2 * np.pi * fk * tf * dt)
        w = -(2 * np.pi * fk) ** 2 / (nt - 2)
        w2 = ics * factor / (nt - 2)
        for fld in range(u.shape[1]):
            uf = factor * np.tensordot(phf, u[tf, fld].astype(np.float64), axes=(0, 0))
            if ic == "as":
                pha = np.exp(1j * 2 * np.pi * fk * ta * dt)
                vf = np.tensordot(pha, v[ta, fld].astype(np.float64), axes=(0, 0))
                g -= w * (uf * vf).real
                continue
            for t in ta:
                U = uf * np.exp(1j * 2 * np.pi * fk * t * dt)
                vt = v[t, fld].astype(np.float64)
                g -= w * U.real * vt * m - w2 * inner_grad(U.real, vt)
    return g


def J_adjoint_freq(model, src_coords, wavelet, rec_coords, recin, freq_list, dft_sub=None,
                   is_residual=False, return_obj=False, born_fwd=False, nlind=False, misfit=None,
                   ic="as"):
    """interface.py:348-408 (on-the-fly DFT gradient), evaluated from the stored wavefield
    histories instead of on the fly (small cases only), see dft_gradient."""
    if born_fwd:
        rec, u, _ = born(model, src_coords, rec_coords, wavelet, save=True, nlind=nlind, ic=ic)
    else:
        rec, u, _ = forward(model, src_coords, rec_coords, wavelet, save=True)
    f, residual = Loss(rec, recin, float(model.critical_dt), is_residual=is_residual,
                       misfit=misfit)
    _, v, _ = forward(model, rec_coords, None, residual, save=True, fw=False)
    g = dft_gradient(u, v, float(model.critical_dt), freq_list, dft_sub, ic=ic,
                     m=model.m.astype(np.float64) if ic != "as" else None,
                     inner_grad=inner_grad_np(model) if ic != "as" else None)
    g = g.astype(model.lib._dtype)
    if return_obj:
        return f, g
    return g


def set_fast(on, precision="f32"):
    """Toggle the vectorised constant-density fast path (bit-identical to the generic one)."""
    _lib(precision).oracle_set_fast(int(bool(on)))


def set_threads(n, precision="f32"):
    """Number of OpenMP threads of the oracle loops (returns the value in effect)."""
    return int(_lib(precision).oracle_set_threads(int(n)))


def time_steps(model, nsteps, with_gradient=True, t_sub=1):
    """Bounded CPU-baseline sample (bench.py only): seconds for nsteps fwd (+ nsteps adj+IC)."""
    s = ctypes.c_double(0)
    model.lib.oracle_time_steps(model.handle(), int(nsteps), int(with_gradient), int(t_sub),
                                ctypes.byref(s))
    return s.value
