"""Host-side mirror of the reference's `pysource.interface` (src/pysource/interface.py): the
functions the Julia driver calls through `wrapcall_data(ac.<fn>, modelPy, ...)`
(src/TimeModeling/Modeling/python_interface.jl:30-198, misfit_fg.jl:69-73), with the same names,
argument meaning, defaults and return tuples -- but executed by libjudi_b200 on a B200.

Conventions: coordinates (npoint, ndim) in metres; wavelets / shot records (nt, ntrace);
grids (nx[,ny],nz).  Outputs are F-ordered numpy arrays, i.e. exactly the memory layout Julia
uses, so the `ccall` twin of these wrappers needs no copies.  When the wavelet / data are torch
CUDA tensors (given time-fastest: contiguous (ntrace, nt)) everything stays in HBM and the
results are written into the caller's `*_out` device tensors.

Out of this backend's scope (raise NotImplementedError): `wri_func`.
"""
import ctypes
import numpy as np

from . import _lib as L
from .models import is_device_tensor


def _coords(c, ndim):
    """(npoint, ndim) -> SoA float32 (Julia's hcat(x[,y],z) column-major memory)."""
    c = np.asarray(c, dtype=np.float32)
    if c.ndim == 1:
        c = c.reshape(1, -1)
    assert c.shape[1] == ndim, "coordinates must be (npoint, %d)" % ndim
    return np.asfortranarray(c)


def _traces(a):
    """(nt, ntrace) host array -> time-fastest float32 memory; returns (array, nt, ntrace)."""
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    a = np.asfortranarray(a)
    return a, a.shape[0], a.shape[1]


def _dev_traces(t):
    assert t.is_contiguous() and t.dim() == 2, "device traces must be contiguous (ntrace, nt)"
    return t, int(t.shape[1]), int(t.shape[0])


def _ptr(a):
    if a is None:
        return None
    if is_device_tensor(a):
        return a.data_ptr()
    return a.ctypes.data


def _with_f0(fn):
    """`f0` (kHz) only matters for visco-acoustic models: it sets the relaxation times of the next
    call (propagators.py:46).  None keeps the model's value (reference default 0.015)."""
    import functools
    import inspect
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        bound = sig.bind(*args, **kwargs)
        model = bound.arguments.get("model")
        if model is not None and hasattr(model, "set_f0"):
            model.set_f0(bound.arguments.get("f0"))
        return fn(*args, **kwargs)
    return wrapper


def _grid_out(model, out):
    if out is not None:
        return out
    return np.zeros(model.shape_pml, dtype=np.float32, order="F")


# Forward wrappers Pr*F*Ps'*q
@_with_f0
def forward_rec(model, src_coords, wavelet, rec_coords, f0=None, illum=False, fw=True,
                rec_out=None, illum_out=None):
    """Modeling of a point source with receivers Pr*F*Ps^T*q (interface.py:17-47).

    Returns (shot record (nt, nrec), illumination or None).  With fw=False the adjoint
    propagator is run (time reversed), as the reference does for F'.
    """
    dev = is_device_tensor(wavelet)
    sc = _coords(src_coords, model.dim)
    rc = _coords(rec_coords, model.dim)
    if dev:
        q, nt, nsrc = _dev_traces(wavelet)
        assert rec_out is not None and is_device_tensor(rec_out)
        rec = rec_out
        I = illum_out if illum else None
    else:
        q, nt, nsrc = _traces(wavelet)
        rec = np.zeros((nt, rc.shape[0]), dtype=np.float32, order="F")
        I = _grid_out(model, illum_out) if illum else None
    assert nsrc == sc.shape[0], "one wavelet column per source point"
    L.check(L.lib().judi_forward_rec(model.h, int(bool(fw)), nsrc, sc.ctypes.data, _ptr(q), nt,
                                     rc.shape[0], rc.ctypes.data, _ptr(rec), _ptr(I),
                                     L.DATA_ON_DEVICE if dev else 0))
    return rec, I


# F*Ps'*q
@_with_f0
def forward_no_rec(model, src_coords, wavelet, f0=None, illum=False, fw=True):
    """Forward modeling of a point source without receiver (interface.py:83-111): returns the
    wavefield history (nt, nx[,ny],nz) on the padded grid and the illumination."""
    sc = _coords(src_coords, model.dim)
    q, nt, nsrc = _traces(wavelet)
    u = np.zeros(model.shape_pml + (nt,), dtype=np.float32, order="F")  # memory: time slowest
    I = _grid_out(model, None) if illum else None
    L.check(L.lib().judi_forward_no_rec(model.h, int(bool(fw)), nsrc, sc.ctypes.data, _ptr(q), nt,
                                        _ptr(u), _ptr(I), 0))
    return np.moveaxis(u, -1, 0), I


# Linearized modeling d/dm (Pr*F*Ps'*q)
def _reverse_time(a):
    """(nt, ntrace) host traces with the time axis reversed (F-ordered copy)."""
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a[::-1])


def _require_time_reversible(model, traces, what):
    if getattr(model, "is_viscoacoustic", False):
        raise NotImplementedError("%s with fw=False on a visco-acoustic model (its adjoint scheme is "
                                  "not the time mirror of the forward one)" % what)
    if is_device_tensor(traces):
        raise NotImplementedError("%s with fw=False takes host traces" % what)


@_with_f0
def born_rec(model, src_coords, wavelet, rec_coords, ic="as", f0=None, illum=False, fw=True,
             rec_out=None, illum_out=None):
    """Linearized (Born) modeling of a point source for the model's dm (interface.py:208-241).

    fw=False (the Jacobian of the adjoint propagator F', `born(..., fw=False)`): for every
    propagator but the visco-acoustic one the adjoint scheme is the forward scheme run in reversed
    time (`u.dt.T` is the time mirror of `u.dt`, kernels.py:61-69; checked bitwise in tests/), so
    the call is the forward Born modelling of the time-reversed wavelet, reversed."""
    if not fw:
        _require_time_reversible(model, wavelet, "born_rec")
        rec, I = born_rec(model, src_coords, _reverse_time(wavelet), rec_coords, ic=ic, f0=f0,
                          illum=illum, fw=True, illum_out=illum_out)
        return _reverse_time(rec), I
    dev = is_device_tensor(wavelet)
    sc = _coords(src_coords, model.dim)
    rc = _coords(rec_coords, model.dim)
    if dev:
        q, nt, nsrc = _dev_traces(wavelet)
        assert rec_out is not None and is_device_tensor(rec_out)
        rec = rec_out
        I = illum_out if illum else None
    else:
        q, nt, nsrc = _traces(wavelet)
        rec = np.zeros((nt, rc.shape[0]), dtype=np.float32, order="F")
        I = _grid_out(model, illum_out) if illum else None
    L.check(L.lib().judi_born_rec(model.h, nsrc, sc.ctypes.data, _ptr(q), nt, rc.shape[0],
                                  rc.ctypes.data, L.IC_CODES[ic], _ptr(rec), _ptr(I),
                                  L.DATA_ON_DEVICE if dev else 0))
    return rec, I


@_with_f0
def J_adjoint(model, src_coords, wavelet, rec_coords, recin, is_residual=False,
              checkpointing=False, n_checkpoints=None, t_sub=1, return_obj=False, freq_list=None,
              dft_sub=None, ic="as", illum=False, ws=None, f0=None, born_fwd=False, nlind=False,
              misfit=None, fw=True, snapshot_region="padded", grad_out=None, f_out=None,
              illum_out=None, accumulate=False):
    """Adjoint Jacobian on a shot record (interface.py:279-345): standard zero-lag
    cross-correlation (:411-470) or checkpointed (:473-562).

    Returns `(f, g, Iu, Iv)` if return_obj else `(g, Iu, Iv)`, g on the padded grid (the caller
    crops it with remove_padding, misfit_fg.jl:77).  `misfit(dsyn, dobs) -> (f, residual)` is
    the optional user misfit (misfit_fg.jl:55-64).  Extra keywords of this backend:
    `snapshot_region` ('padded' = reference semantics | 'physical'), and for device-resident
    runs `grad_out`, `f_out`, `illum_out=(Iu, Iv)`, `accumulate`.
    """
    if not fw:
        # Jacobian of the adjoint propagator (`gradient(..., fw=False)`): the time mirror of the
        # forward problem (see born_rec).  Exact when the snapshots are taken on every step (or on a
        # pattern that is symmetric under time reversal); the on-the-fly DFT keeps its own path.
        _require_time_reversible(model, wavelet, "J_adjoint")
        nt_ = np.asarray(wavelet).shape[0]
        if freq_list is not None or not (t_sub == 1 or checkpointing or (nt_ - 1) % t_sub == 0):
            raise NotImplementedError("J_adjoint with fw=False needs t_sub = 1 (or checkpointing, or "
                                      "(nt - 1) % t_sub == 0) and no frequency list")
        if misfit is not None:
            user_misfit = misfit

            def misfit(dsyn, dobs):     # the callback sees the traces in their own time order
                fv, r = user_misfit(dsyn[::-1], dobs[::-1])
                return fv, np.asarray(r)[::-1]
        return J_adjoint(model, src_coords, _reverse_time(wavelet), rec_coords, _reverse_time(recin),
                         is_residual=is_residual, checkpointing=checkpointing,
                         n_checkpoints=n_checkpoints, t_sub=t_sub, return_obj=return_obj, ic=ic,
                         illum=illum, ws=ws, f0=f0, born_fwd=born_fwd, nlind=nlind, misfit=misfit,
                         fw=True, snapshot_region=snapshot_region, grad_out=grad_out, f_out=f_out,
                         illum_out=illum_out, accumulate=accumulate)
    dev = is_device_tensor(wavelet)
    sc = _coords(src_coords, model.dim) if src_coords is not None else None
    rc = _coords(rec_coords, model.dim)
    wsp = _weights(model, ws) if ws is not None else None
    o = L.JudiJopts()
    o.is_residual = int(bool(is_residual))
    o.checkpointing = int(bool(checkpointing))
    o.n_checkpoints = int(n_checkpoints) if n_checkpoints else 0
    o.t_sub = int(t_sub)
    o.ic = L.IC_CODES[ic]
    o.born_fwd = int(bool(born_fwd))
    o.nlind = int(bool(nlind))
    o.illum = int(bool(illum))
    o.snapshot_region = L.SNAP_PHYSICAL if snapshot_region == "physical" else L.SNAP_PADDED
    o.ws = _ptr(wsp)
    freqs = None
    if freq_list is not None and len(freq_list) > 0:
        # J_adjoint_freq (interface.py:348-408): frequencies in kHz, dft_sub in time steps
        freqs = np.ascontiguousarray(freq_list, dtype=np.float32)
        o.nfreq = int(freqs.size)
        o.dft_sub = int(dft_sub) if dft_sub else 1
        o.freq_list = freqs.ctypes.data
    flags = (L.DATA_ON_DEVICE if dev else 0) | (L.GRAD_ACCUMULATE if accumulate else 0)
    Iu = Iv = None
    if dev:
        q, nt, nsrc = _dev_traces(wavelet)
        d, ntd, nrec = _dev_traces(recin)
        assert grad_out is not None and is_device_tensor(grad_out)
        g = grad_out
        fbuf = f_out
        if illum:
            Iu, Iv = illum_out
    elif is_device_tensor(grad_out):
        # host traces, outputs summed in HBM (one copy back after the reduction over shots/ranks)
        q, nt, nsrc = _traces(wavelet)
        d, ntd, nrec = _traces(recin)
        g, fbuf = grad_out, f_out
        assert fbuf is None or is_device_tensor(fbuf)
        dev = True
        flags |= L.OUT_ON_DEVICE
        if illum:
            Iu, Iv = illum_out
    else:
        q, nt, nsrc = _traces(wavelet)
        d, ntd, nrec = _traces(recin)
        g = grad_out if grad_out is not None else _grid_out(model, None)
        fbuf = ctypes.c_float(0.0 if not accumulate or f_out is None else float(f_out))
        if illum:
            Iu, Iv = _grid_out(model, None), _grid_out(model, None)
    assert ntd == nt and nrec == rc.shape[0], "data shape does not match geometry / wavelet"

    cb = ctypes.cast(None, L.MISFIT_FN)
    if misfit is not None:
        def _cb(psyn, pobs, pres, nt_, nrec_, _user):
            shp = (nt_, nrec_)
            dsyn = np.ctypeslib.as_array(psyn, shape=(nrec_, nt_)).T
            dobs = np.ctypeslib.as_array(pobs, shape=(nrec_, nt_)).T
            res = np.ctypeslib.as_array(pres, shape=(nrec_, nt_)).T
            fval, r = misfit(dsyn, dobs)
            res[...] = np.asarray(r, dtype=np.float32).reshape(shp)
            return float(fval)
        cb = L.MISFIT_FN(_cb)

    fptr = None
    if return_obj or fbuf is not None:
        fptr = _ptr(fbuf) if dev else ctypes.addressof(fbuf)
    if sc is None:
        nsrc = 0
    L.check(L.lib().judi_J_adjoint(model.h, nsrc, sc.ctypes.data if sc is not None else None,
                                   _ptr(q), nt, nrec,
                                   rc.ctypes.data, _ptr(d), ctypes.byref(o),
                                   ctypes.cast(cb, ctypes.c_void_p), None, fptr, _ptr(g),
                                   _ptr(Iu), _ptr(Iv), flags))
    if return_obj:
        f = fbuf if dev else np.float32(fbuf.value)
        return f, g, Iu, Iv
    return g, Iu, Iv


def fg_multishot(model, shots, grad_out, f_out, is_residual=False, checkpointing=False,
                 n_checkpoints=None, t_sub=1, ic="as", born_fwd=False, nlind=False,
                 snapshot_region="padded", accumulate=False, f0=None):
    """Objective and gradient of several source experiments in ONE library call, summed in HBM
    (`judi_fg_multishot`: the per-worker part of multi_src_fg!, propagation.jl:93-107).

    `shots`: sequence of (src_coords, wavelet, rec_coords, dobs) with host arrays; `grad_out`
    (padded grid) and `f_out` (one float) are device tensors that receive the sums."""
    model.set_f0(f0)
    assert is_device_tensor(grad_out) and is_device_tensor(f_out)
    o = L.JudiJopts()
    o.is_residual = int(bool(is_residual))
    o.checkpointing = int(bool(checkpointing))
    o.n_checkpoints = int(n_checkpoints) if n_checkpoints else 0
    o.t_sub = int(t_sub)
    o.ic = L.IC_CODES[ic]
    o.born_fwd = int(bool(born_fwd))
    o.nlind = int(bool(nlind))
    o.snapshot_region = L.SNAP_PHYSICAL if snapshot_region == "physical" else L.SNAP_PADDED
    arr = (L.JudiShot * len(shots))()
    keep = []
    for k, (src_coords, wavelet, rec_coords, recin) in enumerate(shots):
        sc, rc = _coords(src_coords, model.dim), _coords(rec_coords, model.dim)
        q, nt, nsrc = _traces(wavelet)
        d, ntd, nrec = _traces(recin)
        assert ntd == nt and nrec == rc.shape[0], "data shape does not match geometry / wavelet"
        keep += [sc, rc, q, d]
        arr[k].nsrc, arr[k].src_coords, arr[k].wavelet = nsrc, sc.ctypes.data, q.ctypes.data
        arr[k].nt, arr[k].nrec, arr[k].rec_coords, arr[k].recin = nt, nrec, rc.ctypes.data, d.ctypes.data
    flags = L.OUT_ON_DEVICE | (L.GRAD_ACCUMULATE if accumulate else 0)
    L.check(L.lib().judi_fg_multishot(model.h, len(shots), arr, ctypes.byref(o), None, None,
                                      f_out.data_ptr(), grad_out.data_ptr(), flags))
    return f_out, grad_out


def fg_multishot_reduced(model, shots, sum_padding=False, is_residual=False, checkpointing=False,
                         n_checkpoints=None, t_sub=1, ic="as", born_fwd=False, nlind=False,
                         snapshot_region="padded", f0=None):
    """multi_src_fg! + reduce! + remove_padding of one worker in ONE library call
    (`judi_fg_multishot_reduced`; propagation.jl:93-107, distributed.jl:60-95, misfit_fg.jl:77):
    this rank's `shots` (possibly none) are summed in HBM, cropped on the device and all-reduced
    over the context's communicator.  Host arrays in, (f, physical-grid gradient) of ALL ranks' shots
    out -- the call `julia/B200Backend.jl:b200_fg_multishot_reduced` makes."""
    model.set_f0(f0)
    o = L.JudiJopts()
    o.is_residual = int(bool(is_residual))
    o.checkpointing = int(bool(checkpointing))
    o.n_checkpoints = int(n_checkpoints) if n_checkpoints else 0
    o.t_sub = int(t_sub)
    o.ic = L.IC_CODES[ic]
    o.born_fwd = int(bool(born_fwd))
    o.nlind = int(bool(nlind))
    o.snapshot_region = L.SNAP_PHYSICAL if snapshot_region == "physical" else L.SNAP_PADDED
    arr = (L.JudiShot * max(len(shots), 1))()
    keep = []
    for k, (src_coords, wavelet, rec_coords, recin) in enumerate(shots):
        sc, rc = _coords(src_coords, model.dim), _coords(rec_coords, model.dim)
        q, nt, nsrc = _traces(wavelet)
        d, ntd, nrec = _traces(recin)
        assert ntd == nt and nrec == rc.shape[0], "data shape does not match geometry / wavelet"
        keep += [sc, rc, q, d]
        arr[k].nsrc, arr[k].src_coords, arr[k].wavelet = nsrc, sc.ctypes.data, q.ctypes.data
        arr[k].nt, arr[k].nrec, arr[k].rec_coords, arr[k].recin = nt, nrec, rc.ctypes.data, d.ctypes.data
    g = np.zeros(model.shape, dtype=np.float32, order="F")
    f = ctypes.c_float(0)
    L.check(L.lib().judi_fg_multishot_reduced(model.h, len(shots), arr, ctypes.byref(o), None, None,
                                              int(bool(sum_padding)), ctypes.addressof(f),
                                              g.ctypes.data, 0))
    return np.float32(f.value), g


def _weights(model, w):
    """Extended-source weights on the padded grid (python_interface.jl:129 zero-pads them)."""
    if is_device_tensor(w):
        return w
    w = np.asarray(w, dtype=np.float32)
    if w.shape == model.shape:
        w = np.pad(w, model.padsizes, mode="constant")
    assert w.shape == model.shape_pml, "weights must have the (padded) model shape"
    return np.asfortranarray(w)


# Pr*F*Pw'*w
@_with_f0
def forward_rec_w(model, weight, wavelet, rec_coords, f0=None, illum=False, fw=True):
    """Forward modeling of an extended source with receivers Pr*F*Pw^T*w (interface.py:50-80)."""
    rc = _coords(rec_coords, model.dim)
    q, nt, _ = _traces(wavelet)
    w = _weights(model, weight)
    rec = np.zeros((nt, rc.shape[0]), dtype=np.float32, order="F")
    I = _grid_out(model, None) if illum else None
    L.check(L.lib().judi_forward_rec_w(model.h, int(bool(fw)), _ptr(w), _ptr(q), nt, rc.shape[0],
                                       rc.ctypes.data, _ptr(rec), _ptr(I), 0))
    return rec, I


# Pw*F'*Pr'*d_obs
@_with_f0
def adjoint_w(model, rec_coords, data, wavelet, f0=None, illum=False, fw=True):
    """Adjoint modeling of a shot record for an extended source Pw*F^T*Pr^T*d (interface.py:174-204):
    returns the spatial distribution on the padded grid."""
    rc = _coords(rec_coords, model.dim)
    d, nt, nrec = _traces(data)
    q, ntq, _ = _traces(wavelet)
    assert ntq == nt and nrec == rc.shape[0]
    w = _grid_out(model, None)
    I = _grid_out(model, None) if illum else None
    L.check(L.lib().judi_adjoint_w(model.h, int(bool(fw)), nrec, rc.ctypes.data, _ptr(d), _ptr(q),
                                   nt, _ptr(w), _ptr(I), 0))
    return w, I


# d/dm (Pr*F*Pw'*w)
@_with_f0
def born_rec_w(model, weight, wavelet, rec_coords, ic="as", f0=None, illum=False, fw=True):
    """Linearized modeling of an extended source (interface.py:244-276)."""
    rc = _coords(rec_coords, model.dim)
    q, nt, _ = _traces(wavelet)
    w = _weights(model, weight)
    rec = np.zeros((nt, rc.shape[0]), dtype=np.float32, order="F")
    I = _grid_out(model, None) if illum else None
    L.check(L.lib().judi_born_rec_w(model.h, _ptr(w), _ptr(q), nt, rc.shape[0], rc.ctypes.data,
                                    L.IC_CODES[ic], _ptr(rec), _ptr(I), 0))
    return rec, I


def _wf_src(model, u):
    u = np.asarray(u, dtype=np.float32)
    assert u.shape[1:] == model.shape_pml, "wavefield source must be (nt, padded grid)"
    # memory: time slowest, x fastest
    return np.asfortranarray(np.moveaxis(u, 0, -1)), u.shape[0]


# Pr*F*u
@_with_f0
def forward_wf_src(model, u, rec_coords, f0=None, illum=False, fw=True):
    """Forward modeling of a full wavefield source Pr*F*u (interface.py:114-142)."""
    rc = _coords(rec_coords, model.dim)
    us, nt = _wf_src(model, u)
    rec = np.zeros((nt, rc.shape[0]), dtype=np.float32, order="F")
    I = _grid_out(model, None) if illum else None
    L.check(L.lib().judi_forward_wf_src(model.h, int(bool(fw)), _ptr(us), nt, rc.shape[0],
                                        rc.ctypes.data, _ptr(rec), _ptr(I), 0))
    return rec, I


# F*u
@_with_f0
def forward_wf_src_norec(model, u, f0=None, illum=False, fw=True):
    """Forward modeling of a full wavefield source without receivers F*u (interface.py:145-171)."""
    us, nt = _wf_src(model, u)
    out = np.zeros(model.shape_pml + (nt,), dtype=np.float32, order="F")
    I = _grid_out(model, None) if illum else None
    L.check(L.lib().judi_forward_wf_src_norec(model.h, int(bool(fw)), _ptr(us), nt, _ptr(out),
                                              _ptr(I), 0))
    return np.moveaxis(out, -1, 0), I


def wri_func(*a, **k):
    """Wavefield-reconstruction inversion (interface.py:568) is outside this backend's hot path."""
    raise NotImplementedError("wri_func is outside the B200 backend's hot path (see DESIGN.md)")
