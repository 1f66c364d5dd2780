"""Mirror of JUDI's per-shot driver above the propagator calls
(src/TimeModeling/Modeling/time_modeling_serial.jl:9-51 `_time_modeling`,
src/TimeModeling/Modeling/python_interface.jl:30-198 `devito_interface`) with the steps that
bracket every call executed by libjudi_b200 on the device:

  time_resample                    src/TimeModeling/Utils/time_utilities.jl:32-48, 84
  remove_padding                   src/TimeModeling/Utils/auxiliaryFunctions.jl:79-89
  limit_model_to_receiver_area     src/TimeModeling/Utils/auxiliaryFunctions.jl:129-173

`Geometry` / `Options` carry only the fields this path reads (Geometry: xloc/yloc/zloc, dt, nt,
t0 -- GeometryStructure.jl; Options: JUDIOptions fields of OptionsStructure.jl that reach the
propagators).
"""
import ctypes
import numpy as np

from . import _lib as L
from . import interface
from .models import Model, get_context, is_device_tensor


class Geometry(object):
    """One source or receiver geometry (one shot): coordinates in metres and the time axis
    `t0:dt:t` in ms (GeometryStructure.jl `GeometryIC`, single-source view)."""

    def __init__(self, xloc, yloc, zloc, dt, t, t0=0.0):
        self.xloc = np.atleast_1d(np.asarray(xloc, dtype=np.float32))
        self.zloc = np.atleast_1d(np.asarray(zloc, dtype=np.float32))
        self.yloc = None if yloc is None else np.atleast_1d(np.asarray(yloc, dtype=np.float32))
        self.dt, self.t, self.t0 = np.float32(dt), np.float32(t), np.float32(t0)
        self.nt = int(np.floor((self.t - self.t0) / self.dt)) + 1

    @property
    def taxis(self):
        return (self.t0, self.dt, self.nt)

    def coords(self, ndim):
        """setup_grid (auxiliaryFunctions.jl:316-326): hcat(x[,y],z); 2-D models drop y."""
        if ndim == 3:
            return np.stack([self.xloc, self.yloc, self.zloc], axis=1).astype(np.float32)
        return np.stack([self.xloc, self.zloc], axis=1).astype(np.float32)


class Options(object):
    """The JUDIOptions fields that reach this path (OptionsStructure.jl:7-60), same defaults."""

    def __init__(self, space_order=8, free_surface=False, limit_m=False, buffer_size=1000.0,
                 optimal_checkpointing=False, subsampling_factor=1, sum_padding=False,
                 dt_comp=None, IC="as", f0=0.015, snapshot_region="padded", nbl=40,
                 frequencies=(), dft_subsampling_factor=1):
        self.space_order = space_order
        self.free_surface = free_surface
        self.limit_m = limit_m
        self.buffer_size = buffer_size
        self.optimal_checkpointing = optimal_checkpointing
        self.subsampling_factor = subsampling_factor
        self.sum_padding = sum_padding
        self.dt_comp = dt_comp
        self.IC = IC
        self.f0 = f0
        self.snapshot_region = snapshot_region
        self.nbl = nbl
        self.frequencies = list(frequencies)          # kHz; non-empty selects on-the-fly DFT
        self.dft_subsampling_factor = dft_subsampling_factor


# ---- time_resample ---------------------------------------------------------------------------
def _resampled_length(t_in, t_new):
    t0i, dti, nti = np.float32(t_in[0]), np.float32(t_in[1]), int(t_in[2])
    t0n, dtn, ntn = np.float32(t_new[0]), np.float32(t_new[1]), int(t_new[2])
    if dtn == dti:
        return nti - int(np.trunc(np.float32(t0n - t0i) / dtn))
    if np.fmod(dtn, dti) == 0:
        rate = int(np.trunc(dtn / dti))
        return (nti + rate - 1) // rate - int(np.trunc(np.float32(t0n - t0i) / dtn))
    return ntn


def time_resample(data, t_in, t_new, out=None, context=None):
    """time_resample(data, t_in::StepRangeLen, t_new::StepRangeLen) (time_utilities.jl:32-48).

    `data`: (nt_in, ntrace) host array, or a torch CUDA tensor given time fastest
    (contiguous (ntrace, nt_in)); `t_in` / `t_new`: (first, step, length).  Returns the
    resampled traces in the same convention (device in -> device out)."""
    ctx = context or get_context()
    nt_out = _resampled_length(t_in, t_new)
    if is_device_tensor(data):
        import torch
        assert data.is_contiguous() and data.dim() == 2
        ntr, nt_in = int(data.shape[0]), int(data.shape[1])
        res = out if out is not None else torch.empty((ntr, nt_out), dtype=torch.float32,
                                                      device=data.device)
        flags = L.DATA_ON_DEVICE | L.OUT_ON_DEVICE
        pin, pout = data.data_ptr(), res.data_ptr()
    else:
        a = np.asarray(data, dtype=np.float32)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        a = np.asfortranarray(a)
        nt_in, ntr = a.shape
        res = np.zeros((nt_out, ntr), dtype=np.float32, order="F")
        flags = 0
        pin, pout = a.ctypes.data, res.ctypes.data
    assert nt_in == int(t_in[2]), "data length does not match its time axis"
    L.check(L.lib().judi_time_resample(ctx.h, pin, nt_in, ntr, float(t_in[0]), float(t_in[1]),
                                       pout, nt_out, float(t_new[0]), float(t_new[1]), flags))
    return res


def time_resample_to_dt(data, geometry, dt_new, context=None):
    """time_resample(data, G_in, dt_new) (time_utilities.jl:15-18): same start and end time."""
    t0, t = geometry.t0, np.float32(geometry.t0 + geometry.dt * (geometry.nt - 1))
    n_new = int(np.floor(np.float32(t - t0) / np.float32(dt_new))) + 1
    return time_resample(data, geometry.taxis, (t0, np.float32(dt_new), n_new), context=context)


def time_resample_to_geometry(data, dt_in, geometry, geometry_in=None, context=None):
    """time_resample(data, dt_in, G_out) (time_utilities.jl:66-69) and, with the source geometry,
    time_resample(data, dt_in, G_in, G_out) (:71-79, what post_process_src calls): the modelled
    traces start at t0 = min(t0 of the source, t0 of the receivers)."""
    nt_in = data.shape[1] if is_device_tensor(data) else np.asarray(data).shape[0]
    t0 = np.float32(0) if geometry_in is None else np.float32(min(geometry_in.t0, geometry.t0))
    return time_resample(data, (t0, np.float32(dt_in), nt_in), geometry.taxis, context=context)


# ---- _maybe_pad_t0 ---------------------------------------------------------------------------
def _jdiv(a, b):
    """Julia's div(a, b) on Float32 (rounds towards zero)."""
    return int(np.trunc(np.float32(a) / np.float32(b)))


def maybe_pad_t0(q, q_geom, d, d_geom, dt):
    """_maybe_pad_t0 (time_utilities.jl:92-139): zero-pad the source and the data so that both cover
    min(t0) .. max(t) when their geometries start or end at different times (SEG-Y data with a
    recording delay).  `d` = None stands for "no data yet" (forward modelling, :153-154): only the
    padded source is returned."""
    only_q = d is None
    if only_q:
        nd = int(np.floor(np.float32(d_geom.t - d_geom.t0) / np.float32(dt))) + 1
        d = np.zeros((nd, 1), dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    d = np.asarray(d, dtype=np.float32)
    z = lambda n, like: np.zeros((max(n, 0), like.shape[1]), dtype=np.float32)
    dt0 = np.float32(d_geom.t0 - q_geom.t0)
    Dt = np.float32(d_geom.t - q_geom.t)
    dsize = q.shape[0] - d.shape[0]
    if dsize == 0 and dt0 == 0 and Dt == 0:
        pass
    elif dsize == 0 and dt0 != 0 and Dt != 0:
        assert np.sign(dt0) == np.sign(Dt), "source and data are shifted in opposite directions"
        pad = _jdiv(abs(dt0), dt)
        if dt0 > 0:
            d, q = np.vstack([z(pad, d), d]), np.vstack([q, z(pad, q)])
        else:
            d, q = np.vstack([d, z(pad, d)]), np.vstack([z(pad, q), q])
    elif dsize != 0:
        ts, te = min(q_geom.t0, d_geom.t0), max(q_geom.t, d_geom.t)
        d = np.vstack([z(_jdiv(d_geom.t0 - ts, dt), d), d, z(_jdiv(te - d_geom.t, dt), d)])
        q = np.vstack([z(_jdiv(q_geom.t0 - ts, dt), q), q, z(_jdiv(te - q_geom.t, dt), q)])
        left = q.shape[0] - d.shape[0]
        if left > 0:
            d = np.vstack([d, z(left, d)])
        elif left < 0:
            q = np.vstack([q, z(-left, q)])
    else:
        raise ValueError("data and source have different t0 %s and t %s and are not compatible in size "
                         "for padding: %s" % ((d_geom.t0, q_geom.t0), (d_geom.t, q_geom.t),
                                              (q.shape[0], d.shape[0])))
    q = np.asfortranarray(q)
    return q if only_q else (q, np.asfortranarray(d))


# ---- remove_padding --------------------------------------------------------------------------
def remove_padding(g, model, true_adjoint=False, out=None):
    """remove_padding(gradient, modelPy.padsizes; true_adjoint) on the device.  `g` is a padded
    grid array (F-ordered numpy (nx[,ny],nz)) or a torch CUDA tensor (x fastest); returns the
    physical-grid array as F-ordered numpy, or fills the CUDA tensor `out`."""
    flags = 0
    if is_device_tensor(g):
        flags |= L.DATA_ON_DEVICE
        pin = g.data_ptr()
    else:
        g = np.asfortranarray(np.asarray(g, dtype=np.float32))
        assert g.shape == model.shape_pml
        pin = g.ctypes.data
    if out is not None and is_device_tensor(out):
        flags |= L.OUT_ON_DEVICE
        res, pout = out, out.data_ptr()
    else:
        res = out if out is not None else np.zeros(model.shape, dtype=np.float32, order="F")
        assert res.flags.f_contiguous and res.shape == model.shape
        pout = res.ctypes.data
    L.check(L.lib().judi_remove_padding(model.h, pin, int(bool(true_adjoint)), pout, flags))
    return res


# ---- limit_model_to_receiver_area ----------------------------------------------------------------
def limit_model_to_receiver_area(src_geom, rec_geom, origin, spacing, shape, buffer, params,
                                 pert=None):
    """auxiliaryFunctions.jl:129-173: crop the model arrays to the x[,y] range of the sources and
    receivers plus `buffer`; returns (origin, shape, params, pert)."""
    ndim = len(shape)
    axes = [("xloc",), ("yloc",)] if ndim == 3 else [("xloc",)]
    inds = []
    for d, (name,) in enumerate(axes):
        o, h, n = np.float32(origin[d]), np.float32(spacing[d]), shape[d]
        loc_r, loc_s = getattr(rec_geom, name), getattr(src_geom, name)
        lo = min(loc_r.min(), loc_s.min())
        hi = max(loc_r.max(), loc_s.max())
        lo = max(o, np.float32(lo - np.float32(buffer)))
        hi = min(np.float32(o + h * np.float32(n - 1)), np.float32(hi + np.float32(buffer)))
        i0 = int(np.floor(np.float32(lo - o) / h)) + 1
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
    return new_origin, new_shape, newp, newpert


# ---- the per-shot driver ---------------------------------------------------------------------
def time_modeling(origin, spacing, shape, params, src_geom, src_data, rec_geom, rec_data, dm, op,
                  options=None, fw=True, illum=False, context=None):
    """_time_modeling (time_modeling_serial.jl:9-51) for one source experiment.

    `params`: dict of model arrays / scalars (m, rho, epsilon, delta, theta, phi);
    `op` in {"forward", "adjoint", "born", "adjoint_born"} as in the reference.
    Returns the shot record resampled to the receiver geometry (nt_rec, nrec), or the gradient
    on the (full, un-limited) physical grid.
    """
    o = options or Options()
    ndim = len(shape)
    if op == "born" and (dm is None or not np.any(np.asarray(dm))):
        # time_modeling_serial.jl:22-24: J*0 returns zeros without propagating
        return np.zeros((rec_geom.nt, rec_geom.xloc.size), dtype=np.float32, order="F")
    full_shape, full_origin = tuple(shape), tuple(origin)
    if o.limit_m:
        origin, shape, params, dm = limit_model_to_receiver_area(
            src_geom, rec_geom, origin, spacing, shape, o.buffer_size, params, pert=dm)
    kw = dict(params)
    if dm is not None:
        kw["dm"] = np.asarray(dm, dtype=np.float32).reshape(shape)
    model = Model(origin, spacing, shape, space_order=o.space_order, nbl=o.nbl,
                  fs=o.free_surface, dt=o.dt_comp, context=context, **kw)
    dt_comp = np.float32(model.critical_dt)
    q_in = time_resample_to_dt(src_data, src_geom, dt_comp, context=model.ctx)
    sc, rc = src_geom.coords(ndim), rec_geom.coords(ndim)
    if op != "adjoint_born":
        q_in = maybe_pad_t0(q_in, src_geom, None, rec_geom, dt_comp)     # python_interface.jl:34, 88
    if op in ("forward", "adjoint") and rec_data is None:
        # `op` only selects the post-processing; the direction is `fw` alone, as `_prop_fw(F)`
        # hands it to devito_interface (propagation.jl:9,52-54; python_interface.jl:42)
        rec, _ = interface.forward_rec(model, sc, q_in, rc, f0=o.f0, illum=illum, fw=fw)
        return time_resample_to_geometry(rec, dt_comp, rec_geom, geometry_in=src_geom, context=model.ctx)
    if op == "born":
        rec, _ = interface.born_rec(model, sc, q_in, rc, ic=o.IC, f0=o.f0, illum=illum, fw=fw)
        return time_resample_to_geometry(rec, dt_comp, rec_geom, geometry_in=src_geom, context=model.ctx)
    if op == "adjoint_born":
        d_in = time_resample_to_dt(rec_data, rec_geom, dt_comp, context=model.ctx)
        q_in, d_in = maybe_pad_t0(q_in, src_geom, d_in, rec_geom, dt_comp)   # python_interface.jl:109
        g, _, _ = interface.J_adjoint(model, sc, q_in, rc, d_in, is_residual=True,
                                      t_sub=o.subsampling_factor,
                                      checkpointing=o.optimal_checkpointing, ic=o.IC, f0=o.f0,
                                      freq_list=o.frequencies or None,
                                      dft_sub=o.dft_subsampling_factor, fw=fw,
                                      snapshot_region=o.snapshot_region)
        g = remove_padding(g, model, true_adjoint=o.sum_padding)
        if o.limit_m and tuple(shape) != full_shape:
            # the gradient of a limited model is embedded in the full grid (PhysicalParameter
            # arithmetic with origins, ModelStructure.jl:186-240)
            full = np.zeros(full_shape, dtype=np.float32, order="F")
            start = [int(round(float((origin[d] - full_origin[d]) / spacing[d])))
                     for d in range(ndim)]
            full[tuple(slice(s, s + n) for s, n in zip(start, shape))] = g
            return full
        return g
    raise ValueError("unknown operator %r" % (op,))


def multi_src_fg(origin, spacing, shape, params, src_geom, src_data, rec_geom, rec_data, dm=None,
                 options=None, nlind=False, lin=False, misfit=None, illum=False, context=None):
    """_multi_src_fg (misfit_fg.jl:8-91) for ONE source experiment given on its own geometries'
    time axes: limit the model, build it, resample source and data to the computational time step,
    pad for different recording delays, `J_adjoint(..., return_obj=True, is_residual=False)`, crop
    (or fold) the gradient.  Returns (f, gradient on the full physical grid[, Iu, Iv]) -- the
    per-shot work of `fwi_objective` (`lin=False`) and `lsrtm_objective` (`lin=True`, with `dm`)."""
    o = options or Options()
    ndim = len(shape)
    full_shape, full_origin = tuple(shape), tuple(origin)
    if o.limit_m:
        origin, shape, params, dm = limit_model_to_receiver_area(
            src_geom, rec_geom, origin, spacing, shape, o.buffer_size, params, pert=dm)
    kw = dict(params)
    if dm is not None:
        kw["dm"] = np.asarray(dm, dtype=np.float32).reshape(shape)
    model = Model(origin, spacing, shape, space_order=o.space_order, nbl=o.nbl,
                  fs=o.free_surface, dt=o.dt_comp, context=context, **kw)
    dt_comp = np.float32(model.critical_dt)
    q_in = time_resample_to_dt(src_data, src_geom, dt_comp, context=model.ctx)
    d_in = time_resample_to_dt(rec_data, rec_geom, dt_comp, context=model.ctx)
    q_in, d_in = maybe_pad_t0(q_in, src_geom, d_in, rec_geom, dt_comp)
    sc, rc = src_geom.coords(ndim), rec_geom.coords(ndim)
    f, g, Iu, Iv = interface.J_adjoint(model, sc, q_in, rc, d_in, is_residual=False,
                                       t_sub=o.subsampling_factor, checkpointing=o.optimal_checkpointing,
                                       freq_list=o.frequencies or None, dft_sub=o.dft_subsampling_factor,
                                       ic=o.IC, f0=o.f0, born_fwd=lin, nlind=nlind, misfit=misfit,
                                       return_obj=True, illum=illum, snapshot_region=o.snapshot_region)

    def post(a, fold):
        a = remove_padding(a, model, true_adjoint=fold)
        if o.limit_m and tuple(shape) != full_shape:
            full = np.zeros(full_shape, dtype=np.float32, order="F")
            start = [int(round(float((origin[d] - full_origin[d]) / spacing[d]))) for d in range(ndim)]
            full[tuple(slice(s_, s_ + n) for s_, n in zip(start, shape))] = a
            return full
        return np.array(a, copy=True, order="F")
    out = (np.float32(f), post(g, o.sum_padding))
    if illum:
        out += (post(Iu, False), post(Iv, False))
    return out


def time_modeling_w(origin, spacing, shape, params, weights, wavelet, rec_geom, rec_data, dm, op,
                    options=None, fw=True, illum=False, context=None):
    """The extended-source methods of `devito_interface` (python_interface.jl:124-198) behind
    `_time_modeling`: `weights` are the judiWeights of one experiment on the physical grid,
    `wavelet` the judiLRWF time signature sampled on the receiver geometry's time axis.

      op = "forward"        d   = Pr F Pw' w           -> (nt_rec, nrec)        (:125-140)
      op = "adjoint"        w   = Pw F' Pr' d          -> physical grid          (:143-158)
      op = "born"           d   = Jw dm                -> (nt_rec, nrec)        (:161-176)
      op = "adjoint_born"   g   = Jw' d                -> physical grid          (:179-198)
    """
    o = options or Options()
    ndim = len(shape)
    kw = dict(params)
    if dm is not None:
        kw["dm"] = np.asarray(dm, dtype=np.float32).reshape(shape)
    model = Model(origin, spacing, shape, space_order=o.space_order, nbl=o.nbl,
                  fs=o.free_surface, dt=o.dt_comp, context=context, **kw)
    dt_comp = np.float32(model.critical_dt)
    rc = rec_geom.coords(ndim)
    q_in = time_resample_to_dt(wavelet, rec_geom, dt_comp, context=model.ctx)
    wpad = None
    if weights is not None:
        # pad_array(reshape(weights, shape), modelPy.padsizes; mode=:zeros)
        wpad = np.asfortranarray(np.pad(np.asarray(weights, dtype=np.float32).reshape(shape),
                                        model.padsizes, mode="constant"))
    if op == "forward":
        rec, _ = interface.forward_rec_w(model, wpad, q_in, rc, f0=o.f0, illum=illum, fw=fw)
        return time_resample_to_geometry(rec, dt_comp, rec_geom, context=model.ctx)
    if op == "born":
        rec, _ = interface.born_rec_w(model, wpad, q_in, rc, ic=o.IC, f0=o.f0, illum=illum)
        return time_resample_to_geometry(rec, dt_comp, rec_geom, context=model.ctx)
    d_in = time_resample_to_dt(rec_data, rec_geom, dt_comp, context=model.ctx)
    q_in, d_in = maybe_pad_t0(q_in, rec_geom, d_in, rec_geom, dt_comp)
    if op == "adjoint":
        w, _ = interface.adjoint_w(model, rc, d_in, q_in, f0=o.f0, illum=illum, fw=fw)
        return np.array(remove_padding(w, model, true_adjoint=False), copy=True, order="F")
    if op == "adjoint_born":
        g, _, _ = interface.J_adjoint(model, None, q_in, rc, d_in, is_residual=True, ws=wpad,
                                      t_sub=o.subsampling_factor, checkpointing=o.optimal_checkpointing,
                                      freq_list=o.frequencies or None, dft_sub=o.dft_subsampling_factor,
                                      ic=o.IC, f0=o.f0, snapshot_region=o.snapshot_region)
        return np.array(remove_padding(g, model, true_adjoint=o.sum_padding), copy=True, order="F")
    raise ValueError("unknown operator %r" % (op,))


def time_modeling_wf(origin, spacing, shape, params, src_geom, src_data, rec_geom, options=None,
                     fw=True, illum=False, context=None):
    """The full-wavefield methods of `devito_interface` (python_interface.jl:44-81):

      src_geom given, rec_geom None    u     = F Ps' q      wavefield history (nt, padded grid), :45-56
      src_geom None,  rec_geom given   d_obs = Pr F u       shot record on the receiver time axis, :59-69
      both None                        u_out = F u_in       wavefield history, :72-80

    A wavefield source `src_data` is (nt, padded grid) on the computational time axis, as a
    judiWavefield carries it; wavefield results come back the same way with their dt (what
    `post_process(..., Val(:forward))` wraps into a judiWavefield)."""
    o = options or Options()
    ndim = len(shape)
    model = Model(origin, spacing, shape, space_order=o.space_order, nbl=o.nbl,
                  fs=o.free_surface, dt=o.dt_comp, context=context, **dict(params))
    dt_comp = np.float32(model.critical_dt)
    if src_geom is not None:
        q_in = time_resample_to_dt(src_data, src_geom, dt_comp, context=model.ctx)
        u, _ = interface.forward_no_rec(model, src_geom.coords(ndim), q_in, f0=o.f0, illum=illum, fw=fw)
        return u, dt_comp
    if rec_geom is not None:
        rec, _ = interface.forward_wf_src(model, src_data, rec_geom.coords(ndim), f0=o.f0, illum=illum, fw=fw)
        return time_resample_to_geometry(rec, dt_comp, rec_geom, context=model.ctx)
    u, _ = interface.forward_wf_src_norec(model, src_data, f0=o.f0, illum=illum, fw=fw)
    return u, dt_comp
