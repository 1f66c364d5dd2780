"""Host-side mirror of the reference's `pysource.models.Model` (src/pysource/models.py:123-527)
for the B200 backend: same constructor arguments and the attributes the Julia driver reads
(`critical_dt`, `padsizes`, `shape`, `dim`, `spacing`, `origin`; python_interface.jl:33,38,
time_modeling_serial.jl:59,70-71, misfit_fg.jl:40,77).  All arithmetic happens in libjudi_b200.

Arrays follow pysource's convention, shape (nx[,ny],nz).  Julia arrays arrive as F-ordered
numpy arrays and are passed to the C-ABI without a copy.  Device-resident inputs are torch CUDA
tensors given x-fastest, i.e. contiguous with shape (nz[,ny],nx).
"""
import ctypes
import numpy as np

from . import _lib as L

_contexts = {}


class Context(object):
    """One per (process, GPU): owns the CUDA stream and memory pool of the backend."""

    def __init__(self, device=0):
        self.device = int(device)
        self.h = L.check_ptr(L.lib().judi_ctx_create(self.device))

    def set_option(self, key, value):
        L.check(L.lib().judi_ctx_set_option(self.h, key.encode(), int(value)))

    def synchronize(self):
        L.check(L.lib().judi_ctx_synchronize(self.h))

    @property
    def stream(self):
        return L.lib().judi_ctx_stream(self.h)

    @property
    def launch_count(self):
        return int(L.lib().judi_ctx_launch_count(self.h))

    def stencil_stats(self):
        ms, n, pts = ctypes.c_double(0), ctypes.c_longlong(0), ctypes.c_double(0)
        L.check(L.lib().judi_ctx_stencil_stats(self.h, ctypes.byref(ms), ctypes.byref(n),
                                               ctypes.byref(pts)))
        out = {"ms": ms.value, "launches": n.value, "points": pts.value, "kinds": {}}
        for kind, name in enumerate(("plain", "save", "ic", "born")):
            L.check(L.lib().judi_ctx_stencil_stats_kind(self.h, kind, ctypes.byref(ms),
                                                        ctypes.byref(n)))
            if n.value:
                out["kinds"][name] = {"ms": ms.value, "launches": n.value}
        return out

    def reset_stats(self):
        L.check(L.lib().judi_ctx_reset_stats(self.h))


def host_register(a):
    """Page-lock a host array that outlives many shots (model, observed data, gradient buffer):
    later copies of it run at PCIe speed.  Returns the array."""
    a = np.asarray(a)
    L.check(L.lib().judi_host_register(ctypes.c_void_p(a.ctypes.data), a.nbytes))
    return a


def host_unregister(a):
    L.check(L.lib().judi_host_unregister(ctypes.c_void_p(np.asarray(a).ctypes.data)))


def get_context(device=None):
    """Context of `device` (default: LOCAL_RANK, else 0), created on first use."""
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def is_device_tensor(a):
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda") and bool(a.is_cuda)


def _as_param(field, shape, keep):
    """Translate array | scalar | None into a judi_param (+ whether it lives on the device)."""
    p = L.JudiParam()
    if field is None:
        p.kind = L.PARAM_NONE
        return p, False
    if is_device_tensor(field):
        assert tuple(field.shape) == tuple(reversed(shape)) and field.is_contiguous(), \
            "device tensors must be contiguous with shape (nz[,ny],nx)"
        import torch
        assert field.dtype == torch.float32
        keep.append(field)
        p.kind = L.PARAM_ARRAY
        p.data = field.data_ptr()
        return p, True
    if np.ndim(field) == 0:
        p.kind = L.PARAM_SCALAR
        p.value = float(field)
        return p, False
    a = np.asarray(field)
    if a.shape != tuple(shape):
        raise ValueError("parameter of shape %s does not match the model shape %s"
                         % (a.shape, tuple(shape)))
    a = np.asfortranarray(a, dtype=np.float32)   # x fastest; no copy for Julia arrays
    keep.append(a)
    p.kind = L.PARAM_ARRAY
    p.data = a.ctypes.data
    return p, False


class Model(object):
    """The physical model used in seismic inversion, resident in HBM.

    Same signature as pysource `Model(origin, spacing, shape, space_order=8, nbl=40, m=...,
    epsilon=..., delta=..., theta=..., phi=..., rho=..., b=..., dm=..., fs=False, dt=...)`.
    `qp` makes the model visco-acoustic (models.py:208-214, kernels.py:80-141); `f0` (kHz) is the
    peak frequency of its relaxation times, the value the reference passes with every call
    (`f0=options.f0`): set it here or per call with `set_f0`.  Elastic (`lam`, `mu`) models are
    outside this backend's path.
    """

    def __init__(self, origin, spacing, shape, space_order=8, nbl=40, dtype=np.float32,
                 m=None, epsilon=None, delta=None, theta=None, phi=None, rho=None, b=None,
                 qp=None, lam=None, mu=None, dm=None, fs=False, context=None, **kwargs):
        if lam is not None or mu is not None:
            raise NotImplementedError("elastic models are out of the B200 backend's scope "
                                      "(acoustic, visco-acoustic and TTI only)")
        if np.dtype(dtype) != np.float32:
            raise NotImplementedError("the B200 backend computes in float32 like the reference")
        if m is None:
            raise ValueError("squared slowness m is required")
        self.ctx = context or get_context()
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.nbl = int(nbl)
        self.fs = bool(fs)
        self.origin = tuple(np.float32(o) for o in origin)
        self.spacing = tuple(np.float32(s) for s in spacing)
        self.space_order = int(space_order)
        self._dt = kwargs.get("dt")
        self.is_tti = any(p is not None for p in (epsilon, delta, theta, phi))
        self.is_viscoacoustic = qp is not None
        self.f0 = float(kwargs.get("f0", 0.015))
        self.is_elastic = False
        keep = []
        d = L.JudiModelDesc()
        d.ndim = self.dim
        for k in range(self.dim):
            d.shape[k] = self.shape[k]
            d.spacing[k] = float(spacing[k])
            d.origin[k] = float(origin[k])
        d.nbl, d.space_order, d.fs = self.nbl, self.space_order, int(self.fs)
        d.dt = float(self._dt) if self._dt else 0.0
        on_dev = []
        for name, val in (("m", m), ("rho", rho), ("b", b), ("epsilon", epsilon),
                          ("delta", delta), ("theta", theta),
                          ("phi", phi if self.dim == 3 else None), ("dm", dm)):
            p, dev = _as_param(val, self.shape, keep)
            setattr(d, name, p)
            if p.kind == L.PARAM_ARRAY:
                on_dev.append(dev)
        if on_dev and any(on_dev) and not all(on_dev):
            raise ValueError("model arrays must be all host or all device resident")
        d.params_on_device = int(bool(on_dev) and all(on_dev))
        self.h = L.check_ptr(L.lib().judi_model_create(self.ctx.h, ctypes.byref(d)))
        self._has_dm = d.dm.kind == L.PARAM_ARRAY or (d.dm.kind == L.PARAM_SCALAR and d.dm.value != 0)
        if self.is_viscoacoustic:
            qv = qp if (is_device_tensor(qp) or np.ndim(qp) > 0) else np.full(self.shape, qp, dtype=np.float32)
            pq, dev = _as_param(qv, self.shape, keep)
            L.check(L.lib().judi_model_set_visco(self.h, ctypes.byref(pq), self.f0, int(dev)))

    def set_f0(self, f0):
        """Peak frequency (kHz) used by the following calls on a visco-acoustic model."""
        if f0 is None or not self.is_viscoacoustic or float(f0) == self.f0:
            return
        self.f0 = float(f0)
        L.check(L.lib().judi_model_set_visco(self.h, None, self.f0, 0))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                L.lib().judi_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- what the Julia driver reads -----------------------------------------------------
    @property
    def critical_dt(self):
        return np.float32(L.lib().judi_model_critical_dt(self.h))

    @property
    def padsizes(self):
        out = (ctypes.c_int * (2 * self.dim))()
        L.check(L.lib().judi_model_padsizes(self.h, out))
        return tuple((out[2 * k], out[2 * k + 1]) for k in range(self.dim))

    @property
    def shape_pml(self):
        out = (ctypes.c_int * self.dim)()
        L.check(L.lib().judi_model_padded_shape(self.h, out))
        return tuple(out[k] for k in range(self.dim))

    @property
    def dtype(self):
        return np.float32

    @property
    def dt(self):
        return self._dt

    # ---- perturbation ----------------------------------------------------------------------
    @property
    def has_dm(self):
        return self._has_dm

    def set_dm(self, dm):
        """Replace the squared-slowness perturbation used by Born modelling."""
        keep = []
        p, dev = _as_param(dm, self.shape, keep)
        L.check(L.lib().judi_model_set_dm(self.h, ctypes.byref(p), int(dev)))
        self._has_dm = p.kind == L.PARAM_ARRAY or (p.kind == L.PARAM_SCALAR and p.value != 0)

    # ---- test hooks ------------------------------------------------------------------------
    def get_field(self, name):
        """Padded field `name` ('m', 'damp', 'irho', ...) as an array of the padded shape."""
        out = np.zeros(self.shape_pml, dtype=np.float32, order="F")
        L.check(L.lib().judi_model_get_field(self.h, name.encode(), out.ctypes.data))
        return out

    def sparse_locate(self, coords):
        """Corner cells (ijk), weights and validity of Devito-style linear interpolation."""
        c = np.asfortranarray(np.asarray(coords, dtype=np.float32).reshape(-1, self.dim))
        npnt, nc = c.shape[0], 2 ** self.dim
        ijk = np.zeros((npnt, nc, 3), dtype=np.int32)
        w = np.zeros((npnt, nc), dtype=np.float32)
        valid = np.zeros((npnt, nc), dtype=np.int32)
        L.check(L.lib().judi_sparse_locate(self.h, npnt, c.ctypes.data, ijk.ctypes.data,
                                           w.ctypes.data, valid.ctypes.data))
        return ijk, w, valid
