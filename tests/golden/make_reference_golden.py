"""Golden vectors produced by EXECUTING the reference's own Python (src/pysource) in this container.

Devito is not installable here, so the Devito-generated stencil loops cannot run; but every part of
the hot path that src/pysource states in plain Python / numpy / sympy can, once `devito` is
replaced by an import stub that only provides symbols (no semantics of its own beyond sympy's):

  * models.Model._cfl_coeff / _thomsen_scale / _max_vp / critical_dt / padsizes   (models.py:269-272, 430-482)
  * FD_utils.sa_tti / R_mat / thomsen_mat: the rotated TTI flux (P, M) as a function of the
    staggered gradients, with `grad` returning symbols and `div` capturing its argument   (FD_utils.py:24-105)
  * sensitivity.crosscorr_time / isic_time / fwi_time / basic_src / isic_src / fwi_src: signs and
    factors of the imaging conditions and Born sources, on symbolic fields   (sensitivity.py:55-69, 106-122, 157-233)
  * sensitivity.Loss on numpy data (plain, is_residual, (background, linearised) tuple)   (sensitivity.py:236-276)
  * fields.wavefield_subsampled: number of saved snapshots   (fields.py:123-153)
  * kernels.SLS_2nd_order: the visco-acoustic update of memory variable and pressure, forward and
    adjoint, with `solve` done by sympy after u.dt2, u.dt, u.dt.T are replaced by their time
    stencils (the one-sided u.dt of SURVEY A.3 is THIS script's assumption, stated in the stub)   (kernels.py:80-141)

  * models.damp_op: the list of Eq / Inc over left / right SubDimensions that initialises the
    absorbing-layer field, interpreted by this script on a numpy array (SubDimension.left covers
    the first `thickness` indices of its parent, .right the last ones: Devito's documented
    meaning, [dv]) for both abc types, with and without free surface   (models.py:53-98)

  * geom_utils.geom_expr: which field and time level the source is injected into (forward /
    adjoint, single field / TTI pair), the injected expression src * dt^2 / (m * irho), and what
    the receivers interpolate, with PointSource / Receiver replaced by capturing symbols;
    sources.RickerSource.wavelet and sources.TimeAxis on numpy   (geom_utils.py:31-96, sources.py:11-176)

  * fields_exprs.otf_dft and sensitivity.grad_expr -> crosscorr_freq: the on-the-fly DFT of the
    forward wavefield and the frequency-domain imaging condition, evaluated on random time series
    of one grid point (ConditionalDimension(factor=k): tsave = time / k on the steps with
    time % k == 0, [dv]; the complex right-hand side stored into the real gradm keeps its real part, [dv])
    (fields_exprs.py:140-169, 195-212, sensitivity.py:27-52, 72-104)

Run from the repository root (needs /root/reference):  python tests/golden/make_reference_golden.py
Output: tests/golden/reference_golden.json (committed; the tests never read /root/reference).
"""
import json
import os
import sys
import types

import numpy as np
import sympy as sp

REF = os.environ.get("JUDI_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------
# import stub for devito: symbols only
class _Anything(object):
    """Placeholder for every devito name the imported modules mention but this script never calls."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything()


class _Module(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything


def _as_tuple(x, type=None, length=None):
    if x is None:
        return ()
    if isinstance(x, (tuple, list)):
        return tuple(x)
    return (x,)


class _Trig(object):
    def __init__(self, f):
        self.f = f
        self.__sympy_class__ = f

    def __call__(self, *a):
        return self.f(*a)


CAPTURED = {}


GRAD_OF = {}


def _grad(u, shift=None):
    assert shift == .5
    if not isinstance(u, Field):       # gradient of an expression: fresh symbols, argument recorded
        k = len(GRAD_OF)
        GRAD_OF[k] = u
        return sp.Matrix([sp.Symbol("gexpr%d_%s" % (k, d)) for d in "xyz"])
    return sp.Matrix([sp.Symbol("g%s_%s" % (u.name, d)) for d in "xyz"][:u.dim] if u.dim == 3 else
                     [sp.Symbol("g%s_x" % u.name), sp.Symbol("g%s_z" % u.name)])


def _div(v, shift=None):
    assert shift == -.5
    key = "div%d" % len(CAPTURED)
    CAPTURED[key] = v
    return sp.Symbol(key)


def _tensor(name=None, grid=None, components=None, symmetric=None, diagonal=None):
    return sp.Matrix(components)


class _Dt(sp.Symbol):
    @property
    def T(self):
        return sp.Symbol(self.name + "T")


FIELDS = []


class Field(sp.Symbol):
    """A symbolic (time) function: u, u.dt2 and its staggered gradient are independent symbols."""
    def __new__(cls, name, dim=3):
        obj = sp.Symbol.__new__(cls, name)
        obj.dim = dim
        FIELDS.append(obj)
        return obj

    @property
    def dt2(self):
        return sp.Symbol(self.name + "_dt2")

    @property
    def dt(self):
        return _Dt(self.name + "_dt")

    @property
    def forward(self):
        return sp.Symbol(self.name + "_f")

    @property
    def backward(self):
        return sp.Symbol(self.name + "_b")

    @property
    def laplace(self):
        return sp.Symbol(self.name + "_lap")

    @property
    def indices(self):
        t = types.SimpleNamespace(spacing=sp.Symbol("dt_save"))
        return (t,)


# ---------------------------------------------------------------------------------------------
# damp_op (models.py:53-98): symbols for the Devito objects + an interpreter of the Eq / Inc list
class _Dim(object):
    def __init__(self, name):
        self.name = name
        self.spacing = sp.Symbol("h_" + name)
        self.symbolic_min = sp.Symbol(name + "_m")
        self.symbolic_max = sp.Symbol(name + "_M")


class _SubDim(sp.Symbol):
    def __new__(cls, name, parent, thickness, side):
        obj = sp.Symbol.__new__(cls, name)
        obj.parent, obj.thickness, obj.side = parent, int(thickness), side
        return obj

    @classmethod
    def left(cls, name=None, parent=None, thickness=None):
        return cls(name, parent, thickness, "left")

    @classmethod
    def right(cls, name=None, parent=None, thickness=None):
        return cls(name, parent, thickness, "right")


class _DampFunction(object):
    def __init__(self, name=None, grid=None, space_order=0):
        self.dimensions = grid.dimensions

    def subs(self, mapping):
        (d, sub), = mapping.items()
        return ("on", d, sub)


class _Grid(object):
    def __init__(self, shape):
        self.dimensions = tuple(_Dim(n) for n in ("x", "y", "z")[:len(shape) - 1] + ("z",))[:len(shape)]
        if len(shape) == 2:
            self.dimensions = (_Dim("x"), _Dim("z"))


def run_damp_eqs(eqs, shape, spacing):
    """Apply the Eq / Inc list to a float64 array of `shape` (index ranges as Devito defines them
    for left / right SubDimensions of thickness t: [0, t) and [N - t, N))."""
    out = None
    for e in eqs:
        if e[0] == "Eq":
            out = np.full(shape, float(e[2]), dtype=np.float64)
            continue
        assert e[0] == "Inc"
        _, (_, d, sub), val = e
        ax = [dd.name for dd in DAMP_DIMS].index(d.name)
        N = shape[ax]
        idx = range(0, sub.thickness) if sub.side == "left" else range(N - sub.thickness, N)
        for i in idx:
            v = float(sp.sympify(val).xreplace({sub: sp.Integer(i), d.symbolic_min: sp.Integer(0),
                                                d.symbolic_max: sp.Integer(N - 1),
                                                d.spacing: sp.Float(spacing[ax])}))
            sl = [slice(None)] * len(shape)
            sl[ax] = i
            out[tuple(sl)] += v
    return out


DAMP_DIMS = ()


def _solve(expr, target):
    """devito.solve for these linear equations, after writing the time derivatives as stencils:
    u.dt2 = (u+ - 2u + u-)/dt^2, u.dt = (u+ - u)/dt (Devito's one-sided first derivative for
    time_order 2, SURVEY A.3 [dv]) and its transpose u.dt.T = (u- - u)/dt."""
    dt = sp.Symbol("dt")
    byname = {}
    for f in FIELDS:
        u = sp.Symbol(f.name)
        byname[f.name] = u
        byname[f.name + "_dt2"] = (f.forward - 2 * u + f.backward) / dt ** 2
        byname[f.name + "_dt"] = (f.forward - u) / dt
        byname[f.name + "_dtT"] = (f.backward - u) / dt
    expr = sp.sympify(expr)
    expr = expr.xreplace({s_: byname[s_.name] for s_ in expr.free_symbols if s_.name in byname})
    sol = sp.solve(expr, target)
    assert len(sol) == 1
    return sol[0]


def install_stub():
    dv = _Module("devito")
    dv.cos, dv.sin = _Trig(sp.cos), _Trig(sp.sin)
    dv.grad, dv.div = _grad, _div
    dv.TensorFunction = _tensor
    dv.Differentiable = sp.Expr
    dv.Function = type("Function", (), {})
    dv.Constant = type("Constant", (), {})
    dv.Eq = lambda *a, **k: ("Eq",) + a
    dv.solve = _solve
    dv.switchconfig = lambda **k: (lambda f: f)
    tools = _Module("devito.tools")
    tools.as_tuple = _as_tuple
    tools.memoized_func = lambda f: f
    sys.modules["devito"] = dv
    sys.modules["devito.tools"] = tools
    types_mod = _Module("devito.types")
    types_mod.SparseTimeFunction = type("SparseTimeFunction", (), {"__rkwargs__": ()})
    sys.modules["devito.types"] = types_mod
    for sub in ("devito.builtins", "devito.arch", "devito.arch.compiler",
                "devito.types.utils", "devito.symbolics", "devito.finite_differences"):
        sys.modules[sub] = _Module(sub)
    sys.path.insert(0, os.path.join(REF, "src", "pysource"))


def main():
    install_stub()
    import models           # noqa: the reference's module, unmodified
    import FD_utils
    import sensitivity
    import fields
    out = {"reference": "slimgroup/JUDI.jl src/pysource executed with a devito import stub",
           "cases": {}}
    rng = np.random.default_rng(20261019)

    # ---- 1. CFL / critical dt / padsizes ------------------------------------------------------
    class Probe(object):
        _cfl_coeff = models.Model._cfl_coeff
        _thomsen_scale = models.Model._thomsen_scale
        _max_vp = models.Model._max_vp
        critical_dt = models.Model.critical_dt
        padsizes = models.Model.padsizes
        dt = property(lambda self: self._dt)

    dts = []
    for ndim in (2, 3):
        for so in (4, 8, 12, 16):
            for tti in (False, True):
                for (vmax, h, user_dt) in ((1.5, 10., None), (4.7, 12.5, None), (5.5, 25., None),
                                           (3.0, 10., 0.75), (3.0, 10., 5.0)):
                    p = Probe()
                    p.is_elastic = False
                    p.is_tti = tti
                    p.space_order = so
                    p.grid = types.SimpleNamespace(dim=ndim)
                    p.dim = ndim
                    p.spacing = (h,) * ndim
                    p.dtype = np.float32
                    v = np.linspace(1.5, vmax, 7).astype(np.float32)
                    p.m = (1 / v ** 2).astype(np.float32)
                    eps = (0.25 * rng.random(7)).astype(np.float32)
                    p.epsilon = (1 + 2 * eps).astype(np.float32)     # models.py:218
                    p._dt = user_dt
                    p.nbl, p.fs = 40, False
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        dt = p.critical_dt
                    dts.append(dict(ndim=ndim, so=so, tti=tti, h=h, m=p.m.tolist(), eps=eps.tolist(),
                                    user_dt=user_dt, cfl=float(p._cfl_coeff), dt=float(dt)))
    out["cases"]["critical_dt"] = dts
    pads = []
    for ndim in (2, 3):
        for fs in (False, True):
            p = Probe()
            p.nbl, p.fs, p.dim = 17, fs, ndim
            pads.append(dict(ndim=ndim, fs=fs, nbl=17, padsizes=[list(t) for t in p.padsizes]))
    out["cases"]["padsizes"] = pads

    # ---- 2. rotated TTI flux --------------------------------------------------------------------
    tti_cases = []
    for ndim in (2, 3):
        CAPTURED.clear()
        model = types.SimpleNamespace(dim=ndim, grid=None, irho=sp.Symbol("b"), epsilon=sp.Symbol("eps"),
                                      delta=sp.Symbol("delta"), theta=sp.Symbol("theta"))
        if ndim == 3:
            model.phi = sp.Symbol("phi")
        u, v = Field("u", ndim), Field("v", ndim)
        FD_utils.sa_tti(u, v, model)
        PI, MI = CAPTURED["div0"], CAPTURED["div1"]
        syms = sorted(PI.free_symbols | MI.free_symbols, key=lambda s: s.name)
        fP = sp.lambdify(syms, PI, "numpy")
        fM = sp.lambdify(syms, MI, "numpy")
        for k in range(12):
            vals = {}
            for s in syms:
                if s.name == "b":
                    vals[s.name] = float(rng.uniform(0.4, 1.2))
                elif s.name == "eps":
                    vals[s.name] = 0.0
                elif s.name == "delta":
                    vals[s.name] = float(1 + 2 * rng.uniform(0.0, 0.15))
                elif s.name in ("theta", "phi"):
                    vals[s.name] = float(rng.uniform(-1.2, 1.2))
                else:
                    vals[s.name] = float(rng.standard_normal())
            vals["eps"] = vals["delta"] + float(2 * rng.uniform(0.0, 0.2))   # eps~ >= delta~
            args = [vals[s.name] for s in syms]
            tti_cases.append(dict(ndim=ndim, inputs=vals,
                                  P=[float(x) for x in np.array(fP(*args), dtype=np.float64).ravel()],
                                  M=[float(x) for x in np.array(fM(*args), dtype=np.float64).ravel()]))
    out["cases"]["tti_flux"] = tti_cases

    # ---- 3. imaging conditions and Born sources (signs and factors) -------------------------------
    ic_cases = []
    for nf in (1, 2):
        for name in ("as", "isic", "fwi"):
            model = types.SimpleNamespace(irho=sp.Symbol("irho"), m=sp.Symbol("m"), dm=sp.Symbol("dm"),
                                          is_tti=(nf == 2))
            us = tuple(Field("u%d" % f) for f in range(nf))
            vs = tuple(Field("v%d" % f) for f in range(nf))
            expr = sensitivity.ic_dict[name](us, vs, model)
            CAPTURED.clear()
            src = sensitivity.ls_dict[name](model, us)
            src = src if isinstance(src, tuple) else (src,)
            # laplacian(u, dm*irho) went through FD_utils.laplacian -> div(c * grad): captured
            lap = {k: list(vv) for k, vv in CAPTURED.items()}
            syms = sorted(set().union(expr.free_symbols, *[s.free_symbols for s in src],
                                      *[sp.Matrix(vv).free_symbols for vv in lap.values()]),
                          key=lambda s: s.name)
            vals = {s.name: float(rng.uniform(0.3, 1.7) * rng.choice([-1, 1])) for s in syms
                    if not s.name.startswith("div")}
            sub = {s: vals[s.name] for s in syms if s.name in vals}
            # a captured div(c*grad u) is handed back as the symbol divK: give it a value too
            for k in lap:
                vals[k] = float(rng.standard_normal())
                sub[sp.Symbol(k)] = vals[k]
            ic_cases.append(dict(nf=nf, ic=name, inputs=vals,
                                 grad_expr=float(expr.subs(sub)),
                                 lin_src=[float(s.subs(sub)) for s in src],
                                 div_args={k: [float(sp.sympify(c).subs(sub)) for c in vv]
                                           for k, vv in lap.items()}))
    out["cases"]["imaging"] = ic_cases

    # ---- 4. Loss -------------------------------------------------------------------------------------
    class Rec(object):
        def __init__(self, a):
            self.data = types.SimpleNamespace(_local=a.copy())

    loss = []
    nt, nr, dt = 23, 5, 1.375
    dsyn = rng.standard_normal((nt, nr)).astype(np.float32)
    dlin = rng.standard_normal((nt, nr)).astype(np.float32)
    dobs = rng.standard_normal((nt, nr)).astype(np.float32)
    for mode in ("plain", "is_residual", "nlind"):
        if mode == "nlind":
            f, r = sensitivity.Loss((Rec(dsyn), Rec(dlin)), dobs, dt)
        else:
            f, r = sensitivity.Loss(Rec(dsyn), dobs, dt, is_residual=(mode == "is_residual"))
        loss.append(dict(mode=mode, dt=dt, f=float(f), residual=np.asarray(r, dtype=np.float64).tolist()))
    # user misfit (sensitivity.py:255-263): f scaled by dt, residual as returned; with the
    # (background, linearised) tuple the callback sees dsyn[0] and dobs - dsyn[1]
    l1 = lambda a_, b_: (float(np.sum(np.abs(a_ - b_))), np.sign(a_ - b_).astype(np.float32))
    for mode in ("misfit", "misfit_nlind"):
        if mode == "misfit_nlind":
            f, r = sensitivity.Loss((Rec(dsyn), Rec(dlin)), dobs, dt, misfit=l1)
        else:
            f, r = sensitivity.Loss(Rec(dsyn), dobs, dt, misfit=l1)
        loss.append(dict(mode=mode, dt=dt, f=float(f), residual=np.asarray(r, dtype=np.float64).tolist()))
    out["cases"]["loss"] = dict(dsyn=dsyn.tolist(), dlin=dlin.tolist(), dobs=dobs.tolist(), results=loss)

    # ---- 5. number of saved snapshots -------------------------------------------------------------------
    saves = []

    class TF(object):
        def __init__(self, **kw):
            self.kw = kw

    fields.dvp = types.SimpleNamespace(TimeFunction=TF)
    fields.ConditionalDimension = lambda **k: ("cd", k.get("factor"))
    fields.compression_mode = lambda: None
    model = types.SimpleNamespace(grid=types.SimpleNamespace(time_dim="time"))
    for nt in (3, 10, 101, 260, 1042, 1705):
        for t_sub in (1, 2, 4, 7, 16):
            wf = fields.wavefield_subsampled(model, Field("u"), nt, t_sub)
            saves.append(dict(nt=nt, t_sub=t_sub, nsave=None if wf is None else int(wf[0].kw["save"])))
    out["cases"]["nsave"] = saves

    # ---- 6. visco-acoustic SLS update ---------------------------------------------------------------
    import kernels
    kernels.memory_field = lambda p: Field("r" + p.name)
    visco = []
    for fw in (True, False):
        CAPTURED.clear(); GRAD_OF.clear()
        model = types.SimpleNamespace(qp=sp.Symbol("qp"), irho=sp.Symbol("b"), damp=sp.Symbol("damp"),
                                      m=sp.Symbol("m"))
        pf = Field("p")
        (u_r, u_p), _ = kernels.SLS_2nd_order(model, pf, fw=fw, q=sp.Symbol("q"), f0=sp.Symbol("f0"))
        r_new, p_new = u_r[2], u_p[2]
        tgt_r = sp.Symbol("rp_f" if fw else "rp_b")
        assert u_r[1] == tgt_r and u_p[1] == sp.Symbol("p_f" if fw else "p_b")
        # laplacian(v, b) = div(b * grad(v)): each captured div is one application of L
        for k in range(6):
            vals = dict(dt=float(rng.uniform(0.5, 2.0)), f0=float(rng.uniform(0.008, 0.03)),
                        qp=float(rng.uniform(20., 200.)), b=float(rng.uniform(0.4, 1.2)),
                        m=float(rng.uniform(0.05, 0.5)), damp=float(rng.uniform(0.7, 1.0)),
                        p=float(rng.standard_normal()), rp=float(rng.standard_normal()),
                        q=float(rng.standard_normal()))
            vals["p_b" if fw else "p_f"] = float(rng.standard_normal())     # the other time level
            for name in CAPTURED:
                vals[name] = float(rng.standard_normal())
            if fw:
                vals["div1"] = vals["div0"]      # both are laplacian(p, b)
            def nsub(e, extra={}):       # substitute by NAME (Field / _Dt are Symbol subclasses)
                e = sp.sympify(e)
                table = dict(vals, **extra)
                return float(e.xreplace({s_: sp.Float(table[s_.name]) for s_ in e.free_symbols}))
            rn = nsub(r_new)
            pn = nsub(p_new, {tgt_r.name: rn})
            entry = dict(fw=fw, inputs=vals, r_new=rn, p_new=pn)
            if not fw:
                # what the two Laplacians of the adjoint pressure equation are applied to, as the
                # coefficients of p and of r- :  (1 + tt) p  and  (tt / (b t_s)) r-
                e0, e1 = GRAD_OF[0], GRAD_OF[1]
                entry["w_factor_p"] = nsub(e0) / vals["p"]
                entry["w_factor_r"] = nsub(e1, {tgt_r.name: rn}) / rn
            visco.append(entry)
    out["cases"]["visco"] = visco

    # ---- 7. acoustic and TTI time update ------------------------------------------------------------
    upd = []
    for kind in ("acoustic", "acoustic_rho", "tti"):
        for fw in (True, False):
            CAPTURED.clear()
            model = types.SimpleNamespace(irho=(sp.Symbol("b") if kind != "acoustic" else 0.8),
                                          m=sp.Symbol("m"), damp=sp.Symbol("damp"), fs=False,
                                          physical=None, dim=3, grid=None, epsilon=sp.Symbol("eps"),
                                          delta=sp.Symbol("delta"), theta=sp.Symbol("theta"),
                                          phi=sp.Symbol("phi"))
            if kind == "tti":
                u1, u2 = Field("u1"), Field("u2")
                eqs, _ = kernels.tti_kernel(model, u1, u2, fw=fw, q=(sp.Symbol("q1"), sp.Symbol("q2")))
                names = ["u1", "u2"]
            else:
                u1 = Field("u1")
                eqs, _ = kernels.acoustic_kernel(model, u1, fw=fw, q=sp.Symbol("q1"))
                names = ["u1"]
            for k in range(3):
                vals = dict(dt=float(rng.uniform(0.5, 2.0)), b=float(rng.uniform(0.4, 1.2)),
                            m=float(rng.uniform(0.05, 0.5)), damp=float(rng.uniform(0.0, 0.05)))
                for n_ in names:
                    vals[n_] = float(rng.standard_normal())
                    vals[n_ + ("_b" if fw else "_f")] = float(rng.standard_normal())
                    vals["q" + n_[1]] = float(rng.standard_normal())
                    vals[n_ + "_lap"] = float(rng.standard_normal())
                for name in CAPTURED:
                    vals[name] = float(rng.standard_normal())

                def nsub(e):
                    e = sp.sympify(e)
                    return float(e.xreplace({s_: sp.Float(vals[s_.name]) for s_ in e.free_symbols}))
                # which spatial term each equation carries: b * u.laplace (scalar density) or a
                # captured div(...) symbol (density field / TTI: div0 -> H0, div1 -> H1)
                upd.append(dict(kind=kind, fw=fw, inputs=vals, irho=(0.8 if kind == "acoustic" else vals["b"]),
                                new=[nsub(e[2]) for e in eqs]))
    out["cases"]["update"] = upd


    # ---- 8. absorbing-layer field ---------------------------------------------------------------------
    global DAMP_DIMS
    models.Grid = _Grid
    models.Function = _DampFunction
    models.SubDimension = _SubDim
    models.Abs, models.sin = sp.Abs, sp.sin
    models.Eq = lambda *a, **k: ("Eq",) + a
    models.Inc = lambda *a, **k: ("Inc",) + a
    models.Operator = lambda eqs, **k: eqs
    damps = []
    for ndim, nbl, n, h in ((2, 9, (7, 5), (10., 12.5)), (2, 40, (5, 4), (25., 25.)),
                            (3, 5, (3, 2, 3), (10., 20., 15.))):
        for fs in (False, True):
            for abc in ("damp", "mask"):
                pads = tuple((nbl, nbl) for _ in range(ndim - 1)) + (((0 if fs else nbl), nbl),)
                shape = tuple(n_ + l + r for n_, (l, r) in zip(n, pads))
                g = _Grid(shape)
                DAMP_DIMS = g.dimensions
                _orig = models.Grid
                models.Grid = lambda shp, _g=g: _g
                eqs = models.damp_op(ndim, pads, abc, fs)
                models.Grid = _orig
                arr = run_damp_eqs(eqs, shape, h)
                entry = dict(ndim=ndim, nbl=nbl, shape=list(n), spacing=list(h), fs=fs, abc=abc,
                             padsizes=[list(t) for t in pads])
                if arr.size <= 2100:
                    entry["damp"] = np.round(arr, 13).tolist()
                else:       # the reference's default layer width: axis lines through the middle + diagonal
                    mid = tuple(s_ // 2 for s_ in shape)
                    entry["lines"] = [np.round(arr[tuple(slice(None) if a == ax else mid[a] for a in range(ndim))], 13).tolist()
                                      for ax in range(ndim)]
                    entry["diag"] = [round(float(arr[(i,) * ndim]), 13) for i in range(min(shape))]
                damps.append(entry)
    out["cases"]["damp"] = damps

    # ---- 9. source injection / receiver interpolation expressions, Ricker wavelet, time axis ---------
    import geom_utils
    import sources

    class Sparse(sp.Symbol):
        def __new__(cls, name=None, grid=None, ntime=None, coordinates=None):
            obj = sp.Symbol.__new__(cls, name)
            obj.data = np.zeros((ntime, coordinates.shape[0]), dtype=np.float32)
            return obj

        def inject(self, field=None, expr=None):
            return [("inject", str(field), expr)]

        def interpolate(self, expr=None):
            return [("interpolate", str(expr))]

    geom_utils.PointSource = Sparse
    geom_utils.Receiver = Sparse
    geo = []
    for tti in (False, True):
        for fw in (True, False):
            model = types.SimpleNamespace(irho=sp.Symbol("b"), m=sp.Symbol("m"), is_elastic=False, is_tti=tti,
                                          fs=False, grid=types.SimpleNamespace(time_dim=types.SimpleNamespace(
                                              spacing=sp.Symbol("dt"))),
                                          __init_abox__=lambda *a, **k: None)
            u = (Field("u1"), Field("u2")) if tti else Field("u1")
            ex = geom_utils.geom_expr(model, u, src_coords=np.zeros((2, 3)), rec_coords=np.zeros((3, 3)),
                                      wavelet=np.ones((7, 2), dtype=np.float32), fw=fw)
            inj = [e for e in ex if e[0] == "inject"]
            itp = [e for e in ex if e[0] == "interpolate"]
            assert len(inj) == 1 and len(itp) == 1
            vals = dict(b=float(rng.uniform(0.4, 1.2)), m=float(rng.uniform(0.05, 0.5)),
                        dt=float(rng.uniform(0.5, 2.0)))
            src_sym = [s_ for s_ in inj[0][2].free_symbols if s_.name.startswith("src")][0]
            vals[src_sym.name] = float(rng.standard_normal())
            geo.append(dict(tti=tti, fw=fw, inject_field=inj[0][1], interpolate=itp[0][1], inputs=vals,
                            src_name=src_sym.name,
                            injected=float(inj[0][2].xreplace({s_: sp.Float(vals[s_.name])
                                                               for s_ in inj[0][2].free_symbols}))))
    out["cases"]["geometry"] = geo
    rick = []
    # the Julia helper the oracle restates spreads nt = floor(tmax / dt) + 1 samples over [0, tmax];
    # the two time axes coincide when tmax is a multiple of dt, where the wavelets are compared
    for (f0, dt_, tmax) in ((0.008, 2.0, 2000.), (0.010, 0.5, 1000.), (0.025, 4.0, 1200.),
                            (0.008, 1.921, 2000.), (0.010, 0.75, 1000.)):
        ax = sources.TimeAxis(start=0., step=dt_, stop=tmax)
        self_ = types.SimpleNamespace(f0=f0, a=None, t0=None)
        w = sources.RickerSource.wavelet(self_, ax.time_values)
        entry = dict(f0=f0, dt=dt_, tmax=tmax, num=int(ax.num), stop=float(ax.stop))
        if abs(ax.stop - tmax) < 1e-9:
            entry["wavelet"] = [float(x) for x in np.round(w[:160], 12)]
        rick.append(entry)
    out["cases"]["ricker"] = rick

    # ---- 10. on-the-fly DFT and the frequency-domain imaging condition ---------------------------------
    import fields_exprs

    class TimeDim(sp.Symbol):
        def __new__(cls, name):
            obj = sp.Symbol.__new__(cls, name, integer=True)
            obj.spacing = sp.Symbol("dt")
            obj.symbolic_max = sp.Symbol("time_M")
            return obj

    class Mode(sp.Symbol):
        """A DFT mode uf[freq_dim, x, ...]: indexing gives something with `.dimensions`."""
        def __getitem__(self, i):
            return types.SimpleNamespace(dimensions=(sp.Symbol("freq_dim"),))

    time = TimeDim("time")
    fields_exprs.ConditionalDimension = lambda name=None, parent=None, factor=None: sp.Symbol("tsave", integer=True)
    fields_exprs.Inc = lambda *a, **k: ("Inc",) + a
    fields_exprs.exp = sp.exp
    sensitivity.exp = sp.exp
    sensitivity.Eq = lambda *a, **k: ("Eq",) + a
    fsym = sp.Symbol("f")
    sensitivity.frequencies = lambda freq, fdim=None: (fsym, len(freq))
    dfts = []
    kappa = sp.Symbol("kappa")
    # inner_grad(a, b) = grad(a).grad(b) is bilinear: a stand-in kappa * a * b keeps every factor visible
    sensitivity.inner_grad = lambda a_, b_: kappa * a_ * b_
    tsave_sym = sp.Symbol("tsave", integer=True)
    for ic_name in ("as", "isic", "fwi"):
      for nfld in (1, 2):
        for factor in (None, 1, 2, 3):
            us = tuple(Field("u%d" % (i_ + 1)) for i_ in range(nfld))
            for u_ in us:
                u_.grid = types.SimpleNamespace(time_dim=time)
            modes = tuple(Mode("uf%d" % (i_ + 1)) for i_ in range(nfld))
            fields_exprs.fourier_modes = lambda u, freq, _m=modes: (_m, fsym)
            freqs = [0.004, 0.011]
            fwd = fields_exprs.otf_dft(us, freqs, sp.Symbol("dt"), factor=factor)
            assert len(fwd) == nfld and all(e[0] == "Inc" and e[1] == m_ for e, m_ in zip(fwd, modes))
            model = types.SimpleNamespace(grid=types.SimpleNamespace(time_dim=time), physical=None,
                                          m=sp.Symbol("m"), irho=sp.Symbol("b"))
            vs = tuple(Field("v%d" % (i_ + 1)) for i_ in range(nfld))
            gradm = sp.Symbol("gradm")
            (eq,) = sensitivity.grad_expr(gradm, modes, vs, model, w=None, freq=freqs, dft_sub=factor, ic=ic_name)
            assert eq[0] == "Eq" and eq[1] == gradm
            rhs = eq[2]
            names = sorted(s_.name for s_ in rhs.free_symbols)
            # evaluate both on random time series of one grid point
            nt, dtv = 14, 1.75
            mv, kv = float(rng.uniform(0.05, 0.5)), float(rng.uniform(0.2, 2.0))
            useries = rng.standard_normal((nt, nfld))
            vseries = rng.standard_normal((nt, nfld))
            k = factor or 1
            uf = np.zeros((len(freqs), nfld), dtype=np.complex128)
            for t in range(1, nt - 1):
                if t % k:
                    continue
                for qi, fv in enumerate(freqs):
                    for fi in range(nfld):
                        e = fwd[fi][2]
                        sub = {fsym: sp.Float(fv), sp.Symbol("dt"): sp.Float(dtv), us[fi]: sp.Float(useries[t, fi]),
                               time: sp.Integer(t), tsave_sym: sp.Integer(t // k)}
                        uf[qi, fi] += complex(e.xreplace(sub))
            # an equation that mentions the conditional dimension runs on its steps only [dv]
            gk = k if tsave_sym in rhs.free_symbols else 1
            g = 0.0
            for t in range(nt - 2, 0, -1):
                if t % gk:
                    continue
                for qi, fv in enumerate(freqs):
                    sub = {fsym: sp.Float(fv), sp.Symbol("dt"): sp.Float(dtv), time: sp.Integer(t),
                           tsave_sym: sp.Integer(t // gk), sp.Symbol("time_M"): sp.Integer(nt - 2),
                           gradm: sp.Float(0), sp.Symbol("m"): sp.Float(mv), kappa: sp.Float(kv)}
                    for fi in range(nfld):
                        sub[modes[fi]] = sp.sympify(complex(uf[qi, fi]))
                        sub[vs[fi]] = sp.Float(vseries[t, fi])
                    g += complex(rhs.xreplace(sub)).real
            dfts.append(dict(ic=ic_name, nfields=nfld, dft_sub=factor, freqs=freqs, nt=nt, dt=dtv, m=mv, kappa=kv,
                             forward_symbols=sorted(s_.name for s_ in fwd[0][2].free_symbols),
                             gradient_symbols=names, u=useries.tolist(), v=vseries.tolist(),
                             uf_re=uf.real.tolist(), uf_im=uf.imag.tolist(), gradm=g))
    out["cases"]["dft"] = dfts

    path = os.path.join(HERE, "reference_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
