"""CPU tests: host logic of the oracle against the reference's formulas, the C-ABI library
(loads without a GPU, exports every symbol the header declares, fails loudly instead of falling
back), and the Julia-side helpers restated in the package."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, model_args
from oracle import oracle as O


def test_cfl_coefficients_match_survey_values():
    """models.py:440-456: cfl(2-D)=0.5177 (so 4/8), 0.4721 (so16); cfl(3-D)=0.4227, 0.3855."""
    assert abs(O.cfl_coeff(8, 2) - 0.5177) < 1e-4
    assert abs(O.cfl_coeff(4, 2) - 0.5177) < 1e-4
    assert abs(O.cfl_coeff(16, 2) - 0.4721) < 1e-4
    assert abs(O.cfl_coeff(8, 3) - 0.4227) < 1e-4
    assert abs(O.cfl_coeff(16, 3) - 0.3855) < 1e-4


def test_critical_dt_rounding_and_user_dt():
    """models.py:466-482: '%.3e' rounding; a smaller user dt wins, a larger one is ignored."""
    n = (401, 401, 201)
    m = np.full((5, 5, 5), 1 / 5.5 ** 2, dtype=np.float32)
    mod = O.Model((0, 0, 0), (25., 25., 25.), (5, 5, 5), space_order=8, nbl=4, m=m)
    assert mod.critical_dt == np.float32(1.921)      # C3 dt
    assert O.Model((0, 0, 0), (25., 25., 25.), (5, 5, 5), nbl=4, m=m, dt=1.5).critical_dt == np.float32(1.5)
    assert O.Model((0, 0, 0), (25., 25., 25.), (5, 5, 5), nbl=4, m=m, dt=3.0).critical_dt == np.float32(1.921)
    del n


def test_damping_profile_properties():
    """models.py:54-96: zero inside the domain and in the 3 innermost layer cells, separable sum,
    symmetric, increasing outwards, coefficient 1.5 ln(1000)/(nbl-3) / h."""
    mod = O.Model((0, 0), (10., 20.), (31, 21), nbl=15, m=np.ones((31, 21), np.float32))
    d = mod.damp
    assert d.shape == (61, 51)
    assert not d[12:-12, 12:-12].any()
    assert d[11, 25] > 0 and d[12, 25] == 0
    assert np.allclose(d[:, 25], d[::-1, 25])
    px, pz = d[:, 25], d[30, :]
    assert np.allclose(d, px[:, None] + pz[None, :], rtol=1e-6)
    assert np.all(np.diff(px[:12]) < 0)
    nb = 12
    pos = (nb + 1) / nb
    expect = 1.5 * np.log(1000.) / nb * (pos - np.sin(2 * np.pi * pos) / (2 * np.pi)) / 10.
    assert abs(px[0] - expect) < 1e-6 * expect
    assert abs(pz[0] / px[0] - 0.5) < 1e-6          # divided by the spacing of its own axis


def test_padding_is_edge_replication_and_its_adjoint():
    """auxiliaryFunctions.jl:52-89: remove_padding(true_adjoint) is the adjoint of pad_array."""
    rng = np.random.default_rng(0)
    nb = ((3, 4), (2, 5))
    x = rng.standard_normal((6, 7))
    y = rng.standard_normal((13, 14))
    px = O.pad_array(x, nb)
    assert np.array_equal(px[:3, 2:-5], np.broadcast_to(x[0], (3, 7)))
    a = np.vdot(px, y)
    b = np.vdot(x, O.remove_padding(y, nb, true_adjoint=True))
    assert abs(a - b) < 1e-12 * abs(a)
    assert np.array_equal(O.remove_padding(y, nb), y[3:-4, 2:-5])
    import judi_jl_b200 as J
    assert np.array_equal(J.objectives.remove_padding(y, nb, True), O.remove_padding(y, nb, True))


def test_model_parameters_edge_padded_and_tti_scaling():
    args = model_args("tti", 2, 8)
    mod = O.Model(**args)
    assert mod.m.shape == mod.shape_pml
    assert np.array_equal(mod.m[:20, 30], np.full(20, mod.m[20, 30]))
    eps, _ = mod.tti["epsilon"]
    assert np.allclose(eps[20:-20, 20:-20], 1 + 2 * args["epsilon"])   # models.py:218
    assert mod.padsizes == ((20, 20), (20, 20))
    assert mod.origin_pml == (np.float32(-200.), np.float32(-200.))


def test_ricker_wavelet_matches_julia_formula():
    q = O.ricker_wavelet(1000., 2., 0.01)
    assert q.shape == (501, 1)
    t = np.arange(501) * 2.0
    r = np.pi * 0.01 * (t - 100.0)
    assert np.allclose(q[:, 0], (1 - 2 * r ** 2) * np.exp(-r ** 2), atol=1e-5)


def test_sparse_weights_partition_unity_and_skip_outside():
    args = model_args("iso", 3, 8)
    mod = O.Model(**args)
    pts = np.array([[13.7, 25.2, 31.9], [-150., 0., 0.], [0., 0., 0.], [400., 360., 320.]],
                   dtype=np.float32)
    ijk, w, valid = O.sparse_locate(mod, pts)
    assert np.allclose(w.sum(axis=1), 1.0, atol=1e-6)
    assert valid[0].all() and not valid[1].any()           # second point is outside the padded grid
    assert tuple(ijk[0, 0]) == (11, 12, 13)                # floor((13.7+100)/10) etc.
    assert w[2, 0] == 1.0 and not w[2, 1:].any()           # on a node: all weight on one corner


# ---- C-ABI ----------------------------------------------------------------------------------
def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "judi_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(judi_[a-zA-Z0-9_]+)\s*\(", txt)) - {"judi_misfit_fn"})


def test_library_loads_and_exports_every_declared_symbol():
    import judi_jl_b200 as J
    from judi_jl_b200 import _lib
    assert os.path.exists(J.LIB_PATH), "run python judi.jl_b200/build.py"
    L = ctypes.CDLL(J.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), "symbol %s declared in include/judi_b200.h is not exported" % name
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes binding covers the header
    assert _lib.lib().judi_abi_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly (no oracle / CPU path behind it)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import judi_jl_b200 as J
    with pytest.raises(J.JudiB200Error, match="no CPU fallback"):
        J.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "judi.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def test_out_of_scope_entry_points_fail_loudly():
    import judi_jl_b200 as J
    with pytest.raises(NotImplementedError):
        J.interface.wri_func()
    # all ten entry points of src/pysource/interface.py exist under the same names
    for name in ("forward_rec", "forward_rec_w", "forward_no_rec", "forward_wf_src",
                 "forward_wf_src_norec", "adjoint_w", "born_rec", "born_rec_w", "J_adjoint",
                 "wri_func"):
        assert callable(getattr(J.interface, name))


# ---- the Julia `ccall` twin (julia/B200Backend.jl) against the header ---------------------------
def _header_prototypes():
    """name -> (return class, [argument classes]) from include/judi_b200.h; a class is one of
    'int', 'float', 'size_t', 'longlong', 'ptr', 'void'."""
    txt = open(os.path.join(ROOT, "include", "judi_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)

    def cls(decl):
        decl = decl.strip()
        if "*" in decl or "[" in decl or "judi_misfit_fn" in decl:
            return "ptr"
        base = re.sub(r"\b(const|unsigned|struct)\b", "", decl).split()
        t = base[0] if base else "void"
        return {"int": "int", "float": "float", "size_t": "size_t", "long": "longlong",
                "void": "void", "double": "double"}[t]
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(judi_[a-z0-9_A-Z]+)\s*\(([^;{}]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if name == "judi_misfit_fn" or "typedef" in ret:
            continue
        args = [a for a in (x.strip() for x in args.replace("\n", " ").split(",")) if a and a != "void"]
        protos[name] = (cls(ret + (" *" if ret.strip().endswith("*") else "")), [cls(a) for a in args])
    return protos


_JL_CLASS = {"Cint": "int", "Cfloat": "float", "Csize_t": "size_t", "Clonglong": "longlong",
             "Cvoid": "void", "Cstring": "ptr", "Cdouble": "double"}


def _julia_ccalls():
    """[(name, return class, [argument classes], number of values passed)] of every ccall in
    julia/B200Backend.jl."""
    src = open(os.path.join(ROOT, "julia", "B200Backend.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    out = []
    for m in re.finditer(r"ccall\(\(:(judi_[a-z0-9_A-Z]+), LIB\)\s*,", src):
        # balanced scan of the rest of the call
        i, depth, parts, cur = m.end(), 1, [], ""
        while depth > 0:
            c = src[i]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if c == "," and depth == 1:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += c
            i += 1
        parts.append(cur.strip())
        ret, argt, vals = parts[0], parts[1], parts[2:]
        assert argt.startswith("(") and argt.endswith(")"), (m.group(1), argt)
        inner = argt[1:-1].strip()
        types, d, cur = [], 0, ""
        for c in inner:
            if c == "{":
                d += 1
            elif c == "}":
                d -= 1
            if c == "," and d == 0:
                types.append(cur.strip())
                cur = ""
            else:
                cur += c
        if cur.strip():
            types.append(cur.strip())

        def jcls(t):
            return "ptr" if t.startswith(("Ptr{", "Ref{")) else _JL_CLASS[t]
        out.append((m.group(1), jcls(ret), [jcls(t) for t in types], len([v for v in vals if v])))
    return out


def test_julia_ccall_twin_matches_the_header():
    """Every ccall of julia/B200Backend.jl (the binding INTEGRATION.md hands to a JUDI maintainer;
    no Julia in this image, so it cannot be executed here) names an exported symbol and passes the
    number and kind (int / float / size_t / pointer) of arguments the header declares."""
    protos = _header_prototypes()
    assert set(protos) == set(_header_symbols())
    calls = _julia_ccalls()
    assert len(calls) >= 20
    for name, ret, types, nvals in calls:
        assert name in protos, "%s is not declared in include/judi_b200.h" % name
        pret, pargs = protos[name]
        assert ret == pret, "%s: Julia return type class %s, header %s" % (name, ret, pret)
        assert types == pargs, "%s: Julia argument types %s, header %s" % (name, types, pargs)
        assert nvals == len(types), "%s: %d values for %d declared arguments" % (name, nvals, len(types))
    # the entry points of the hot path all have a twin
    named = {c[0] for c in calls}
    for must in ("judi_model_create", "judi_forward_rec", "judi_born_rec", "judi_J_adjoint",
                 "judi_fg_multishot", "judi_fg_multishot_reduced", "judi_comm_init",
                 "judi_forward_rec_w", "judi_adjoint_w", "judi_born_rec_w", "judi_forward_wf_src",
                 "judi_forward_wf_src_norec", "judi_remove_padding", "judi_time_resample"):
        assert must in named, must


def test_ctypes_binding_matches_the_header_prototypes():
    """Argument count and int / float / pointer class of every ctypes signature against the header."""
    from judi_jl_b200 import _lib
    protos = _header_prototypes()

    def ccls(t):
        if t is None:
            return "void"
        if t in (ctypes.c_int,):
            return "int"
        if t is ctypes.c_float:
            return "float"
        if t is ctypes.c_size_t:
            return "size_t"
        if t is ctypes.c_longlong:
            return "longlong"
        return "ptr"
    for name, (res, args) in _lib.SIGNATURES.items():
        pret, pargs = protos[name]
        assert ccls(res) == pret, (name, res, pret)
        assert [ccls(a) for a in args] == pargs, (name, [ccls(a) for a in args], pargs)


def _header_structs():
    """name -> [(field, class)] for every `typedef struct { ... } judi_xxx;` of the header; class is
    'int', 'float', 'ptr', an embedded struct name, or ('int' | 'float', n) for a fixed array."""
    txt = open(os.path.join(ROOT, "include", "judi_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef struct\s*\{(.*?)\}\s*(judi_[a-z_]+)\s*;", txt, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            base, names = decl.rsplit(" ", 1)[0], None
            # "judi_param epsilon, delta, theta, phi" / "const float *data" / "int shape[3]"
            mm = re.match(r"(const\s+)?([a-z_]+)\s+(.*)", decl)
            typ, names = mm.group(2), [n.strip() for n in mm.group(3).split(",")]
            for n in names:
                if n.startswith("*"):
                    fields.append((n.lstrip("*"), "ptr"))
                elif "[" in n:
                    fields.append((n.split("[")[0], (typ, int(n.split("[")[1].rstrip("]")))))
                else:
                    fields.append((n, typ))
        out[m.group(2)] = fields
    return out


def test_struct_layouts_agree_between_header_ctypes_and_julia():
    """A field out of order in a binding is silent memory corruption: compare field names, order and
    type class of judi_param / judi_model_desc / judi_jopts / judi_shot across the header, the ctypes
    Structures the GPU tests use and the Julia structs of julia/B200Backend.jl."""
    from judi_jl_b200 import _lib
    H = _header_structs()
    assert set(H) >= {"judi_param", "judi_model_desc", "judi_jopts", "judi_shot"}
    cmap = {"judi_param": _lib.JudiParam, "judi_model_desc": _lib.JudiModelDesc,
            "judi_jopts": _lib.JudiJopts, "judi_shot": _lib.JudiShot}

    def ccls(t):
        if t is ctypes.c_int:
            return "int"
        if t is ctypes.c_float:
            return "float"
        if t is ctypes.c_void_p or hasattr(t, "contents"):
            return "ptr"
        if issubclass(t, ctypes.Array):
            return (ccls(t._type_), t._length_)
        if issubclass(t, ctypes.Structure):
            return {v: k for k, v in cmap.items()}[t]
        raise AssertionError(t)
    for name, S in cmap.items():
        assert [(f, ccls(t)) for f, t in S._fields_] == H[name], name

    # Julia: `struct JudiX  # judi_x` ... `end`, fields `name::Type` separated by ';' or newlines
    src = open(os.path.join(ROOT, "julia", "B200Backend.jl")).read()
    jl = {}
    for m in re.finditer(r"struct\s+(Judi[A-Za-z]+)\s*#\s*(judi_[a-z_]+)\s*\n(.*?)\nend", src, flags=re.S):
        body = re.sub(r"#[^\n]*", "", m.group(3))
        fields = []
        for f in re.split(r"[;\n]", body):
            f = f.strip()
            if not f:
                continue
            n, t = f.split("::")
            if t.startswith("Ptr{"):
                c = "ptr"
            elif t.startswith("NTuple{"):
                k, et = re.match(r"NTuple\{(\d+),\s*(\w+)\}", t).groups()
                c = ({"Cint": "int", "Cfloat": "float"}[et], int(k))
            elif t in ("Cint", "Cfloat"):
                c = {"Cint": "int", "Cfloat": "float"}[t]
            else:
                c = {"JudiParam": "judi_param"}[t]
            fields.append((n.strip(), c))
        jl[m.group(2)] = fields
    assert set(jl) >= {"judi_param", "judi_model_desc", "judi_jopts"}, sorted(jl)
    for name, fields in jl.items():
        assert fields == H[name], (name, fields, H[name])
