"""ctypes binding of libjudi_b200.so (the C-ABI declared in include/judi_b200.h).

The product path has no CPU fallback: if the shared library is missing it must be built
(`python judi.jl_b200/build.py`), and without a CUDA device every compute call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjudi_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)

PARAM_NONE, PARAM_SCALAR, PARAM_ARRAY = 0, 1, 2
IC_CODES = {"as": 0, "isic": 1, "fwi": 2}
DATA_ON_DEVICE, GRAD_ACCUMULATE, OUT_ON_DEVICE = 1, 2, 4
SNAP_PADDED, SNAP_PHYSICAL = 0, 1


class JudiParam(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("value", ctypes.c_float), ("data", ctypes.c_void_p)]


class JudiModelDesc(ctypes.Structure):
    _fields_ = [("ndim", ctypes.c_int), ("shape", ctypes.c_int * 3),
                ("spacing", ctypes.c_float * 3), ("origin", ctypes.c_float * 3),
                ("nbl", ctypes.c_int), ("space_order", ctypes.c_int), ("fs", ctypes.c_int),
                ("dt", ctypes.c_float),
                ("m", JudiParam), ("rho", JudiParam), ("b", JudiParam),
                ("epsilon", JudiParam), ("delta", JudiParam), ("theta", JudiParam),
                ("phi", JudiParam), ("dm", JudiParam),
                ("params_on_device", ctypes.c_int)]


class JudiJopts(ctypes.Structure):
    _fields_ = [("is_residual", ctypes.c_int), ("checkpointing", ctypes.c_int),
                ("n_checkpoints", ctypes.c_int), ("t_sub", ctypes.c_int), ("ic", ctypes.c_int),
                ("born_fwd", ctypes.c_int), ("nlind", ctypes.c_int), ("illum", ctypes.c_int),
                ("snapshot_region", ctypes.c_int), ("ws", ctypes.c_void_p),
                ("nfreq", ctypes.c_int), ("dft_sub", ctypes.c_int), ("freq_list", ctypes.c_void_p)]


class JudiShot(ctypes.Structure):
    _fields_ = [("nsrc", ctypes.c_int), ("src_coords", ctypes.c_void_p), ("wavelet", ctypes.c_void_p),
                ("nt", ctypes.c_int), ("nrec", ctypes.c_int), ("rec_coords", ctypes.c_void_p),
                ("recin", ctypes.c_void_p)]


MISFIT_FN = ctypes.CFUNCTYPE(ctypes.c_float, c_float_p, c_float_p, c_float_p, ctypes.c_int,
                             ctypes.c_int, ctypes.c_void_p)

# name -> (restype, argtypes); also the list of symbols the header declares (tests check it)
SIGNATURES = {
    "judi_abi_version": (ctypes.c_int, []),
    "judi_last_error": (ctypes.c_char_p, []),
    "judi_ctx_create": (ctypes.c_void_p, [ctypes.c_int]),
    "judi_ctx_destroy": (None, [ctypes.c_void_p]),
    "judi_ctx_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "judi_host_unregister": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_ctx_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "judi_ctx_launch_count": (ctypes.c_longlong, [ctypes.c_void_p]),
    "judi_ctx_stencil_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                              ctypes.POINTER(ctypes.c_longlong),
                                              ctypes.POINTER(ctypes.c_double)]),
    "judi_ctx_stencil_stats_kind": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_double),
                                                   ctypes.POINTER(ctypes.c_longlong)]),
    "judi_ctx_reset_stats": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_ctx_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "judi_ctx_trim": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "judi_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_comm_size": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_comm_rank": (ctypes.c_int, [ctypes.c_void_p]),
    "judi_comm_allreduce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "judi_fg_reduce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_fg_multishot_reduced": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(JudiShot),
                                                 ctypes.POINTER(JudiJopts), ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_model_create": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.POINTER(JudiModelDesc)]),
    "judi_model_destroy": (None, [ctypes.c_void_p]),
    "judi_model_critical_dt": (ctypes.c_float, [ctypes.c_void_p]),
    "judi_model_padsizes": (ctypes.c_int, [ctypes.c_void_p, c_int_p]),
    "judi_model_padded_shape": (ctypes.c_int, [ctypes.c_void_p, c_int_p]),
    "judi_model_set_dm": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(JudiParam), ctypes.c_int]),
    "judi_model_set_visco": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(JudiParam), ctypes.c_float,
                                            ctypes.c_int]),
    "judi_model_get_field": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]),
    "judi_sparse_locate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "judi_forward_rec": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int]),
    "judi_born_rec": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_J_adjoint": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.POINTER(JudiJopts), ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_fg_multishot": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(JudiShot),
                                         ctypes.POINTER(JudiJopts), ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_forward_no_rec": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_forward_rec_w": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int]),
    "judi_adjoint_w": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_born_rec_w": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_forward_wf_src": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "judi_forward_wf_src_norec": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_int]),
    "judi_remove_padding": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_int]),
    "judi_time_resample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_int]),
}

_lib = None


class JudiB200Error(RuntimeError):
    pass


def lib():
    """Load the shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JudiB200Error("%s is missing: build it with `python judi.jl_b200/build.py` "
                                "(the B200 backend has no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise JudiB200Error(lib().judi_last_error().decode())


def check_ptr(p):
    if not p:
        raise JudiB200Error(lib().judi_last_error().decode())
    return p
