"""judi.jl_b200 -- B200-native (sm_100a) drop-in for the wave-equation hot path of
slimgroup/JUDI.jl: the part of `src/pysource` that Devito-generated code executes.

  csrc/          hand-written CUDA kernels + the C-ABI (include/judi_b200.h)
  models.py      mirror of pysource.models.Model (handle to the HBM-resident padded model)
  interface.py   mirror of pysource.interface (forward_rec, born_rec, J_adjoint, ...)
  objectives.py  mirror of the Julia L2/L5 drivers (multi_src_fg!, fwi_objective,
                 lsrtm_objective): shot partition over GPUs + one all-reduce
  time_modeling.py  mirror of the per-shot driver (_time_modeling / devito_interface):
                 limit_m, device time_resample, device remove_padding

The directory name contains a dot, so the package is imported through the top-level shim
`judi_jl_b200` (see /judi_jl_b200.py).
"""
from ._lib import JudiB200Error, LIB_PATH  # noqa: F401
from .models import Model, Context, get_context, host_register, host_unregister  # noqa: F401
from . import interface  # noqa: F401
from .interface import (forward_rec, forward_no_rec, born_rec, J_adjoint, forward_rec_w,  # noqa: F401
                        adjoint_w, born_rec_w, forward_wf_src, forward_wf_src_norec)
from . import objectives  # noqa: F401
from . import time_modeling  # noqa: F401
from .time_modeling import Geometry, Options, time_resample  # noqa: F401
