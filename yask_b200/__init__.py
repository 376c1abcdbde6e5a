"""yask_b200: B200-native execution engine for the hot path of intel/yask
(yk_solution::run_solution for the iso3dfd / awp / ssg stencils).

The product is libyask_b200.so (CUDA for sm_100a behind the C ABI in include/yask_b200.h);
this package only holds the ctypes binding (capi), the synthetic-input generator (synth) and the
multi-GPU launcher glue.  There is no CPU compute path.
"""
from . import capi, synth  # noqa: F401
from .capi import Solution, YaskError  # noqa: F401
