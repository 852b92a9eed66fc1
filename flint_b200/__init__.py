"""flint_b200 -- B200-native batched adaptive-ERK ensemble integrator behind FLINT's object API.

Package contents (only what the hot path needs):
  csrc/        CUDA kernels (sm_100a) + the C ABI (include/flint_b200.h)
  api.py       host-side mirror of the reference's DiffEqSys / ERK_class interface (ctypes)
  problems.py  the reference test programs' workloads (BASELINE.json configs)
  build.py     nvcc build of libflint_b200.so
"""
from .api import *  # noqa: F401,F403
from . import api, problems  # noqa: F401
