"""montecarlosolvers_b200 -- B200-native annealing sweeps behind the MonteCarloSolvers call surface.

Drop-in modules (same function names and positional arguments as the reference's Cython modules,
SURVEY.md 8b):  `qmc`, `sa`, `svmc`, `tools`.  Everything computes on the GPU through the C ABI of
libmcs_b200.so (include/mcs_b200.h); there is no CPU fallback.
"""
from . import _lib
from ._lib import Instance, State, McsError, empty_pinned, reseed, device_count  # noqa: F401
from . import qmc, sa, svmc, tools, parallel  # noqa: F401

__all__ = ["qmc", "sa", "svmc", "tools", "parallel", "Instance", "State", "McsError", "empty_pinned", "reseed",
           "device_count"]
__version__ = "0.1.0"
