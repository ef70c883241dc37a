"""smoke-simulation_b200 -- B200-native (sm_100a) smoke simulation step.

The product is the shared library ``libsmoke_b200.so`` in this directory: hand-written CUDA kernels
behind the C ABI of ``include/smoke_b200.h`` plus the reference's own C++ entry points
(``host/smokeSimulation.cuh``).  This Python package is only the host-side mirror used by tests and
bench.py: a ctypes binding (``binding.py``), the multi-GPU slab plumbing (``slab.py``) and the synthetic scenes of
SURVEY.md section 8(d) (``scenes.py``).

The directory name contains a hyphen; import it as ``smoke_simulation_b200`` (the module of that name at
the repository root points its ``__path__`` here).
"""
from . import scenes, slab  # noqa: F401
from .binding import (  # noqa: F401
    LIB_PATH, SmokeSim, SmokeError, load_library, build_library, declared_symbols, selfcheck_omega,
    SMOKE, U, V, W, MASK, NOW, PAST, BUF0, BUF1,
)
