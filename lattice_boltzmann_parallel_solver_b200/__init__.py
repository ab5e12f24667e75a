"""B200-native D2Q9 fp64 lattice Boltzmann time step behind the function signatures of
SimonSchrodi/lattice_boltzmann_parallel_solver's `src/lattice_boltzmann_method.py`,
`src/boundary_conditions.py`, `src/boundary_utils.py` and `src/parallelization_utils.py`.

    from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method, boundary_utils, ...

or, to drive the reference's own `experiments.py` / `main.py` unchanged, put
`lattice_boltzmann_parallel_solver_b200/dropin` first on PYTHONPATH.

All array work runs in hand-written sm_100a CUDA kernels behind the C-ABI of include/lbm_b200.h
(`liblbm_b200.so`, built by `__graft_entry__.build()`); there is no CPU fallback.
"""
from . import _native                                  # noqa: F401
from . import lattice_boltzmann_method                 # noqa: F401
from . import parallelization_utils                    # noqa: F401
from . import boundary_conditions                      # noqa: F401
from . import boundary_utils                           # noqa: F401
from .engine import Lattice, LatticeArray              # noqa: F401

__version__ = '0.1.0'
