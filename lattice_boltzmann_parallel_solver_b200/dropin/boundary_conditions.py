"""Flat-import shim: lets the reference's drivers (`from boundary_conditions import ...` with src/ on PYTHONPATH,
reference Makefile:3) pick up the B200 implementation. Put this directory first on PYTHONPATH."""
import lattice_boltzmann_parallel_solver_b200.boundary_conditions as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
