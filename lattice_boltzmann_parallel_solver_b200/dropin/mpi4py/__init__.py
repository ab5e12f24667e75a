"""Stand-in for `mpi4py` when it is not installed: the reference's drivers do `from mpi4py import MPI` and use
MPI.COMM_WORLD.Get_size/Get_rank/Create_cart and the Cartesian communicator's Get_coords/Shift/Sendrecv
(src/experiments.py:606-614, src/parallelization_utils.py:18-49). Here those names are backed by
torch.distributed (one process per GPU under torchrun) — see lattice_boltzmann_parallel_solver_b200/dist.py."""
from . import MPI  # noqa: F401
