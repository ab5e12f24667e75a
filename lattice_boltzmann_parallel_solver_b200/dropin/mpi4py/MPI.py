from lattice_boltzmann_parallel_solver_b200.dist import CartComm, WorldComm, comm_world

Intracomm = WorldComm
Cartcomm = CartComm


def __getattr__(name):
    if name == 'COMM_WORLD':
        return comm_world()
    raise AttributeError(name)
