"""cylinder_b200 -- B200-native Metropolis engine for the hot path of jklebes/cylinder.

Reference-facing classes (same names / arguments as the reference):
    Cylinder, Cylinder2D, Lattice (On_lattice_simple), LatticeRescaled0thOrder (On_lattice_rescaled_0thorder),
    LatticeRescaled1stOrder (On_lattice_rescaled_1storder), LatticeDecoupled (On_lattice_decoupled),
    MetropolisEngine, run.single_run
Batched / multi-GPU form of the reference's job farm: replicas.ReplicaBatch, run.grid_run.
Everything numerical runs in libcylinder_b200.so (CUDA, sm_100a) through the C-ABI of include/cylinder_b200.h;
importing this package does not need a GPU, using it does.
"""
__all__ = ["Cylinder", "Cylinder1D", "Cylinder2D", "Lattice", "LatticeRescaled0thOrder", "LatticeRescaled1stOrder", "LatticeDecoupled",
           "MetropolisEngine", "ReplicaBatch"]


def __getattr__(name):
    if name == "Cylinder":
        from .system_cylinder import Cylinder
        return Cylinder
    if name == "Cylinder2D":
        from .system_cylinder2D import Cylinder2D
        return Cylinder2D
    if name == "Cylinder1D":
        from .system_cylinder1D import Cylinder1D
        return Cylinder1D
    if name in ("Lattice", "LatticeRescaled0thOrder", "LatticeRescaled1stOrder", "LatticeDecoupled"):
        from . import lattice
        return getattr(lattice, name)
    if name == "MetropolisEngine":
        from .metropolisengine import MetropolisEngine
        return MetropolisEngine
    if name == "ReplicaBatch":
        from .replicas import ReplicaBatch
        return ReplicaBatch
    raise AttributeError(name)
