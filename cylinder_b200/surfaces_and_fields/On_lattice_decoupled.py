"""On_lattice_decoupled.Lattice (reference: surfaces_and_fields/On_lattice_decoupled.py)."""
from ..lattice import LatticeDecoupled as Lattice  # noqa: F401
