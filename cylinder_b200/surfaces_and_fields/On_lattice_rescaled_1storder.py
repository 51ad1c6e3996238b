"""On_lattice_rescaled_1storder.Lattice (reference: surfaces_and_fields/On_lattice_rescaled_1storder.py)."""
from ..lattice import LatticeRescaled1stOrder as Lattice  # noqa: F401
