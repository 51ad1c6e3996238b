"""On_lattice_rescaled_0thorder.Lattice (reference: surfaces_and_fields/On_lattice_rescaled_0thorder.py)."""
from ..lattice import LatticeRescaled0thOrder as Lattice  # noqa: F401
