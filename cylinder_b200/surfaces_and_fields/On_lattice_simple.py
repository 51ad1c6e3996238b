"""On_lattice_simple.Lattice (reference: surfaces_and_fields/On_lattice_simple.py)."""
from ..lattice import Lattice  # noqa: F401
