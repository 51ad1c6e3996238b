from ..system_cylinder2D import Cylinder2D  # noqa: F401
