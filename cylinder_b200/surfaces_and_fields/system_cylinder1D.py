from ..system_cylinder1D import Cylinder1D  # noqa: F401
