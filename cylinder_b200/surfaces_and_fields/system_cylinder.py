from ..system_cylinder import Cylinder  # noqa: F401
