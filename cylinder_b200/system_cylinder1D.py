"""Drop-in mirror of surfaces_and_fields/system_cylinder1D.py (`Cylinder1D`): the field depends on z only, coefficients
[c_-n, ..., c_n].  It is the 2-D model without theta modes (system_cylinder2D.py with num_field_coeffs = (n, 0): the same A, D tables,
B[i, j] = B[i, j, beta = 0]), so everything is delegated to Cylinder2D and its device kernels; the attribute names are the 1-D class's
(`num_field_coeffs`, `len_arrays`, `B_integrals`, `tmp_B_integrals`; system_cylinder1D.py:8-27)."""
import numpy as np

from .system_cylinder2D import Cylinder2D


class Cylinder1D(Cylinder2D):
    def __init__(self, wavenumber, radius, alpha, C, u, n, kappa, gamma, intrinsic_curvature=0, num_field_coeffs=0, device=None):
        super().__init__(wavenumber=wavenumber, radius=radius, alpha=alpha, C=C, u=u, n=n, kappa=kappa, gamma=gamma,
                         intrinsic_curvature=intrinsic_curvature, num_field_coeffs=(num_field_coeffs, 0), device=device)
        self.num_field_coeffs = num_field_coeffs
        self.len_arrays = 2 * num_field_coeffs + 1

    def _coeffs(self, field_coeffs):
        return super()._coeffs(np.asarray(field_coeffs, dtype=complex).reshape(1, self.len_arrays))

    tmp_B_integrals = property(lambda self: self.tmp_B_matrix[:, :, 0, 0])
    B_integrals = property(lambda self: self.B_matrix[:, :, 0, 0])
