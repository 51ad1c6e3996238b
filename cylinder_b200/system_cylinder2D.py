"""Drop-in mirror of surfaces_and_fields/system_cylinder2D.py (`Cylinder2D`, with the parts of `Cylinder1D` it
inherits): complex field on the perturbed cylinder decomposed in Fourier modes c[beta, j].

The reference tabulates Toeplitz tensors A, B, D by 184 adaptive quadratures per amplitude proposal and contracts
them by einsum (system_cylinder2D.py:103-198).  Here the energy is one surface integral evaluated on the device
(cyl_fourier_energy_f64), so there are no temporary matrices: `save_temporary_matrices` only remembers
which amplitude `calc_field_energy` refers to, which is all it means in the reference (run.py:291).
"""
import numpy as np
import torch

from . import ops
from .system_cylinder import Cylinder


class Cylinder2D(Cylinder):
    def __init__(self, wavenumber, radius, alpha, C, u, n, kappa, gamma, intrinsic_curvature=0, num_field_coeffs=(0, 0),
                 device=None):
        super().__init__(wavenumber=wavenumber, radius=radius, kappa=kappa, gamma=gamma,
                         intrinsic_curvature=intrinsic_curvature, device=device)
        self.alpha, self.C, self.u, self.n = alpha, C, u, n
        self.num_field_coeffs_z, self.num_field_coeffs_theta = num_field_coeffs
        self.len_arrays_z = 2 * self.num_field_coeffs_z + 1
        self.len_arrays_theta = 2 * self.num_field_coeffs_theta + 1
        self._params = ops.make_params(self.device, wavenumber=wavenumber, radius=radius, gamma=gamma, kappa=kappa,
                                       intrinsic_curvature=intrinsic_curvature, alpha=alpha, u=u, C=C, n=n, temperature=0.0)
        self._amplitude = 0.0           # amplitude the "stored matrices" belong to
        self._tmp_amplitude = 0.0       # amplitude of the last calc_field_energy_amplitude_change

    def _coeffs(self, field_coeffs):
        c = np.ascontiguousarray(field_coeffs, dtype=complex).reshape(1, self.len_arrays_theta, self.len_arrays_z)
        return torch.from_numpy(c).to(self.device)

    def _energy(self, amplitude, field_coeffs):
        a = torch.tensor([float(amplitude)], dtype=torch.float64, device=self.device)
        return ops.fourier_energy(self._params, self.num_field_coeffs_z, self.num_field_coeffs_theta, a,
                                  self._coeffs(field_coeffs)).item()

    def calc_field_energy(self, field_coeffs):
        """Energy at the last ACCEPTED amplitude (system_cylinder2D.py:151-167)."""
        return self._energy(self._amplitude, field_coeffs)

    def calc_field_energy_amplitude_change(self, amplitude, field_coeffs):
        """Energy of the same field on a proposed shape (:176-198)."""
        self._tmp_amplitude = float(amplitude)
        return self._energy(amplitude, field_coeffs)

    def save_temporary_matrices(self):
        self._amplitude = self._tmp_amplitude

    def calc_surface_energy(self, amplitude):
        """Cylinder1D.calc_surface_energy (system_cylinder1D.py:189-194): effective_gamma * area integral + kappa * bending
        (no 2 pi, kappa not halved -- not the normalisation of Cylinder.calc_surface_energy)."""
        a = torch.tensor([float(amplitude)], dtype=torch.float64, device=self.device)
        return ops.fourier_surface_energy(self._params, a).item()

    # batched form: many chains / amplitudes in one launch
    def calc_field_energy_batch(self, amplitudes, field_coeffs):
        """amplitudes [B], field_coeffs [B, 2*nth+1, 2*nz+1] (numpy or torch) -> energies [B] (torch, device)."""
        amps = torch.as_tensor(amplitudes, dtype=torch.float64, device=self.device).reshape(-1)
        c = torch.as_tensor(np.array(field_coeffs, dtype=complex), device=self.device).reshape(amps.numel(), self.len_arrays_theta, self.len_arrays_z)
        return ops.fourier_energy(self._params.expand(amps.numel(), -1).contiguous(), self.num_field_coeffs_z,
                                  self.num_field_coeffs_theta, amps, c.contiguous())
