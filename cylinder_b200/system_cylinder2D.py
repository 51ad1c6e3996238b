"""Drop-in mirror of surfaces_and_fields/system_cylinder2D.py (`Cylinder2D`, with the parts of `Cylinder1D` it
inherits): complex field on the perturbed cylinder decomposed in Fourier modes c[beta, j].

The reference tabulates Toeplitz tensors A, B, D by 184 adaptive quadratures per amplitude proposal and contracts
them by einsum (system_cylinder2D.py:103-198).  Here the energy is one surface integral evaluated on the device
(cyl_fourier_energy_f64): the energies never touch matrices, and `save_temporary_matrices` only has to remember
which amplitude `calc_field_energy` refers to (run.py:291).  The reference's tensor attributes (`tmp_A_integrals`, `tmp_A_matrix`,
`tmp_D_matrix`, `tmp_B_matrix`, `A_matrix`, `D_matrix`, `B_matrix`, `tmp_A_integrals_0`; tests/test_matrixA.py, tests/test_1Dvs2D.py
of the reference read them) are still there for callers that want them: filled on demand by cyl_fourier_tensor_tables_f64 (one warp per
element, 256-node periodic trapezoid instead of two scipy quad calls) with the reference's own index layout.
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
        self._tensors = {}              # amplitude -> (A integrals, A matrix, D matrix, B matrix), filled when somebody reads them
        self._A0_override = None        # set by evaluate_A_integral_0 (the reference overwrites tmp_A_integrals_0 there)
        self.quartic_selection_rule = self.generate_quartic_combinations(self.len_arrays_theta)

    # ---- the reference's tensor tables, on demand (system_cylinder2D.py:24-147) ----------------------------------------------
    def generate_quartic_combinations(self, len_arrays):
        """(:45-56) theta-index combinations with index1 + index2 == index3 + index4."""
        r = range(len_arrays)
        return [(a, b, c, d) for a in r for b in r for c in r for d in r if a + b == c + d]

    def _tables(self, amplitude):
        key = float(amplitude)
        if key not in self._tensors:
            nz, nth, lz, lth = self.num_field_coeffs_z, self.num_field_coeffs_theta, self.len_arrays_z, self.len_arrays_theta
            a = torch.tensor([key], dtype=torch.float64, device=self.device)
            Aint, Bd = ops.fourier_tensor_tables(self._params, nz, nth, a)
            Aint, Bd = Aint[0].cpu().numpy(), Bd[0].cpu().numpy()
            A = np.zeros((lz, lz), dtype=complex)
            D = np.zeros((lz, lz, lz, lz), dtype=complex)
            for i in range(lz):                                           # :112-118, the reference's (transposed) fill
                A[i, ::-1] = Aint[i + 2 * nz:i + 4 * nz + 1]
                for j in range(lz):
                    for k in range(lz):
                        D[i, j, k, ::-1] = Aint[i + j - k + 2 * nz:i + j - k + 4 * nz + 1]
            Bm = np.zeros((lz, lz, lth, lth), dtype=complex)              # :134-147: only beta' == beta is non-zero
            for b in range(lth):
                Bm[:, :, b, b] = Bd[:, :, b]
            if len(self._tensors) > 64:
                self._tensors.clear()
            self._tensors[key] = (Aint, A, D, Bm)
        return self._tensors[key]

    def evaluate_A_integrals(self, amplitude):
        """(:103-119) fills tmp_A_integrals / tmp_A_matrix / tmp_D_matrix / tmp_A_integrals_0 for `amplitude`."""
        self._tmp_amplitude = float(amplitude)
        self._A0_override = None
        self._tables(amplitude)

    def evaluate_B_integrals(self, amplitude):
        """(:134-147) fills tmp_B_matrix for `amplitude`."""
        self._tmp_amplitude = float(amplitude)
        self._tables(amplitude)

    def evaluate_A_integral_0(self, amplitude):
        """(:122-131) the diff = 0 element alone (the surface area integral)."""
        nz = self.num_field_coeffs_z
        self._A0_override = complex(self._tables(amplitude)[0][4 * nz])

    tmp_A_integrals = property(lambda self: self._tables(self._tmp_amplitude)[0])
    tmp_A_matrix = property(lambda self: self._tables(self._tmp_amplitude)[1])
    tmp_D_matrix = property(lambda self: self._tables(self._tmp_amplitude)[2])
    tmp_B_matrix = property(lambda self: self._tables(self._tmp_amplitude)[3])
    A_matrix = property(lambda self: self._tables(self._amplitude)[1])
    D_matrix = property(lambda self: self._tables(self._amplitude)[2])
    B_matrix = property(lambda self: self._tables(self._amplitude)[3])

    @property
    def tmp_A_integrals_0(self):
        """As the reference leaves it: element [0] of the list after evaluate_A_integrals (:119 -- the diff = -4 nz element, not the
        area, unless nz = 0), the diff = 0 integral after evaluate_A_integral_0."""
        return self._A0_override if self._A0_override is not None else self.tmp_A_integrals[0]

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
