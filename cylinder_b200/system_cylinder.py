"""Drop-in mirror of surfaces_and_fields/system_cylinder.py (`Cylinder`): bare sinusoidally perturbed cylinder
r(z) = r(a)(1 + a sin kz) with surface tension and bending rigidity.  Same constructor, method names and argument
order as the reference; every number is evaluated by the CUDA library (cyl_geometry_eval_f64, cyl_surface_energy_f64).
"""
import torch

from . import ops


class Cylinder:
    def __init__(self, wavenumber, radius, kappa, gamma=1, intrinsic_curvature=0, device=None):
        assert all(map(lambda x: x >= 0, [wavenumber, radius, kappa, gamma]))       # system_cylinder.py:16
        self.wavenumber = wavenumber
        self.radius = radius
        self.kappa = kappa
        self.gamma = gamma
        self.intrinsic_curvature = intrinsic_curvature
        self.effective_gamma = gamma + kappa / 2 * intrinsic_curvature ** 2          # :22
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._params = ops.make_params(self.device, wavenumber=wavenumber, radius=radius, gamma=gamma, kappa=kappa,
                                       intrinsic_curvature=intrinsic_curvature, alpha=0.0, u=1.0, C=0.0, n=0.0, temperature=0.0)

    # ---- metric factors (:25-41); columns of cyl_geometry_eval_f64 ----
    def _geo(self, amplitude, z):
        a = torch.tensor([float(amplitude)], dtype=torch.float64, device=self.device)
        zz = torch.tensor([float(z)], dtype=torch.float64, device=self.device)
        return ops.geometry_eval(self.wavenumber, self.radius, a, zz)[0].tolist()

    def radius_rescaled(self, amplitude):
        return self._geo(amplitude, 0.0)[0]

    def sqrt_g_theta(self, amplitude, z):
        return self._geo(amplitude, z)[1]

    def g_theta(self, amplitude, z):
        return self._geo(amplitude, z)[2]

    def sqrt_g_z(self, amplitude, z):
        return self._geo(amplitude, z)[3]

    def g_z(self, amplitude, z):
        return self._geo(amplitude, z)[4]

    def A_theta(self, amplitude, z):
        return self._geo(amplitude, z)[5]

    # ---- energies (:91-124) ----
    def _surface_terms(self, amplitude):
        a = torch.tensor([float(amplitude)], dtype=torch.float64, device=self.device)
        E, terms = ops.surface_energy(self._params, a, return_terms=True)
        return E.item(), terms[0, 0].item(), terms[0, 1].item()

    def evaluate_A_integral_0(self, amplitude):
        return self._surface_terms(amplitude)[1]

    def calc_bending_energy(self, amplitude):
        return self._surface_terms(amplitude)[2]

    def calc_surface_energy(self, amplitude):
        return self._surface_terms(amplitude)[0]
