"""Batched chains of the Fourier-mode model (config 2): what `run.single_run(field_type="fourier", method="sequential")`
(run.py:240-294) does for ONE chain through the host engine, here for B chains in one kernel
(cyl_fourier_am_steps_f64): Cylinder2D energies + the adaptive engine (real group = amplitude, complex group = the
(2 nth + 1) x (2 nz + 1) field coefficients, covariance-adapted magnitude proposals, phase jitter)."""
import numpy as np
import torch

from . import _abi, ops
from ._abi import call, ptr, stream
from .metropolisengine import robbins_monro_ratio

HDR = 16
METHODS = {"sequential": 0, "simultaneous": 1, "fixed-amplitude": 2, "no-field": 3}          # run.py:286-317


class FourierChains:
    def __init__(self, params, num_field_coeffs=(2, 1), amplitude=0.0, field_coeffs=None, sampling_width=.05, target_acceptance=.3,
                 device=None, seed=0, chain0=0, init_sigma=.1):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        _abi.lib()
        self.params = ops.make_params(self.device, **params)
        self.B = self.params.shape[0]
        self.nz, self.nth = int(num_field_coeffs[0]), int(num_field_coeffs[1])
        self.ncoef = (2 * self.nz + 1) * (2 * self.nth + 1)
        self.d = self.ncoef + 1
        self.seed, self.chain0, self.steps_done = int(seed), int(chain0), 0
        self.target = float(target_acceptance)
        self.amp = torch.as_tensor(amplitude, dtype=torch.float64).expand(self.B).contiguous().to(self.device)
        if field_coeffs is None:
            # gaussian_complex(sigma=.1) = rect(N(0, sigma), U(0, 2 pi))  (run.py:219), from a seeded torch generator
            g = torch.Generator(device="cpu").manual_seed(self.seed)
            mag = torch.randn(self.B, self.ncoef, generator=g, dtype=torch.float64) * init_sigma
            ph = torch.rand(self.B, self.ncoef, generator=g, dtype=torch.float64) * (2 * np.pi)
            field_coeffs = torch.polar(mag.abs(), ph + (mag < 0) * np.pi)
        self.coeffs = torch.as_tensor(field_coeffs).to(torch.complex128).reshape(self.B, self.ncoef).contiguous().to(self.device)
        n = _abi.lib().cyl_fourier_am_state_doubles(self.nz, self.nth)
        assert n == HDR + self.d + self.d * self.d
        st = torch.zeros(self.B, n, dtype=torch.float64)
        st[:, 0] = st[:, 1] = float(sampling_width)
        st[:, 2:5] = 1.0                                         # counters start at 1 (predecessor :39-42)
        x0 = torch.cat([self.amp.abs().cpu()[:, None], self.coeffs.abs().cpu()], dim=1)
        st[:, HDR:HDR + self.d] = x0                            # running mean starts from the initial state as sample 1
        st[:, HDR + self.d:] = torch.eye(self.d, dtype=torch.float64).reshape(1, -1)
        self.state = st.to(self.device)

    LAYOUTS = {None: 0, "cta": 0x100, "warp": 0x200}          # CYL_FOURIER_LAYOUT_* (default: chosen from the chain count)

    def run(self, n_steps, measure_every=1, fieldsteps_per_ampstep=1, method="sequential", layout=None):
        """n_steps measurements, each after measure_every iterations of the chosen driver (run.py:286-317).  A chain keeps
        one method for its whole life (the Philox move index advances by the method's moves per iteration).  `layout`
        ("cta" / "warp") forces the execution layout; results do not depend on it."""
        m = METHODS[method]
        # Robbins-Monro constants: m = 1 for the amplitude, m = ncoef for the coefficient group, m = d for the joint move
        ratio_real = robbins_monro_ratio(self.target, self.d if m == 1 else 1)
        call("cyl_fourier_am_steps_f64", self.device.index, ptr(self.params), self.nz, self.nth, ptr(self.amp), ptr(self.coeffs),
             ptr(self.state), self.B, self.seed, self.chain0, self.steps_done, int(n_steps), int(measure_every),
             int(fieldsteps_per_ampstep), self.target, ratio_real, robbins_monro_ratio(self.target, max(1, self.ncoef)),
             m | self.LAYOUTS[layout], stream())
        self.steps_done += int(n_steps)
        return self

    # ---- results ----
    @property
    def mean(self):
        """Running means of (|a|, |c_i|) per chain, [B, d]."""
        return self.state[:, HDR:HDR + self.d]

    @property
    def covariance(self):
        return self.state[:, HDR + self.d:].reshape(self.B, self.d, self.d)

    def summary(self):
        s = self.state
        nm = (s[:, 4] - 1).clamp(min=1)
        return dict(sigma_real=s[:, 0], sigma_complex=s[:, 1], real_acceptance=s[:, 5] / (s[:, 2] - 1).clamp(min=1),
                    complex_acceptance=s[:, 6] / (s[:, 3] - 1).clamp(min=1), field_energy=s[:, 7], surface_energy=s[:, 8],
                    amplitude_squared=s[:, 9] / nm, mean_energy=s[:, 10] / nm, abs_amplitude=s[:, HDR])
