"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package `cylinder_b200`).

Restatement, in numpy, of the schedule that `cyl_fourier_am_steps_f64` runs for one chain of the Fourier-mode model:
the reference's driver loop run.py:286-294 (method "sequential") over the engine arithmetic of SURVEY.md Appendix B
(predecessor bytecode __pycache__/metropolis_engine.cpython-36.pyc), with the kernel's Philox4x32-10 draws and a
Cholesky factor for the multivariate-normal proposal.  Energies come from oracle/fourier_oracle.py (pinned to golden
vectors of the reference).  Parity status of the engine arithmetic itself: UNPINNED (metropolisengine 0.2.19 is not in the
reference tree); what IS checked is that the CUDA kernel does exactly what is written here."""
import math

import numpy as np

from . import checkerboard_oracle as cb
from . import fourier_oracle as fo
from . import lattice_oracle as lo

TAG_FREAL, TAG_FCPLX = 3, 4


def _draw(word0, chain, move, tag, seed):
    """(unit normal, uniform) of csrc/cyl_site.cuh:amp_draw from one Philox4x32-10 block."""
    x0, x1, x2, _ = cb.philox4x32_10(word0, chain & 0xFFFFFFFF, move & 0xFFFFFFFF, cb.philox_c3(tag, move), seed & 0xFFFFFFFF, seed >> 32)
    z = np.sqrt(-2.0 * np.log(cb.u01(x0))) * np.cos(2.0 * np.pi * cb.u01(x1))
    return z, cb.u01(x2)


def rm_ratio(target, m):
    return lo.rm_ratio(target, m)


def surface_energy_1d(P, a):
    """gamma_eff * int Gth Gz dz + kappa * bending  (system_cylinder1D.py:189-194), area by the closed form."""
    H0 = P.get("intrinsic_curvature", 0.0)
    gamma_eff = P["gamma"] + P["kappa"] / 2 * H0 ** 2
    bend = lo.bending_energy(P["wavenumber"], P["radius"], a, H0) if P["kappa"] != 0 else 0.0
    return gamma_eff * lo.area_integral_0(P["wavenumber"], P["radius"], a) + P["kappa"] * bend


class Chain:
    def __init__(self, P, nz, nth, a0, c0, sampling_width=.05, target=.3, bound=.99):
        self.P, self.nz, self.nth = P, nz, nth
        self.a, self.c = float(a0), np.array(c0, dtype=complex).ravel()
        self.ncoef, self.d = self.c.size, self.c.size + 1
        self.sig_r = self.sig_c = sampling_width
        self.n_real = self.n_cplx = self.n_meas = 1
        self.acc_r = self.acc_c = 0
        self.target, self.bound = target, bound
        self.mean = np.concatenate([[abs(self.a)], np.abs(self.c)])
        self.cov = np.identity(self.d)
        self.e_field = self.energy(self.a, self.c)
        self.e_surf = surface_energy_1d(P, self.a)

    def energy(self, a, c):
        return fo.field_energy_realspace(self.P, self.nz, self.nth, a, c.reshape(2 * self.nth + 1, 2 * self.nz + 1))

    def accept_rule(self, diff, u):
        T = self.P["temperature"]
        return diff <= 0 or (T > 0 and u <= math.exp(-diff / T))

    def run(self, n_steps, measure_every, fieldsteps, seed, chain, step0=0):
        move = step0 * measure_every * (1 + fieldsteps)
        rr, rc = rm_ratio(self.target, 1), rm_ratio(self.target, max(1, self.ncoef))
        L = np.linalg.cholesky(self.cov[1:, 1:])
        for _ in range(n_steps):
            for _ in range(measure_every):
                z, u = _draw(0, chain, move, TAG_FREAL, seed)
                a_new = self.a + self.sig_r * math.sqrt(self.cov[0, 0]) * float(z)
                accept = False
                if abs(a_new) < self.bound:
                    ef, es = self.energy(a_new, self.c), surface_energy_1d(self.P, a_new)
                    accept = self.accept_rule((ef + es) - (self.e_field + self.e_surf), float(u))
                    if accept:
                        self.a, self.e_field, self.e_surf = a_new, ef, es
                self.n_real += 1
                self.acc_r += accept
                f, cc = max(self.n_real, 200.0), self.sig_r * rr
                self.sig_r = self.sig_r + cc * (1 - self.target) / f if accept else self.sig_r - cc * self.target / f
                move += 1
                for _ in range(fieldsteps):
                    lanes = np.arange(self.ncoef)
                    z, u = _draw(lanes, chain, move, TAG_FCPLX, seed)
                    zph, _ = _draw(lanes + 64, chain, move, TAG_FCPLX, seed)
                    dm = L @ z
                    mag = np.abs(self.c) + self.sig_c * dm
                    ph = np.angle(self.c) + 0.1 * zph
                    cp = mag * np.cos(ph) + 1j * mag * np.sin(ph)
                    ef = self.energy(self.a, cp)
                    accept = self.accept_rule(ef - self.e_field, float(u[0]))
                    if accept:
                        self.c, self.e_field = cp, ef
                    self.n_cplx += 1
                    self.acc_c += accept
                    f, cc = max(self.n_cplx / self.ncoef, 200.0), self.sig_c * rc
                    self.sig_c = self.sig_c + cc * (1 - self.target) / f if accept else self.sig_c - cc * self.target / f
                    move += 1
            self.n_meas += 1
            k = self.n_meas
            x = np.concatenate([[abs(self.a)], np.abs(self.c)])
            mu_old = self.mean.copy()
            self.mean = mu_old * (k - 1) / k + x / k
            if k > 50:
                self.cov = self.cov * (k - 2) / (k - 1) + (np.outer(mu_old, mu_old) - k / (k - 1) * np.outer(self.mean, self.mean)
                                                           + np.outer(x, x) / (k - 1) + (self.sig_r ** 2 / k) * np.identity(self.d))
                L = np.linalg.cholesky(self.cov[1:, 1:])
