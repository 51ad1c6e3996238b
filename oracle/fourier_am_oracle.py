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

TAG_FREAL, TAG_FCPLX, TAG_FALL = 3, 4, 5


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

    def run(self, n_steps, measure_every, fieldsteps, seed, chain, step0=0, method="sequential"):
        """method: the driver of run.py:286-317 -- "sequential", "simultaneous" (engine step_all), "fixed-amplitude",
        "no-field"."""
        do_real = method in ("sequential", "no-field")
        do_cplx = method in ("sequential", "fixed-amplitude")
        joint = method == "simultaneous"
        has_field = method != "no-field"
        per_iter = (1 if do_real else 0) + (fieldsteps if do_cplx else 0) + (1 if joint else 0)
        move = step0 * measure_every * per_iter
        rr = rm_ratio(self.target, self.d if joint else 1)
        rc = rm_ratio(self.target, max(1, self.ncoef))
        if not has_field:
            self.e_field = 0.0
        block = (lambda: self.cov) if joint else (lambda: self.cov[1:, 1:])
        L = np.linalg.cholesky(block()) if has_field else None
        for _ in range(n_steps):
            for _ in range(measure_every):
                if do_real:
                    z, u = _draw(0, chain, move, TAG_FREAL, seed)
                    a_new = self.a + self.sig_r * math.sqrt(self.cov[0, 0]) * float(z)
                    accept = False
                    if abs(a_new) < self.bound:
                        ef = self.energy(a_new, self.c) if has_field else 0.0
                        es = surface_energy_1d(self.P, a_new)
                        accept = self.accept_rule((ef + es) - (self.e_field + self.e_surf), float(u))
                        if accept:
                            self.a, self.e_field, self.e_surf = a_new, ef, es
                    self.n_real += 1
                    self.acc_r += accept
                    f, cc = max(self.n_real, 200.0), self.sig_r * rr
                    self.sig_r = self.sig_r + cc * (1 - self.target) / f if accept else self.sig_r - cc * self.target / f
                    move += 1
                if joint:
                    lanes = np.arange(self.d)
                    z, u = _draw(lanes, chain, move, TAG_FALL, seed)
                    zph, _ = _draw(lanes + 64, chain, move, TAG_FALL, seed)
                    dv = L @ z
                    mag0 = abs(self.a) + self.sig_r * dv[0]
                    a_new = math.copysign(mag0, self.a) if self.a != 0 else mag0
                    mag = np.abs(self.c) + self.sig_r * dv[1:]
                    ph = np.angle(self.c) + 0.1 * zph[1:]
                    cp = mag * np.cos(ph) + 1j * mag * np.sin(ph)
                    accept = False
                    if abs(a_new) < self.bound:
                        ef, es = self.energy(a_new, cp), surface_energy_1d(self.P, a_new)
                        accept = self.accept_rule((ef + es) - (self.e_field + self.e_surf), float(u[0]))
                        if accept:
                            self.a, self.c, self.e_field, self.e_surf = a_new, cp, ef, es
                    self.n_real += 1
                    self.acc_r += accept
                    f, cc = max(self.n_real / self.d, 200.0), self.sig_r * rr
                    self.sig_r = self.sig_r + cc * (1 - self.target) / f if accept else self.sig_r - cc * self.target / f
                    move += 1
                for _ in range(fieldsteps if do_cplx else 0):
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
            dm_ = self.d if has_field else 1
            x = np.concatenate([[abs(self.a)], np.abs(self.c)])[:dm_]
            mu_old = self.mean[:dm_].copy()
            self.mean[:dm_] = mu_old * (k - 1) / k + x / k
            if k > 50:
                mu = self.mean[:dm_]
                self.cov[:dm_, :dm_] = self.cov[:dm_, :dm_] * (k - 2) / (k - 1) + (
                    np.outer(mu_old, mu_old) - k / (k - 1) * np.outer(mu, mu) + np.outer(x, x) / (k - 1)
                    + (self.sig_r ** 2 / k) * np.identity(dm_))
                if has_field:
                    L = np.linalg.cholesky(block())
