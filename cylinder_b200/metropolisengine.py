"""Host-side mirror of `metropolisengine.MetropolisEngine` (third-party, version string "0.2.19", un-vendored in the
reference: run.py:14,19; On_lattice_simple.py:15,80).

The engine owns the GLOBAL parameters of a run -- the real group (shape amplitude a) and, for the Fourier model, the
complex group (field coefficients) -- and calls user-supplied energy functions.  It is control logic around
O(1)-sized state; the energy functions it calls are the CUDA kernels of this package.  The per-site lattice hot loop
never goes through it (On_lattice_simple.py:670 only borrows `metropolis_decision`).

What is restated here and from where (parity is UNPINNED for everything except the accept
rule, because the package source is not in the reference tree):
  * API surface: the reference's call sites -- ctor On_lattice_simple.py:80-82, run.py:237-238,257-258;
    set_reject_condition :83, run.py:281; step_real_group :544, run.py:290,316; step_complex_group run.py:293,309;
    step_all run.py:298; measure run.py:294,302,310,317; measure_real_system On_lattice_simple.py:530,563;
    save_time_series / df run.py:325-326; save_equilibrium_stats run.py:342-346; gaussian_complex run.py:219.
  * accept rule: predecessor bytecode __pycache__/metropolis_engine.cpython-36.pyc (src :226-246), pinned by the
    five known answers of tests/__pycache__/test_metropolis_engine*.pyc.
  * Robbins-Monro widths, recursive mean / covariance: predecessor bytecode :316-397 and the inline copy of the
    same rule at On_lattice_simple.py:95-100,688-696 (Garthwaite, Fan & Sisson, arXiv:1006.3690).
Random draws come from Python's `random` / numpy's global RandomState exactly like the reference, so seeded runs
consume the same streams.
"""
import cmath
import math
import random

import numpy as np


def _ppf(p):
    """Standard normal quantile (Acklam's rational approximation + one Halley step; |error| < 1e-15) -- avoids a
    scipy dependency on the product side."""
    if p <= 0 or p >= 1:
        raise ValueError("p must be in (0,1)")
    a = [-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02,
         -3.066479806614716e+01, 2.506628277459239e+00]
    b = [-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01,
         -1.328068155288572e+01]
    c = [-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00,
         4.374664141464968e+00, 2.938163982698783e+00]
    d = [7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00]
    pl = 0.02425
    if p < pl:
        q = math.sqrt(-2 * math.log(p))
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1)
    elif p <= 1 - pl:
        q = p - 0.5
        r = q * q
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1)
    else:
        q = math.sqrt(-2 * math.log(1 - p))
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1)
    e = 0.5 * math.erfc(-x / math.sqrt(2)) - p
    u = e * math.sqrt(2 * math.pi) * math.exp(x * x / 2)
    return x - u / (1 + x * u / 2)


def robbins_monro_ratio(target_acceptance, m):
    """The constant of On_lattice_simple.py:96-100 / predecessor :316-321."""
    alpha = -_ppf(target_acceptance / 2)
    return ((1 - (1 / m)) * math.sqrt(2 * math.pi) * math.exp(alpha ** 2 / 2) / 2 * alpha
            + 1 / (m * target_acceptance * (1 - target_acceptance)))


class MetropolisEngine:
    def __init__(self, energy_functions, initial_complex_params=None, initial_real_params=None,
                 covariance_matrix_complex=None, sampling_width=0.05, temp=0, complex_sample_method=None,
                 target_acceptance=.3, covariance_matrix_real=None, params_names=None, observables_names=None):
        self.energy_functions = energy_functions
        self.real_params = [float(x) for x in (initial_real_params if initial_real_params is not None else [])]
        self.complex_params = None if initial_complex_params is None else np.array(initial_complex_params, dtype=complex)
        self.num_real = len(self.real_params)
        self.num_complex = 0 if self.complex_params is None else len(self.complex_params)
        self.temp = temp
        self.complex_sample_method = complex_sample_method
        self.target_acceptance = target_acceptance
        # dimension of the sampled space, used by the adaptive rules (predecessor :32,:317); run.py:305,313 edits it
        self.m = max(1, self.num_real + self.num_complex)
        if isinstance(sampling_width, (list, tuple)):                  # the pickled [real_w, complex_w] of run.py:211,335
            self.real_group_sampling_width, self.complex_group_sampling_width = float(sampling_width[0]), float(sampling_width[1])
        else:
            self.real_group_sampling_width = self.complex_group_sampling_width = float(sampling_width)
        self.sampling_width = self.real_group_sampling_width
        self.covariance_matrix_real = np.identity(max(1, self.num_real)) if covariance_matrix_real is None else np.array(covariance_matrix_real, dtype=float)
        if covariance_matrix_complex is None:
            self.covariance_matrix_complex = np.identity(max(1, self.num_complex), dtype=complex)
        else:
            self.covariance_matrix_complex = np.array(covariance_matrix_complex, dtype=complex)
        self.covariance_matrix = np.identity(self.m)                    # joint [|a|, |c_i|...] (step_all)
        self.reject_condition = lambda real_params, complex_params: False
        # counters start at 1 (predecessor :39-42)
        self.real_group_step_counter = 1
        self.complex_group_step_counter = 1
        self.step_counter = 1
        self.measure_step_counter = 1
        self.accepted_counter = 0
        self.real_group_accepted = 0
        self.complex_group_accepted = 0
        self.params_names = params_names
        self.observables_names = observables_names
        # running statistics
        self.mean = self._state()
        self.observables = self._observables()
        self.time_series = []
        self.df = None
        self.equilibrated_means = None
        self.field_profile = None
        self.field_abs_profile = None
        # initial energies: every term any group names, evaluated once (run.py:282 relies on this)
        self.energy = {}
        for group in ("all", "real", "complex"):
            for term, fn in self.energy_functions.get(group, {}).items():
                if term not in self.energy:
                    self.energy[term] = self._real(fn(self.real_params, self.complex_params))

    # ---- helpers -----------------------------------------------------------------------------------------
    @staticmethod
    def _real(x):
        """Energy terms may come back complex with a ~0 imaginary part (system_cylinder1D.py:124,194)."""
        return float(x.real) if isinstance(x, complex) or np.iscomplexobj(x) else float(x)

    def _state(self):
        s = [abs(x) for x in self.real_params]
        if self.complex_params is not None:
            s.extend(np.abs(self.complex_params).tolist())
        return np.array(s, dtype=float)

    def _observables(self):
        """abs_<param> for every parameter, then amplitude_squared (run.py:274-276)."""
        o = self._state().tolist()
        o.append(self.real_params[0] ** 2 if self.real_params else 0.0)
        return np.array(o, dtype=float)

    def set_reject_condition(self, fn):
        self.reject_condition = fn

    def set_temperature(self, temp):
        assert temp >= 0
        self.temp = temp

    @staticmethod
    def gaussian_complex(sigma=0.1):                                     # predecessor :284-293
        return cmath.rect(random.gauss(0, sigma), random.uniform(0, 2 * math.pi))

    def metropolis_decision(self, old_energy, proposed_energy):
        """diff <= 0 -> accept without a draw; T == 0 -> reject; else uniform(0,1) <= exp(-diff/T)."""
        diff = proposed_energy - old_energy
        if diff <= 0:
            return True
        if self.temp == 0:
            return False
        return random.uniform(0, 1) <= math.exp(-diff / self.temp)

    def _rm_update(self, width, accept, counter, m):
        """Robbins-Monro step of On_lattice_simple.py:688-696 with this engine's target acceptance."""
        ratio = robbins_monro_ratio(self.target_acceptance, m)
        f = max(counter / m, 200)
        c = width * ratio
        width = width + c * (1 - self.target_acceptance) / f if accept else width - c * self.target_acceptance / f
        assert width > 0
        return width

    def _group_energy(self, group, real, cplx):
        terms = self.energy_functions[group]
        return {t: self._real(fn(real, cplx)) for t, fn in terms.items()}

    def _decide(self, group, real, cplx):
        new = self._group_energy(group, real, cplx)
        old = sum(self.energy[t] for t in new)
        accept = self.metropolis_decision(old, sum(new.values()))
        if accept:
            self.energy.update(new)
        return accept

    # ---- moves -------------------------------------------------------------------------------------------
    def step_real_group(self):
        """One move of the real parameters (the shape amplitude).  Returns the accept flag."""
        scale = [self.real_group_sampling_width * math.sqrt(self.covariance_matrix_real[i, i]) for i in range(self.num_real)]
        proposed = [x + random.gauss(0, s) for x, s in zip(self.real_params, scale)]
        if self.reject_condition(proposed, self.complex_params):
            accept = False                                               # rejected outright, counts as a rejection
        else:
            accept = self._decide("real", proposed, self.complex_params)
        if accept:
            self.real_params = proposed
            self.real_group_accepted += 1
        self.real_group_step_counter += 1
        self.real_group_sampling_width = self._rm_update(self.real_group_sampling_width, accept, self.real_group_step_counter, 1)
        self.sampling_width = self.real_group_sampling_width
        return accept

    def _propose_complex(self, width, cov):
        n = self.num_complex
        cov_r = np.real(cov)
        if self.complex_sample_method == "magnitude-phase":              # predecessor :213-217, modify_phase :266-273
            dm = np.random.multivariate_normal(np.zeros(n), width ** 2 * cov_r, check_valid="raise")
            out = np.empty(n, dtype=complex)
            for i, c in enumerate(self.complex_params):
                out[i] = cmath.rect(abs(c) + dm[i], cmath.phase(c) + random.gauss(0, 0.1))
            return out
        dre = np.random.multivariate_normal(np.zeros(n), width ** 2 * cov_r, check_valid="raise")
        dim = np.random.multivariate_normal(np.zeros(n), width ** 2 * cov_r, check_valid="raise")
        return self.complex_params + dre + 1j * dim

    def step_complex_group(self):
        proposed = self._propose_complex(self.complex_group_sampling_width, self.covariance_matrix_complex)
        if self.reject_condition(self.real_params, proposed):
            accept = False
        else:
            accept = self._decide("complex", self.real_params, proposed)
        if accept:
            self.complex_params = proposed
            self.complex_group_accepted += 1
        self.complex_group_step_counter += 1
        self.complex_group_sampling_width = self._rm_update(self.complex_group_sampling_width, accept,
                                                            self.complex_group_step_counter, max(1, self.num_complex))
        return accept

    def step_all(self):
        """Joint move of [|a|, |c_i|] from N(0, sigma^2 Sigma) + phase jitter (predecessor :213-218)."""
        n = self.m
        d = np.random.multivariate_normal(np.zeros(n), self.sampling_width ** 2 * self.covariance_matrix[:n, :n], check_valid="raise")
        real = []
        for i, x in enumerate(self.real_params):
            mag = abs(x) + d[i]
            real.append(math.copysign(mag, x) if x != 0 else mag)       # sign of a restored (:218)
        cplx = None
        if self.complex_params is not None:
            cplx = np.array([cmath.rect(abs(c) + d[self.num_real + i], cmath.phase(c) + random.gauss(0, 0.1))
                             for i, c in enumerate(self.complex_params)])
        if self.reject_condition(real, cplx):
            accept = False
        else:
            accept = self._decide("all", real, cplx)
        if accept:
            self.real_params, self.complex_params = real, cplx
            self.accepted_counter += 1
        self.step_counter += 1
        self.sampling_width = self._rm_update(self.sampling_width, accept, self.step_counter, self.m)
        return accept

    # ---- measurement -------------------------------------------------------------------------------------
    def measure(self):
        """Running mean of the state and of the observables; covariance adaptation once 50 samples are in
        (predecessor :376-397)."""
        self.measure_step_counter += 1
        k = self.measure_step_counter
        x = self._state()
        old_mean = self.mean.copy()
        self.mean = self.mean * (k - 1) / k + x / k                     # :60-63
        if k > 50:
            n = len(x)
            cov = self.covariance_matrix[:n, :n]
            cov = cov * (k - 2) / (k - 1) + (np.outer(old_mean, old_mean) - k / (k - 1) * np.outer(self.mean, self.mean)
                                             + np.outer(x, x) / (k - 1) + (self.sampling_width ** 2 / k) * np.identity(n))
            self.covariance_matrix = cov
            if self.num_complex:
                self.covariance_matrix_complex = cov[self.num_real:, self.num_real:].astype(complex)
            if self.num_real:
                self.covariance_matrix_real = cov[:self.num_real, :self.num_real]
        self.observables = self.observables * (k - 1) / k + self._observables() / k   # :66-68
        self._record()

    def measure_real_system(self):
        """Lattice runs: record the amplitude and the energies the lattice assigned (On_lattice_simple.py:562-563)."""
        self.measure_step_counter += 1
        k = self.measure_step_counter
        self.observables = self.observables * (k - 1) / k + self._observables() / k
        self._record()

    def _record(self):
        row = {}
        names = self.params_names or (["amplitude"] + ["coeff_%d" % i for i in range(self.num_complex)])
        for name, v in zip(names, list(self.real_params) + ([] if self.complex_params is None else list(self.complex_params))):
            row[name] = v
        for name, v in zip(self.observables_names or [], self._observables()):
            row[name] = v
        for term, e in self.energy.items():
            row[term + "_energy"] = e
        row["energy"] = sum(self.energy.values())
        self.time_series.append(row)

    def save_time_series(self):
        import pandas as pd
        self.df = pd.DataFrame(self.time_series)

    def save_equilibrium_stats(self, profiles=None, burn_in_fraction=.5):
        """Means over the equilibrated part of the run (the second half of the recorded time series), keyed by the
        observable names -- the columns parallel_single_run.py:113 writes to <var1>_<v1>_<var2>_<v2>_mean.csv -- and the
        time-averaged field profiles when `profiles` = Lattice.get_profiles_df() is given (run.py:341-346)."""
        import pandas as pd
        if self.df is None:
            self.save_time_series()
        n = len(self.df)
        if n == 0:
            raise IndexError("no measurements recorded")
        cut = int(n * burn_in_fraction)
        eq = self.df.iloc[cut:]
        means = {}
        for col in eq.columns:
            vals = eq[col].to_numpy()
            means[col] = np.mean(np.abs(vals)) if np.iscomplexobj(vals) else float(np.mean(vals))
            if col in ("energy",) or col.endswith("_energy"):
                means[col + "_variance"] = float(np.var(np.real(vals)))
        for name, v in zip(self.observables_names or [], self.observables):
            means.setdefault(name, float(v))
        self.equilibrated_means = means
        if profiles is not None:
            prof, prof_abs = profiles
            c0 = int(len(prof) * burn_in_fraction)
            self.field_profile = prof.iloc[c0:].mean(axis=0)
            self.field_abs_profile = prof_abs.iloc[c0:].mean(axis=0)
        else:
            self.field_profile, self.field_abs_profile = pd.Series(dtype=complex), pd.Series(dtype=float)
        return means
