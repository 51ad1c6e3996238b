"""Drop-in mirror of the reference's lattice field model:
  Lattice                  = surfaces_and_fields/On_lattice_simple.py:17-696
  LatticeRescaled0thOrder  = surfaces_and_fields/On_lattice_rescaled_0thorder.py (the class run.py:18 imports)
  LatticeRescaled1stOrder  = surfaces_and_fields/On_lattice_rescaled_1storder.py
  LatticeDecoupled         = surfaces_and_fields/On_lattice_decoupled.py
Same constructor arguments, public attributes and methods; psi lives on the GPU and every energy / move is a kernel
of libcylinder_b200 (no CPU fallback).

Two ways to advance the chain:
  * step_lattice / step_lattice_loc  -- the reference's own sequential proposal, one site at a time: site and
    Gaussian draws from Python's `random` in the reference's order, dE on the device in the reference's operation
    order (cyl_site_delta_f64), accept through `self.me.metropolis_decision` (so the uniform is drawn exactly when
    the reference draws it), Robbins-Monro after every proposal.  A seeded run consumes the same random stream as
    the reference.  This is the compatibility path: one kernel + one read-back per proposal.
  * run / run_fixed_amplitude / sweep -- the production path: checkerboard sweeps with counter-based Philox draws on
    the device (cyl_sweep_smem_* for lattices that fit in shared memory, cyl_sweep_hbm_* above that), global amplitude
    move and measurement fused into the sweep kernel.  Statistically equivalent to the sequential ordering
    (tests/test_oracle_checkerboard.py, tests/test_gpu_statistics.py), not stream-identical to it.
Defects of the reference as committed and the semantics adopted instead are listed in SURVEY.md Appendix C.
"""
import math
import os
import random

import numpy as np
import torch

from . import _abi, ops
from ._abi import SWEEP_ADAPT_SIGMA, SWEEP_AMP_MOVES, SWEEP_MEASURE
from .metropolisengine import MetropolisEngine, robbins_monro_ratio
from .system_cylinder import Cylinder


class Lattice:
    VARIANT = "simple"
    INIT_MAGNITUDE = .1                       # On_lattice_simple.py:164
    INIT_SAMPLING_WIDTH = .005                # On_lattice_simple.py:101

    def __init__(self, amplitude, wavenumber, radius, gamma, kappa, intrinsic_curvature, alpha, u, C, n, temperature,
                 temperature_final, dims=(100, 50), n_substeps=None, fieldsteps_per_ampstep=1, dtype=torch.complex128,
                 device=None):
        # experiment variables (:21-35)
        self.wavenumber, self.radius, self.gamma, self.kappa = wavenumber, radius, gamma, kappa
        self.initial_amplitude = amplitude
        self.intrinsic_curvature = intrinsic_curvature
        self.alpha, self.u, self.C, self.n = alpha, u, C, n
        self.Cnsquared = self.C * self.n ** 2
        self.temperature, self.temperature_final = temperature, temperature_final     # temperature_final is unused (:34)
        self.fieldsteps_per_ampstep = fieldsteps_per_ampstep
        # lattice characteristics (:37-49)
        cylinder_len_z = 2 * math.pi / self.wavenumber
        cylinder_radius_th = 2 * math.pi * self.radius
        self.z_len = int(round(dims[0] / self.wavenumber))
        self.th_len = dims[1]
        self.n_substeps = self.z_len * self.th_len if n_substeps is None else n_substeps
        self.z_pixel_len = cylinder_len_z / self.z_len
        self.th_pixel_len = cylinder_radius_th / self.th_len
        self.amplitude = self.initial_amplitude
        # device state
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        _abi.lib()
        self.dtype = dtype
        self.variant = ops.variant_id(self.VARIANT)
        self._params = ops.make_params(self.device, wavenumber=wavenumber, radius=radius, gamma=gamma, kappa=kappa,
                                       intrinsic_curvature=intrinsic_curvature, alpha=alpha, u=u, C=C, n=n,
                                       temperature=temperature, amp_bound=.8, amp_target_acceptance=.3)
        self._Z = torch.tensor([self.z_len], dtype=torch.int32, device=self.device)
        self._psi_store = torch.zeros(1, self.z_len, self.th_len, dtype=dtype, device=self.device)
        self._psi_stale = False                                # True: the HBM workspace holds the newer field
        self._host_view = None                                 # numpy array last handed out by `.lattice`
        self._hbm, self._hbm_dirty = None, True
        self._delta_out = torch.empty(4, dtype=torch.float64, device=self.device)
        self.surface = Cylinder(wavenumber=wavenumber, radius=radius, gamma=gamma, kappa=kappa,
                                intrinsic_curvature=intrinsic_curvature, device=self.device)
        self._avg_dev = torch.zeros(self.z_len, self.th_len, dtype=torch.complex128, device=self.device)   # avg_lattice, on the device
        self.avg_amplitude_profile = np.zeros((self.z_len))
        self.random_initialize()
        # engine coupled to the bare surface + the field energy the surface change would cause (:72-83)
        energy_fct_surface_term = lambda real_params, complex_params: self.surface.calc_surface_energy(*real_params)   # noqa: E731
        energy_fct_field_term = lambda real_params, complex_params: self.surface_field_energy(                        # noqa: E731
            *real_params, self.amplitude, self.lattice, self.psi_squared, self.dz, self.dth)
        energy_fct_by_params_group = {"real": {"surface": energy_fct_surface_term, "field": energy_fct_field_term},
                                      "all": {"surface": energy_fct_surface_term, "field": energy_fct_field_term}}
        self.me = MetropolisEngine(energy_functions=energy_fct_by_params_group, initial_complex_params=None,
                                   initial_real_params=[float(self.initial_amplitude)], covariance_matrix_complex=None,
                                   sampling_width=.005, temp=self.temperature, complex_sample_method=None)
        self.me.set_reject_condition(lambda real_params, complex_params: abs(real_params[0]) > .8)
        self.lattice_acceptance_counter = 0
        self.amplitude_average = 0
        self.field_average = np.zeros((self.z_len))
        self.field_profile_history = []
        self.field_abs_profile_history = []
        # dynamic step size (:93-101)
        self.acceptance_counter = 0
        self.step_counter = 0
        self.bump_step_counter = 0
        self.target_acceptance = .5
        self.m = 1
        self.ratio = robbins_monro_ratio(self.target_acceptance, self.m)
        self.sampling_width = self.INIT_SAMPLING_WIDTH
        self.bump_sampling_width = .005
        # production path state: the Philox seed is a function of the state of Python's generator (random.seed(...) makes
        # runs repeatable) but does not CONSUME it -- the compatibility path stays on the reference's stream
        self._seed = hash(random.getstate()[1][:8]) & ((1 << 62) - 1)
        self._sweeps_done = 0
        self._chain = ops.new_chain(1, self.device, amplitude=float(self.amplitude), sigma_site=self.sampling_width,
                                    sigma_amp=self.me.real_group_sampling_width)
        self._chain_host = self._chain[0].tolist()
        self._obs, self._prof = ops.new_obs(1, self.z_len, self.device)
        self.energy = None

    # ---- lattice and derived arrays --------------------------------------------------------------------
    @property
    def _psi(self):
        """psi[1, Z, T] in the packed ABI layout.  Large lattices are swept in the tiled HBM workspace; the packed copy is
        refreshed from it only when something asks for it."""
        if self._psi_stale:
            self._hbm.unpack(self._psi_store[0])
            self._psi_stale = False
        return self._psi_store

    @property
    def avg_lattice(self):
        """Running time average of psi(z, theta) (:147-148), kept on the device, downloaded on access."""
        return self._avg_dev.cpu().numpy()

    @avg_lattice.setter
    def avg_lattice(self, value):
        self._avg_dev.copy_(torch.from_numpy(np.ascontiguousarray(value, dtype=complex)))

    @property
    def lattice(self):
        """numpy view of psi, shape (z_len, th_len), complex128 (downloaded; cached until the device copy changes)."""
        if self._host_view is None:
            self._host_view = self._psi[0].to(torch.complex128).cpu().numpy()
        return self._host_view

    @lattice.setter
    def lattice(self, value):
        v = np.ascontiguousarray(value, dtype=complex).reshape(self.z_len, self.th_len)
        self._psi[0].copy_(torch.from_numpy(v).to(self.dtype))
        self._invalidate()

    def _invalidate(self):
        self._host_view = None
        if self._hbm is not None:
            self._hbm_dirty = True

    @property
    def psi_squared(self):
        p = self._psi[0]
        return (p.real * p.real + p.imag * p.imag).to(torch.float64).cpu().numpy()

    @property
    def dz(self):                                               # left differences (:170-182)
        p = self._psi[0].to(torch.complex128)
        return ((p - torch.roll(p, 1, 0)) / self.z_pixel_len).cpu().numpy()

    @property
    def dth(self):
        p = self._psi[0].to(torch.complex128)
        return ((p - torch.roll(p, 1, 1)) / self.th_pixel_len).cpu().numpy()

    def squared(self, c):
        return abs(c * c.conjugate())

    def random_initialize(self):
        """psi = rect(U(0,.1), U(0,2pi)) row-major from Python's generator, as :159-167 (same stream as the reference)."""
        vals = np.empty((self.z_len, self.th_len), dtype=complex)
        for z_index in range(self.z_len):
            for th_index in range(self.th_len):
                mag = random.uniform(0, self.INIT_MAGNITUDE)
                ph = random.uniform(0, 2 * math.pi)
                vals[z_index, th_index] = complex(mag * math.cos(ph), mag * math.sin(ph))      # cmath.rect
        self._psi[0].copy_(torch.from_numpy(vals).to(self.dtype))
        self._invalidate()

    # ---- energies --------------------------------------------------------------------------------------
    def surface_field_energy(self, amplitude, reference_amplitude, lattice=None, psi_squared=None, dz=None, dth=None):
        """Total field energy at trial amplitude (:185-213; 0th-order :247-278).  `psi_squared`, `dz`, `dth` are pure
        functions of `lattice` and are derived on the device; they are accepted for signature
        compatibility.  Passing the array obtained from `.lattice` (or None) evaluates the device-resident field."""
        if lattice is None or lattice is self._host_view:
            if self._psi_stale:                                  # the field lives in the tiled HBM workspace: no repacking
                S = self._hbm.row_moments(self._params)[None]
                return ops.field_energy(S, self._Z, self.th_len, self._params, [float(amplitude)], [float(reference_amplitude)],
                                        self.variant).item()
            psi = self._psi
        else:
            psi = torch.from_numpy(np.ascontiguousarray(lattice, dtype=complex)).to(self.device)[None].contiguous()
        Z = torch.tensor([psi.shape[1]], dtype=torch.int32, device=self.device)
        S = ops.row_moments(psi.contiguous(), Z, self._params)
        return ops.field_energy(S, Z, psi.shape[2], self._params, [float(amplitude)], [float(reference_amplitude)],
                                self.variant).item()

    # ---- the reference's sequential proposal (compatibility path) --------------------------------------
    def step_lattice(self, amplitude):
        index_z, index_th = (random.randrange(-1, (self.z_len - 1)), random.randrange(-1, (self.th_len - 1)))   # :595
        self.step_lattice_loc(amplitude, index_z, index_th)

    def step_lattice_loc(self, amplitude, index_z, index_th):
        if self.dtype != torch.complex128:
            raise _abi.CylinderB200Error("step_lattice_loc is the fp64 compatibility path: construct the Lattice with dtype=torch.complex128")
        z_re, z_im = random.gauss(0, 1), random.gauss(0, 1)           # same draws as random.gauss(0, stepsize) twice (:617)
        out = ops.site_delta(self._psi[0], self._params, amplitude, self.variant, index_z, index_th, z_re, z_im,
                             self.sampling_width, self._delta_out).tolist()
        diff_energy = out[0]
        accept = self.me.metropolis_decision(0, diff_energy)           # :670
        if accept:
            self._psi[0, index_z % self.z_len, index_th % self.th_len] = complex(out[1], out[2])   # :671-685
            self._invalidate()
            self.acceptance_counter += 1
            self.lattice_acceptance_counter += 1
        self.update_sigma(accept)                                      # :686
        return accept, diff_energy

    def update_sigma(self, accept):                                    # :688-696
        self.step_counter += 1
        step_number_factor = max((self.step_counter / self.m, 200))
        steplength_c = self.sampling_width * self.ratio
        if accept:
            self.sampling_width += steplength_c * (1 - self.target_acceptance) / step_number_factor
        else:
            self.sampling_width -= steplength_c * self.target_acceptance / step_number_factor
        assert (self.sampling_width) > 0

    def update_rescale_params(self, amplitude):
        pass

    # ---- whole-lattice proposals (:247-382) ------------------------------------------------------------
    # The proposed field is never materialised: cyl_move_row_moments_* evaluates its row moments with the transform applied
    # on load, cyl_field_energy_f64 gives its energy, cyl_move_apply_* commits it on accept.  Random draws come from
    # Python's generator in the reference's order; the accept goes through the engine like the reference's.
    def _collective(self, amplitude, moves, old_energy):
        S = ops.move_row_moments(self._psi, self._Z, self._params, moves)
        new_energy = ops.field_energy(S, self._Z, self.th_len, self._params, [float(amplitude)], [float(amplitude)], self.variant).item()
        accept = self.me.metropolis_decision(old_energy, new_energy)
        if accept:
            ops.move_apply(self._psi, self._Z, moves)
            self._invalidate()
        return accept, new_energy

    def step_lattice_all(self, amplitude):
        """One complex number added to every site (:247-261); `self.energy` must hold the current field energy, as in the
        reference's run loop (:524)."""
        addition = random.gauss(0, self.sampling_width) + random.gauss(0, self.sampling_width) * 1j              # :248
        accept, new_energy = self._collective(amplitude, ops.make_moves(self.device, ops.MOVE_SHIFT, c=addition), self.energy)
        if accept:
            self.energy = new_energy
            self.me.energy["field"] = new_energy
        self.update_sigma(accept)
        return accept

    def step_lattice_random_wavevector(self, amplitude, qzmax=4, qthmax=8):                                     # :263-265
        wavevector = (random.randint(-qzmax, qzmax + 1) * 2 * math.pi, random.randint(-qthmax, qthmax + 1) * 2 * math.pi)
        return self.step_lattice_wavevector(amplitude, wavevector)

    def wave_array(self, base, wavevector, z_len, th_len):                                                       # :267-271
        qz, qth = wavevector
        return np.fromfunction(lambda i, j: base * np.exp(1j * (i * qz / z_len + j * qth / th_len)), (z_len, th_len))

    def step_lattice_wavevector(self, amplitude, wavevector):
        """A plane wave c exp(i(i qz/Z + j qth/T)) added to the field (:273-300), width `bump_sampling_width`."""
        addition = random.gauss(0, self.bump_sampling_width) + random.gauss(0, self.bump_sampling_width) * 1j    # :274
        moves = ops.make_moves(self.device, ops.MOVE_WAVE, c=addition, qz=wavevector[0], qth=wavevector[1])
        accept, new_energy = self._collective(amplitude, moves, self.energy)
        if accept:
            self.energy = new_energy
        self.update_bump_sigma(accept)
        return accept

    def step_row_rotate(self, amplitude, maxrows=1):
        """A block of rows gets one more (or one less) turn of phase around the cylinder plus a random global phase
        (:303-382): the move that changes the winding number of a row.  Same draws as the reference, in its order."""
        n_rot = random.choice([-1, 1])                                                                           # :304
        phase = random.uniform(0, 2 * math.pi)                                                                   # :305
        num_rows = random.randint(1, maxrows)                                                                    # :306
        row_index = random.randint(0, self.z_len - 1)                                                            # :310
        old_energy = self.surface_field_energy(amplitude, amplitude)
        moves = ops.make_moves(self.device, ops.MOVE_TWIST, r0=row_index, nr=num_rows, n_rot=n_rot, phase=phase)
        accept, new_energy = self._collective(amplitude, moves, old_energy)
        if accept:
            self.energy = new_energy
        return accept

    def update_bump_sigma(self, accept):                               # :700-708 (step factor from step_counter, as committed)
        self.bump_step_counter += 1
        step_number_factor = max((self.step_counter / self.m, 200))
        steplength_c = self.bump_sampling_width * self.ratio
        if accept:
            self.bump_sampling_width += steplength_c * (1 - self.target_acceptance) / step_number_factor
        else:
            self.bump_sampling_width -= steplength_c * self.target_acceptance / step_number_factor
        assert (self.bump_sampling_width) > 0

    # ---- production path: checkerboard sweeps on the device --------------------------------------------
    def _fits_shared_memory(self):
        """The batched kernel keeps the lattice AND its per-row tables in shared memory: ask the library for the real layout
        (cyl_sweep_smem_bytes) instead of guessing from the lattice alone, so tall narrow lattices take the HBM path."""
        need = _abi.lib().cyl_sweep_smem_bytes(int(self.z_len), int(self.th_len), int(self._psi_store.element_size() == 16))
        limit = torch.cuda.get_device_properties(self.device).shared_memory_per_block_optin
        return 0 < need <= limit

    def _push_chain(self):
        c = self._chain_host
        c[_abi.CH_AMPLITUDE] = float(self.amplitude)
        c[_abi.CH_SIGMA_SITE] = float(self.sampling_width)
        c[_abi.CH_SIGMA_AMP] = float(self.me.real_group_sampling_width)
        c[_abi.CH_STEP_COUNTER] = float(self.step_counter)
        self._chain[0].copy_(torch.tensor(c, dtype=torch.float64))

    def _pull_chain(self):
        c = self._chain_host = self._chain[0].tolist()
        self.amplitude = c[_abi.CH_AMPLITUDE]
        self.sampling_width = c[_abi.CH_SIGMA_SITE]
        self.me.real_group_sampling_width = self.me.sampling_width = c[_abi.CH_SIGMA_AMP]
        self.step_counter = int(c[_abi.CH_STEP_COUNTER])
        self.acceptance_counter = self.lattice_acceptance_counter = int(c[_abi.CH_SITE_ACCEPTED])
        self.me.real_params = [self.amplitude]
        return c

    def sweep(self, nsweeps=1, amp_moves=False, measure=False, adapt=True):
        """nsweeps checkerboard sweeps (every site proposed once per sweep), optionally each followed by one global
        amplitude move (run_lattice.py:72-90) and a measurement.  Returns the chain record (list of 16 floats)."""
        flags = (SWEEP_ADAPT_SIGMA if adapt else 0) | (SWEEP_AMP_MOVES if amp_moves else 0) | (SWEEP_MEASURE if measure else 0)
        self._params[0, _abi.PARAM_FIELDS.index("amp_bound")] = self._amp_bound()
        self._push_chain()
        if self._fits_shared_memory():
            ops.sweep_smem(self._psi, self._Z, self._params, self._chain, self._seed, self._sweeps_done, nsweeps, self.variant,
                           flags, 0, self._obs, self._prof)
        else:
            if self._hbm is None:
                self._hbm = ops.HbmWorkspace(self.z_len, self.th_len, self.dtype, self.device)
                self._hbm_dirty = True
            if self._hbm_dirty:
                self._hbm.pack(self._psi_store[0])                 # dirty implies the packed copy is the current one
            self._hbm.sweep(self._params, self._chain[0], self._seed, self._sweeps_done, nsweeps, self.variant, flags, 0,
                            self._obs[0], self._prof[0])
            self._hbm_dirty, self._psi_stale = False, True
        self._sweeps_done += nsweeps
        self._host_view = None
        return self._pull_chain()

    def _amp_bound(self):
        """Threshold of the engine's reject condition, found by bisection: |a| > .8 (:83) unless the driver replaced it
        (|a| >= .99, run.py:281).  The kernel tries a move iff |a'| < bound."""
        rc = self.me.reject_condition
        lo, hi = 0.0, 4.0
        if not rc([hi], None):
            return hi
        for _ in range(60):
            mid = .5 * (lo + hi)
            if rc([mid], None):
                hi = mid
            else:
                lo = mid
        return hi

    def _sweeps_for(self, n_proposals):
        return max(1, int(round(n_proposals / float(self.z_len * self.th_len))))

    def run(self, n_steps, n_sub_steps, sweeps_per_step=None):
        """Lattice.run (:532-563) with the adopted semantics of SURVEY.md Appendix C: burn-in of 500 x 50 proposals;
        then per step: measure, one amplitude move, fieldsteps_per_ampstep x 50 proposals, energy resync.  Proposals
        are executed as whole checkerboard sweeps (>= 1 per step)."""
        n_sub_steps = 50                                                  # :533 (the reference overrides the argument)
        self.sweep(self._sweeps_for(500 * n_sub_steps))
        per_step = sweeps_per_step or self._sweeps_for(self.fieldsteps_per_ampstep * n_sub_steps)
        for i in range(n_steps):
            self.measure_avgs()
            self.measure()
            if per_step > 1:
                self.sweep(per_step - 1)
            c = self.sweep(1, amp_moves=True, measure=True)               # sweep, measurement record, amplitude move
            self.update_rescale_params(self.amplitude)
            # energy resync at the amplitude the step ended with (:550,:560-562)
            self.energy = self.surface_field_energy(self.amplitude, self.amplitude)
            self.me.energy["field"] = self.energy
            self.me.energy["surface"] = c[_abi.CH_E_SURF]
            self.me.measure_real_system()

    def run_fixed_amplitude(self, n_steps, n_sub_steps, sweeps_per_step=None):
        """:519-530 -- field moves on a frozen shape: per step one whole-lattice shift proposal (step_lattice_all, :524-525),
        then the single-site proposals, executed as sweeps."""
        per_step = sweeps_per_step or self._sweeps_for(self.fieldsteps_per_ampstep * n_sub_steps)
        for i in range(n_steps):
            self.measure_avgs()
            self.measure()
            self.energy = self.surface_field_energy(self.amplitude, self.amplitude)            # :524
            self.step_lattice_all(self.amplitude)                                              # :525
            if per_step > 1:
                self.sweep(per_step - 1)
            c = self.sweep(1, amp_moves=False, measure=True)
            self.energy = c[_abi.CH_E_FIELD]
            self.me.energy["field"] = c[_abi.CH_E_FIELD]
            self.me.measure_real_system()

    # ---- observables -----------------------------------------------------------------------------------
    def measure(self):
        """Per z-row sum_theta psi and sum_theta |psi| appended to the histories (:121-131)."""
        if self._psi_stale:
            S = self._hbm.row_moments(self._params).cpu().numpy()       # straight from the tiled layout
        else:
            S = ops.row_moments(self._psi, self._Z, self._params)[0].cpu().numpy()
        self.field_profile_history.append(S[:, 6] + 1j * S[:, 7])
        self.field_abs_profile_history.append(S[:, 5].copy())

    def measure_avgs(self):
        """Running averages of |a| and of psi(z, theta) (:135-156); shares step_counter with update_sigma like the reference."""
        self.amplitude_average *= self.step_counter / float(self.step_counter + 1)
        self.amplitude_average += abs(self.amplitude) / float(self.step_counter + 1)
        assert (self.amplitude_average < 1)
        divisor = float(self.step_counter + 1)
        self._avg_dev.mul_(self.step_counter / divisor).add_(self._psi[0].to(torch.complex128), alpha=1.0 / divisor)
        self.step_counter += 1

    def get_profiles_df(self):
        import pandas as pd
        df_profile = pd.DataFrame(np.array(self.field_profile_history) / self.th_len)          # :113
        df_abs_profile = pd.DataFrame(np.array(self.field_abs_profile_history) / self.th_len)
        return (df_profile, df_abs_profile)

    def record_avgs(self):
        self.avg_amplitude_profile += np.abs(self.lattice).sum(axis=1) / self.th_len

    def plot_save(self, exp_dir, title):
        """CSV outputs of :573-586 (<title>_profile.csv, _snapshot.csv, _avglattice.csv); the PNG is not produced."""
        import pandas as pd
        pd.DataFrame(data=self.field_average).to_csv(os.path.join(exp_dir, title + "_profile.csv"))
        pd.DataFrame(data=self.lattice).to_csv(os.path.join(exp_dir, title + "_snapshot.csv"))
        pd.DataFrame(data=self.avg_lattice).to_csv(os.path.join(exp_dir, title + "_avglattice.csv"))


class LatticeRescaled0thOrder(Lattice):
    """On_lattice_rescaled_0thorder.Lattice: metric-rescaled proposal width, 0th-order background energy."""
    VARIANT = "rescaled_0thorder"

    def __init__(self, amplitude, wavenumber, radius, gamma, kappa, intrinsic_curvature, alpha, u, C, n, temperature,
                 temperature_final, dims=(100, 50), n_substeps=None, fieldsteps_per_ampstep=1, **kw):
        self.short_lengthscale_th = 2 * math.pi * 10 ** (-4)                                   # 0th:30
        self.area_factor = 2 / math.sqrt(math.pi)
        self.energy_const = 0
        self._temperature_for_background = temperature
        # O(u) correction to the background energy (0th:38-58); informational, not part of any move
        upper, lower = 1 * self.area_factor, (1 / max(dims)) * self.area_factor
        try:
            integral = math.pi / C * (math.log(C * upper ** 2 + alpha) - math.log(C * lower ** 2 + alpha))
        except ValueError:
            if C * upper ** 2 + alpha > 0:
                lower = math.sqrt(abs(alpha) / C) + .0000001
                integral = math.pi / C * (math.log(C * upper ** 2 + alpha) - math.log(C * lower ** 2 + alpha))
            else:
                integral = 0
        self.correction = u * temperature ** 3 * 4 * integral ** 2 / (dims[0] * dims[1])
        super().__init__(amplitude, wavenumber, radius, gamma, kappa, intrinsic_curvature, alpha, u, C, n, temperature,
                         temperature_final, dims, n_substeps, fieldsteps_per_ampstep, **kw)
        self.background_energy = np.full((self.z_len), self.get_background_energy((self.z_pixel_len, self.th_pixel_len)))

    def get_background_energy(self, cell_dims, ordered_reference=True):                        # 0th:66-92
        area = cell_dims[0] * cell_dims[1]
        ans = -self._temperature_for_background
        ordered = .5 * self.alpha ** 2 / self.u if (self.alpha < 0 and ordered_reference) else 0
        return ans / area + ordered

    def update_rescale_params(self, amplitude):                                                # 0th:94-104
        a = torch.full((self.z_len,), float(amplitude), dtype=torch.float64, device=self.device)
        z = torch.arange(self.z_len, dtype=torch.float64, device=self.device) * self.z_pixel_len
        g = ops.geometry_eval(self.wavenumber, self.radius, a, z).cpu().numpy()
        for z_index in range(self.z_len):
            cell_dims = (self.z_pixel_len * g[z_index, 3], self.th_pixel_len * g[z_index, 1])
            self.background_energy[z_index] = self.get_background_energy(cell_dims)


class LatticeRescaled1stOrder(LatticeRescaled0thOrder):
    """On_lattice_rescaled_1storder.Lattice: A_theta and the index raise at the interstitial rows, cross terms with an extra
    1/sqrt(g_theta), and an O(u) fluctuation correction in the background energy (1st:66-114).  The reference's constructor
    has no `fieldsteps_per_ampstep` (1st:20-21); it is accepted here and defaults to the base class's 1."""
    VARIANT = "rescaled_1storder"

    def get_background_energy(self, cell_dims, ordered_reference=True):                        # 1st:66-114
        area = cell_dims[0] * cell_dims[1]
        ans = -self._temperature_for_background
        upper_limit = 1 * self.area_factor
        lower_limit = (1 / max((self.z_len, self.th_len))) * self.area_factor
        alpha_cell = self.alpha * area
        hi, lo = self.C * upper_limit ** 2 + alpha_cell, self.C * lower_limit ** 2 + alpha_cell
        if hi > 0 and lo > 0:
            integral = math.pi / self.C * (math.log(hi) - math.log(lo))
        elif hi > 0:                                                                           # 1st:98-102
            lower_limit = math.sqrt(abs(alpha_cell) / self.C) + .0000001
            integral = math.pi / self.C * (math.log(hi) - math.log(self.C * lower_limit ** 2 + alpha_cell))
        else:
            integral = 0                                                                       # 1st:103-105
        correction = self.u * self._temperature_for_background ** 3 * 4 * integral ** 2
        correction /= self.z_len * self.th_len
        ans += correction
        ordered = .5 * self.alpha ** 2 / self.u if (self.alpha < 0 and ordered_reference) else 0
        return ans / area + ordered


class LatticeDecoupled(Lattice):
    """On_lattice_decoupled.Lattice: the field feels the metric, the shape only feels the C n^2 |A_theta psi|^2 coupling
    (dec:179-202).  Differences from `simple`: psi_0 = rect(U(0,.001), U(0,2pi)) (dec:157), initial sampling width 1e-4
    (dec:99), proposal width sigma*sqrt(g) (dec:274-275), bare z-gradient and cross terms (dec:294-324), and
    `surface_field_energy(amplitude)` takes the amplitude only (the other arguments of the base signature are accepted and
    ignored so that the engine wiring of the base class keeps working)."""
    VARIANT = "decoupled"
    INIT_MAGNITUDE = .001
    INIT_SAMPLING_WIDTH = .0001

    def surface_field_energy(self, amplitude, reference_amplitude=None, lattice=None, psi_squared=None, dz=None, dth=None):
        return super().surface_field_energy(amplitude, amplitude, lattice)
