"""Drop-in mirror of the reference driver's `single_run` (run.py:182-353), the one working entry point of the farm
(parallel_single_run.py:101-110 -> run.single_run).  Same arguments, same return tuple
(params_names, observables_names, result_means, covariance_matrix_complex) and the same output files
(<title>.csv time series, <title>_profile.csv, <title>_profile_abs.csv, <title>_snapshot.csv, <title>_avglattice.csv,
last_sigma.pickle, last_cov.pickle); plots are not produced.  `grid_run` is the batched form of the whole farm
(scripts/taskarray.sh): one launch instead of one OS process per grid point.
"""
import os
import pickle

import numpy as np
import torch

from . import system_cylinder as cylinder
from . import system_cylinder2D as cylinder2d
from .lattice import LatticeRescaled0thOrder
from .metropolisengine import MetropolisEngine

me_version = "cylinder_b200"

# module-level defaults of the reference driver (run.py:356-380): used when a value is not a loop variable
alpha = -1
C = 1
u = 1
n = 6
kappa = 0
gamma = 1
temp = .1
intrinsic_curvature = 0
initial_amplitude = 0
radius = 1
wavenumber = .5
num_field_coeffs = (2, 1)
measure_every = 10
fieldsteps_per_ampstep = 10
dims = (100, 50)
n_steps = 1000
field_type = "lattice"
method = "sequential"
notes = ""
EXPERIMENT_TYPES = (("num_field_coeffs", "fieldsteps_per_ampstep"), ("amplitude", "C"), ("wavenumber", "gamma"), ("wavenumber", "kappa"),
                    ("wavenumber", "alpha"), ("wavenumber", "intrinsic_curvature"), ("alpha", "C"))            # run.py:60-99


def single_run(temp, temp_final, n_steps, field_type, method, num_field_coeffs, fieldsteps_per_ampstep, measure_every,
               alpha, C, n, u, gamma, kappa, radius, intrinsic_curvature, n_substeps, dims, wavenumber,
               amplitude=None, field_coeffs=None, outdir=None, title=None, lattice_class=LatticeRescaled0thOrder,
               dtype=torch.complex128):
    if amplitude is None:
        amplitude = initial_amplitude
    num_field_coeffs = tuple(int(x) for x in num_field_coeffs) if num_field_coeffs is not None else (0, 0)   # Appendix C
    fieldsteps_per_ampstep, n = int(fieldsteps_per_ampstep), int(n) if float(n).is_integer() else n
    # warm start (run.py:208-228)
    sampling_width = .05
    if os.path.isfile("./last_sigma.pickle") and os.path.getsize("./last_sigma.pickle"):
        try:
            with open("last_sigma.pickle", "rb") as f:
                sampling_width = pickle.load(f)
        except EOFError:
            sampling_width = .05
    cov = None
    if field_type == "fourier" and method != "no-field":
        ncoef = (2 * num_field_coeffs[0] + 1) * (2 * num_field_coeffs[1] + 1)
        if field_coeffs is None:
            field_coeffs = np.array([MetropolisEngine.gaussian_complex(sigma=.1) for _ in range(ncoef)])
        if os.path.isfile("./last_cov.pickle") and os.path.getsize("./last_cov.pickle"):
            with open("last_cov.pickle", "rb") as f:
                cov = pickle.load(f)
            if cov is not None and len(cov) != ncoef:
                cov = None

    lattice = se = None
    if field_type == "fourier":
        if method == "no-field":
            se = cylinder.Cylinder(wavenumber=wavenumber, radius=radius, kappa=kappa, gamma=gamma,
                                   intrinsic_curvature=intrinsic_curvature)
            surf = lambda real_params, complex_params: se.calc_surface_energy(*real_params)               # noqa: E731
            me = MetropolisEngine(energy_functions={"real": {"surface": surf}, "all": {"surface": surf}},
                                  initial_complex_params=None, initial_real_params=[float(amplitude)],
                                  covariance_matrix_complex=None, sampling_width=sampling_width, temp=temp,
                                  complex_sample_method=None)
        else:
            se = cylinder2d.Cylinder2D(wavenumber=wavenumber, radius=radius, alpha=alpha, C=C, u=u, n=n, kappa=kappa,
                                       gamma=gamma, intrinsic_curvature=intrinsic_curvature, num_field_coeffs=num_field_coeffs)
            zl, tl = num_field_coeffs[0] * 2 + 1, num_field_coeffs[1] * 2 + 1
            field = lambda real_params, complex_params: se.calc_field_energy(np.reshape(complex_params, (tl, zl)))      # noqa: E731
            field_alt = lambda real_params, complex_params: se.calc_field_energy_amplitude_change(                        # noqa: E731
                *real_params, np.reshape(complex_params, (tl, zl)))
            surf = lambda real_params, complex_params: se.calc_surface_energy(*real_params)                             # noqa: E731
            groups = {"complex": {"field": field}, "real": {"field": field_alt, "surface": surf},
                      "all": {"field": field_alt, "surface": surf}}
            me = MetropolisEngine(energy_functions=groups, initial_complex_params=field_coeffs,
                                  initial_real_params=[float(amplitude)], covariance_matrix_complex=cov,
                                  sampling_width=sampling_width, temp=temp, complex_sample_method="magnitude-phase")
    elif field_type == "lattice":
        lattice = lattice_class(amplitude=amplitude, wavenumber=wavenumber, radius=radius, gamma=gamma, kappa=kappa,
                                intrinsic_curvature=intrinsic_curvature, alpha=alpha, u=u, C=C, n=n, temperature=temp,
                                temperature_final=temp_final, dims=dims, n_substeps=n_substeps,
                                fieldsteps_per_ampstep=fieldsteps_per_ampstep, dtype=dtype)
        me = lattice.me
    else:
        raise ValueError("choose field_type from ('lattice', 'fourier'); received %r" % (field_type,))

    params_names = ["amplitude"]
    if field_type == "fourier" and method != "no-field":
        idx = [(a, b) for a in range(-num_field_coeffs[0], num_field_coeffs[0] + 1)
               for b in range(-num_field_coeffs[1], num_field_coeffs[1] + 1)]
        params_names.extend(["coeff_" + str(a) + "_" + str(b) for a, b in idx])
        se.save_temporary_matrices()
    me.params_names = params_names
    observables_names = ["abs_" + name for name in params_names]
    observables_names.extend(["amplitude_squared"])
    me.observables_names = observables_names
    me.set_reject_condition(lambda real_params, complex_params: abs(real_params[0]) >= .99)           # run.py:281

    if field_type == "fourier":
        if method == "sequential":
            for i in range(n_steps):
                for j in range(measure_every):
                    if me.step_real_group():
                        se.save_temporary_matrices()
                    for ii in range(fieldsteps_per_ampstep):
                        me.step_complex_group()
                me.measure()
        elif method == "simultaneous":
            for i in range(n_steps):
                for j in range(measure_every):
                    if me.step_all():
                        se.save_temporary_matrices()
                me.measure()
        elif method == "fixed-amplitude":
            me.m -= 1
            for i in range(n_steps):
                for j in range(measure_every):
                    for ii in range(fieldsteps_per_ampstep):
                        me.step_complex_group()
                me.measure()
        elif method == "no-field":
            me.m = 1
            for i in range(n_steps):
                for j in range(measure_every):
                    me.step_real_group()
                me.measure()
    else:
        if method == "fixed-amplitude":
            lattice.run_fixed_amplitude(n_steps, n_substeps)
        else:
            lattice.run(n_steps, n_substeps)
        if outdir is not None and os.path.isdir(outdir):
            lattice.plot_save(outdir, title)

    if outdir is not None and os.path.isdir(outdir):
        me.save_time_series()
        me.df.to_csv(os.path.join(outdir, title + ".csv"))
    with open("last_sigma.pickle", "wb") as f:
        if field_type == "fourier" and method == "sequential":
            pickle.dump([me.real_group_sampling_width, me.complex_group_sampling_width], f)
        else:
            pickle.dump(me.real_group_sampling_width, f)
    if field_type == "fourier" and method == "sequential":
        with open("last_cov.pickle", "wb") as f:
            pickle.dump(me.covariance_matrix_complex, f)
    import pandas as pd
    profiles = lattice.get_profiles_df() if lattice is not None else None
    me.save_equilibrium_stats(profiles)
    result_means = me.equilibrated_means
    if outdir is not None and os.path.isdir(outdir) and lattice is not None:
        pd.Series(me.field_profile).to_csv(os.path.join(outdir, title + "_profile.csv"))
        pd.Series(me.field_abs_profile).to_csv(os.path.join(outdir, title + "_profile_abs.csv"))
    return me.params_names, me.observables_names, result_means, me.covariance_matrix_complex


def grid_run(var1, values1, var2, values2, n_sweeps, base, dims=(50, 50), variant="rescaled_0thorder", burn_in=None,
             dtype=torch.complex64, seed=0, outdir=None, device=None, group=None):
    """The whole farm of scripts/taskarray.sh in one launch: every (var1, var2) grid point is one replica of a
    ReplicaBatch sharded over the ranks of `group` (if torch.distributed is initialised).  Writes, per grid point,
    <var1>_<v1>_<var2>_<v2>_mean.csv as parallel_single_run.py:113-115 does, so parallel_postprocessing.py keeps working.
    Returns the gathered per-replica means as a dict of numpy arrays (on every rank)."""
    import torch.distributed as dist
    from .replicas import OBS_NAMES, ReplicaBatch, parameter_grid, shard_range
    grid = parameter_grid(**{var1: values1, var2: values2})
    ntot = len(grid[var1])
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(ntot, rank, world)
    params = dict(base)
    for k in (var1, var2):
        params[k] = grid[k][lo:hi]
    batch = ReplicaBatch(params, dims=dims, variant=variant, dtype=dtype, device=device, seed=seed, replica0=lo)
    batch.run(burn_in if burn_in is not None else max(1, n_sweeps // 5), measure=False)
    batch.run(n_sweeps)
    rec = batch.gather_records(group).cpu().numpy()
    counts = np.maximum(rec[:, 0], 1.0)
    means = {name: rec[:, 1 + i] / counts for i, name in enumerate(OBS_NAMES)}
    means[var1], means[var2] = grid[var1], grid[var2]
    if outdir is not None and rank == 0:
        import pandas as pd
        os.makedirs(outdir, exist_ok=True)
        for r in range(ntot):
            row = {k: [v[r]] for k, v in means.items()}
            pd.DataFrame(row).to_csv(os.path.join(outdir, "%s_%s_%s_%s_mean.csv" % (var1, grid[var1][r], var2, grid[var2][r])))
    return means


# ------------------------------------------------------------------------------------------------------
# the 2-D scan driver of run.py:21-179 (loop / plot_save / run_experiment)
# ------------------------------------------------------------------------------------------------------
def loop(single_run_lambda, range1, range2):
    """run.py:21-58: call single_run_lambda(var1, var2) over the grid and collect, per result name, a 2-D list
    [len(range1)][len(range2)]: `<name>_mean`, `time_per_experiment`, `<coeff>_variance`, `<c1>_<c2>_covariance`."""
    import collections
    import timeit
    results = collections.defaultdict(list)
    for var1 in range1:
        results_line = collections.defaultdict(list)
        for var2 in range2:
            start_time = timeit.default_timer()
            param_names, observables_names, means_dict, cov_matrix = single_run_lambda(var1, var2)
            elapsed = timeit.default_timer() - start_time
            coeffs_names = param_names[1:]
            if cov_matrix is None:
                coeffs_names = []
            for name in means_dict:
                results_line[name + "_mean"].append(means_dict[name])
            results_line["time_per_experiment"].append(elapsed)
            for i, name in enumerate(coeffs_names):
                results_line[name + "_variance"].append(cov_matrix[i][i])
                for j, name2 in enumerate(coeffs_names):
                    if j > i:
                        results_line[name + "_" + name2 + "_covariance"].append(cov_matrix[i][j])
        for name in results_line:
            results[name].append(results_line[name])
    return results


def plot_save(range1, range2, results, title, exp_dir='.'):
    """run.py:103-130 without the heat map: <title>.csv, rows = range1, columns = range2 (ranges cut to the data's shape)."""
    import pandas as pd
    range1 = list(range1)[-len(results):]
    range2 = list(range2)[-len(results[0]):]
    pd.DataFrame(index=range1, columns=range2, data=results).to_csv(os.path.join(exp_dir, title + ".csv"))


def run_experiment(exp_type, range1, range2, batched=True, exp_dir=None, seed=0, variant="rescaled_0thorder", **overrides):
    """run.py:133-179: a 2-D grid of simulations over the two variables named by `exp_type`, results as one CSV table per
    observable in out/exp-<timestamp>/ plus notes.csv.  Values that are not scanned come from the module-level defaults
    (as in the reference) or from `overrides`.  With batched=True (lattice field, or Fourier field with a fixed number of
    coefficients) the whole grid is ONE ReplicaBatch / FourierChains launch; otherwise the reference's serial loop over
    single_run."""
    import datetime
    import pandas as pd
    exp_type = tuple(exp_type)
    if exp_type not in EXPERIMENT_TYPES:
        raise KeyError("experiment type not found: %r; set exp_type to one of %r" % (exp_type, EXPERIMENT_TYPES))
    g = dict(globals())
    g.update(overrides)
    range1, range2 = list(range1), list(range2)
    now = datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S")
    if exp_dir is None:
        exp_dir = os.path.join("out", "exp-" + now)
    os.makedirs(exp_dir, exist_ok=True)
    exp_notes = {"experiment type": " ".join(exp_type), "n_steps": g["n_steps"], "me_temp": g["temp"], "C": g["C"], "kappa": g["kappa"],
                 "gamma": g["gamma"], "alpha": g["alpha"], "n": g["n"], "u": g["u"], "range1": range1,
                 "intrinsic curvature": g["intrinsic_curvature"], "range2": range2, "radius": g["radius"],
                 "amplitude": g["initial_amplitude"], "wavenumber": g["wavenumber"], "measure every n ampsteps": g["measure_every"],
                 "total number ampsteps": g["n_steps"] * g["measure_every"], "notes": g["notes"], "start_time": now,
                 "me-version": me_version, "field_type": g["field_type"], "method": g["method"], "dims": g["dims"],
                 "num_field_coeffs": g["num_field_coeffs"], "fieldsteps per ampstep": g["fieldsteps_per_ampstep"]}
    pd.DataFrame.from_dict({k: str(v) for k, v in exp_notes.items()}, orient="index", columns=["value"]).to_csv(
        os.path.join(exp_dir, "notes.csv"))
    names = {"amplitude": "amplitude", "C": "C", "wavenumber": "wavenumber", "gamma": "gamma", "kappa": "kappa", "alpha": "alpha",
             "intrinsic_curvature": "intrinsic_curvature"}
    if not batched or exp_type == EXPERIMENT_TYPES[0]:
        def one(v1, v2):
            kw = dict(temp=g["temp"], temp_final=g["temp"], n_steps=g["n_steps"], field_type=g["field_type"], method=g["method"],
                      num_field_coeffs=g["num_field_coeffs"], fieldsteps_per_ampstep=g["fieldsteps_per_ampstep"],
                      measure_every=g["measure_every"], alpha=g["alpha"], C=g["C"], n=g["n"], u=g["u"], gamma=g["gamma"],
                      kappa=g["kappa"], radius=g["radius"], intrinsic_curvature=g["intrinsic_curvature"],
                      n_substeps=g["dims"][0] * g["dims"][1], dims=g["dims"], wavenumber=g["wavenumber"],
                      amplitude=g["initial_amplitude"], outdir=exp_dir, title="%s%s_%s%s" % (exp_type[0], round(v1, 5), exp_type[1], round(v2, 5)))
            for name, v in zip(exp_type, (v1, v2)):
                kw[name] = v
            return single_run(**kw)
        results = loop(one, range1, range2)
    else:
        from .replicas import OBS_NAMES, ReplicaBatch, parameter_grid
        grid = parameter_grid(**{exp_type[0]: range1, exp_type[1]: range2})
        base = dict(wavenumber=g["wavenumber"], radius=g["radius"], gamma=g["gamma"], kappa=g["kappa"],
                    intrinsic_curvature=g["intrinsic_curvature"], alpha=g["alpha"], u=g["u"], C=g["C"], n=g["n"], temperature=g["temp"],
                    amp_bound=.99)
        amp0 = g["initial_amplitude"]
        for name in exp_type:
            if name == "amplitude":
                amp0 = grid[name]
            else:
                base[names[name]] = grid[name]
        if g["field_type"] == "lattice":
            batch = ReplicaBatch(base, dims=tuple(g["dims"]), variant=variant, seed=seed, amplitude=amp0)
            per = max(1, int(round(g["fieldsteps_per_ampstep"] * 50 / float(g["dims"][0] * g["dims"][1]))))
            batch.run(max(1, int(round(500 * 50 / float(g["dims"][0] * g["dims"][1])))), amp_moves=False, measure=False)
            nb = max(2, min(20, g["n_steps"]))
            series = batch.run_blocks(nb, -(-g["n_steps"] // nb), stride=per)
            means, errs, cut = batch.equilibrated_means(series)
            flat = {k + "_mean": means[k].cpu().numpy() for k in OBS_NAMES}
            flat.update({k + "_err": errs[k].cpu().numpy() for k in OBS_NAMES})
        else:
            from .fourier_chains import FourierChains
            ch = FourierChains(base, num_field_coeffs=g["num_field_coeffs"], amplitude=amp0, seed=seed)
            ch.run(g["n_steps"], measure_every=g["measure_every"], fieldsteps_per_ampstep=g["fieldsteps_per_ampstep"], method=g["method"])
            flat = {k + "_mean": v.cpu().numpy() for k, v in ch.summary().items()}
        shape = (len(range1), len(range2))
        results = {k: np.asarray(v).reshape(shape).tolist() for k, v in flat.items()}
    for name in results:
        plot_save(range1=range1, range2=range2, results=results[name], title=name, exp_dir=exp_dir)
    return results, exp_dir
