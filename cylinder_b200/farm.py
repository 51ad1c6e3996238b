"""The reference's parameter-scan farm, file formats included, as ONE batched launch per GPU.

Reference: `scripts/taskarray.sh` writes `varfile.txt` with `parallel_preprocessing.py`, submits an SGE array
job (`cylinder_array.sh:10-15`) that runs `parallel_single_run.py --input infile.txt --varnames <v1> <v2> --varline <x1> <x2>` once
per line -- one OS process per grid point -- and finally joins the `*_mean.csv` files with `parallel_postprocessing.py`.

Here:
  preprocess(...)   = parallel_preprocessing.py:22-46      (same notes.txt / varfile.txt bytes)
  read_infile(...)  = the `@infile` parser of parallel_single_run.py:35-66 (same options, same defaults)
  run(...)          = every line of varfile.txt is one replica of a ReplicaBatch (field_type "lattice") or one chain of a
                      FourierChains batch (field_type "fourier"); under torchrun the lines are sharded over the ranks and every
                      rank writes the files of its own lines (the reference's processes do not communicate either).
                      Writes <var1>_<x1>_<var2>_<x2>_mean.csv (parallel_single_run.py:112-115) plus the profile CSVs of
                      run.py:341-346, with the strings of varfile.txt in the file names.
  postprocess(...)  = parallel_postprocessing.py:8-36      (one <column>.csv table per column of the mean files)

    python -m cylinder_b200.farm preprocess -n "notes" --var1name alpha --var1range -2 2 .25 --var2name C --var2range .25 4 .25
    torchrun --nproc-per-node 8 -m cylinder_b200.farm run --input infile.txt --varfile varfile.txt
    python -m cylinder_b200.farm postprocess
"""
import argparse
import glob
import os
from collections import defaultdict

import numpy as np

# names a varfile column may carry -> CylParams field (parallel_single_run.py:68-92 `exec`s the assignment, so any of the
# single_run arguments can be scanned; these are the ones that are numbers of the model)
_PARAM_OF = {"alpha": "alpha", "C": "C", "u": "u", "n": "n", "kappa": "kappa", "gamma": "gamma", "radius": "radius",
             "wavenumber": "wavenumber", "intrinsic_curvature": "intrinsic_curvature", "temp": "temperature"}


# ---- parallel_preprocessing.py ---------------------------------------------------------------------------
def preprocess(notes, var1name, var1range, var2name, var2range, outdir="."):
    """notes.txt and varfile.txt exactly as parallel_preprocessing.py:30-46 writes them: header `<v1> <v2>`, then one line
    per grid point, var1 slowest, values `str(round(x, 5))` of `np.arange(start, stop, step)`."""
    r1 = np.arange(var1range[0], var1range[1], var1range[2])
    r2 = np.arange(var2range[0], var2range[1], var2range[2])
    with open(os.path.join(outdir, "notes.txt"), "w") as f:
        f.write(notes)
    with open(os.path.join(outdir, "varfile.txt"), "w") as f:
        f.write(var1name + " " + var2name + "\n")
        for v1 in r1:
            for v2 in r2:
                f.write(str(round(v1, 5)) + " " + str(round(v2, 5)) + "\n")
    return len(r1) * len(r2)


def read_varfile(path):
    """-> ((var1name, var2name), [(x1_string, x2_string), ...]); the strings name the output files."""
    with open(path) as f:
        lines = [ln.split() for ln in f.read().splitlines() if ln.strip()]
    names = tuple(lines[0])
    if len(names) != 2 or any(len(ln) != 2 for ln in lines[1:]):
        raise ValueError("%s: expected a header of two names and two values per line" % path)
    return names, [tuple(ln) for ln in lines[1:]]


# ---- the @infile parser of parallel_single_run.py ---------------------------------------------------------
def _convert_arg_line_to_args(arg_line):
    for arg in arg_line.split():
        if not arg.strip():
            continue
        yield arg


def read_infile(path):
    """Namespace of the run options with the reference's defaults (parallel_single_run.py:35-66)."""
    p = argparse.ArgumentParser(fromfile_prefix_chars='@')
    p.convert_arg_line_to_args = _convert_arg_line_to_args
    p.add_argument('--temp', type=float)
    p.add_argument('--temp_final', type=float, required=False, default=None)
    p.add_argument('--n_steps', type=int)
    p.add_argument('--field_type', type=str)
    p.add_argument('--method', type=str)
    p.add_argument('--num_field_coeffs', required=False, nargs=2)
    p.add_argument('--fieldsteps_per_ampstep', required=False)
    p.add_argument('--measure_every', type=int)
    for name in ("alpha", "C", "n", "u", "gamma", "kappa", "radius", "intrinsic_curvature", "wavenumber", "amplitude"):
        p.add_argument('--' + name, type=float)
    p.add_argument('--dims', type=int, nargs=2)
    p.add_argument('--temperature_lattice', required=False, default=None, type=float)
    p.add_argument('--n_substeps', required=False, default=None, type=int)
    p.add_argument('--field_coeffs', required=False, default=None)
    p.add_argument('--outdir', required=False, default='.', type=str)
    a = p.parse_args(['@' + path])
    if a.temperature_lattice is None:
        a.temperature_lattice = 1
    if a.n_substeps is None and a.dims is not None:
        a.n_substeps = a.dims[0] * a.dims[1]
    if a.temp_final is None:
        a.temp_final = a.temp
    if a.num_field_coeffs is not None:
        a.num_field_coeffs = tuple(int(x) for x in a.num_field_coeffs)
    if a.fieldsteps_per_ampstep is not None:
        a.fieldsteps_per_ampstep = int(a.fieldsteps_per_ampstep)
    return a


def title_of(names, strings):
    """parallel_single_run.py:95."""
    return names[0] + "_" + strings[0] + "_" + names[1] + "_" + strings[1]


def _columns(args, names, lines):
    """Per-line model parameters: the infile's values, overridden by the two scanned columns."""
    n = len(lines)
    cols = {field: np.full(n, float(getattr(args, opt))) for opt, field in _PARAM_OF.items()}
    amp = np.full(n, float(args.amplitude if args.amplitude is not None else 0.0))
    for c, name in enumerate(names):
        vals = np.array([float(ln[c]) for ln in lines])
        if name == "amplitude":
            amp = vals
        elif name in _PARAM_OF:
            cols[_PARAM_OF[name]] = vals
        else:
            raise ValueError("cannot scan %r: not a model parameter (%s, amplitude)" % (name, ", ".join(sorted(_PARAM_OF))))
    return cols, amp


# ---- parallel_single_run.py for every line at once --------------------------------------------------------
def run(infile="infile.txt", varfile="varfile.txt", outdir=".", variant="rescaled_0thorder", seed=0, dtype=None, device=None,
        group=None, n_steps=None, write_profiles=True, write_series=False):
    """Run every line of `varfile` with the options of `infile`; returns {column: array over this rank's lines}.

    field_type "lattice": the adopted schedule of Lattice.run (On_lattice_simple.py:532-563): burn-in of
    500 x 50 proposals, then n_steps x (measurement, amplitude move, fieldsteps_per_ampstep x 50 proposals), proposals executed as
    whole checkerboard sweeps (>= 1 per step); means and their standard errors over the equilibrated part of every line's own
    series (stats.py: MSER cut-off on block means).
    field_type "fourier": the drivers of run.py:286-317 (sequential, simultaneous, fixed-amplitude, no-field) batched
    (FourierChains)."""
    import pandas as pd
    import torch
    import torch.distributed as dist
    from .replicas import OBS_NAMES, ReplicaBatch, shard_range
    args = read_infile(infile)
    names, lines = read_varfile(varfile)
    dist_on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if dist_on else 0
    world = dist.get_world_size(group) if dist_on else 1
    lo, hi = shard_range(len(lines), rank, world)
    mine = lines[lo:hi]
    steps = int(n_steps if n_steps is not None else args.n_steps)
    os.makedirs(outdir, exist_ok=True)
    if not mine:
        return {}
    cols, amp = _columns(args, names, mine)
    cols["amp_bound"] = .99                                                             # run.py:281 (|a| >= .99 rejected)
    out = {}
    profiles = None
    if args.field_type == "lattice":
        batch = ReplicaBatch(cols, dims=tuple(args.dims), variant=variant, dtype=dtype or torch.complex64, device=device, seed=seed,
                             replica0=lo, amplitude=amp)
        # proposals -> whole sweeps by the nominal lattice of the infile (dims[0] x dims[1]), not by this rank's share of the
        # lines: the schedule, hence every replica's result, must not depend on how the lines are sharded
        sweeps = lambda nprop: max(1, int(round(nprop / float(args.dims[0] * args.dims[1]))))          # noqa: E731
        per_step = sweeps((args.fieldsteps_per_ampstep or 1) * 50)                      # :533 n_sub_steps = 50
        batch.run(sweeps(500 * 50), amp_moves=False, measure=False)                     # :534-536
        # the steps are recorded as block means; every line's series is then cut at the end of its transient (MSER on <|a|>,
        # at most half of the run) -- me.save_equilibrium_stats / equilibrated_means of run.py:341-346
        nblocks = max(2, min(20, steps))
        per_block = -(-steps // nblocks)
        series = batch.run_blocks(nblocks, per_block, stride=per_step, series=write_series)
        means, errs, cut = batch.equilibrated_means(series)
        for k in OBS_NAMES:
            out[k] = means[k].cpu().numpy()
            out[k + "_err"] = errs[k].cpu().numpy()
        obs = batch.observables()
        for k in ("site_acceptance", "amplitude_acceptance"):
            out[k] = obs[k].cpu().numpy()
        out["energy"] = out["field_energy"] + out["surface_energy"]
        out["equilibration_cut_steps"] = cut.cpu().numpy() * per_block
        if write_profiles:
            profiles = (means["profile"].cpu().numpy(), means["profile_abs"].cpu().numpy(), batch.Z_host)
    elif args.field_type == "fourier":
        from .fourier_chains import METHODS, FourierChains
        if args.method not in METHODS:
            raise ValueError("farm: method must be one of %s (got %r)" % (sorted(METHODS), args.method))
        ch = FourierChains(cols, num_field_coeffs=args.num_field_coeffs, amplitude=amp, device=device, seed=seed, chain0=lo)
        ch.run(steps, measure_every=args.measure_every or 1, fieldsteps_per_ampstep=args.fieldsteps_per_ampstep or 1,
               method=args.method)
        for k, v in ch.summary().items():
            out[k] = v.cpu().numpy()
        mean = ch.mean.cpu().numpy()
        nz, nth = ch.nz, ch.nth
        idx = [(a, b) for a in range(-nz, nz + 1) for b in range(-nth, nth + 1)]       # run.py:266-267
        for c, (a, b) in enumerate(idx):
            out["abs_coeff_%d_%d" % (a, b)] = mean[:, 1 + c]
    else:
        raise ValueError("farm: field_type must be 'lattice' or 'fourier' (got %r)" % args.field_type)
    frames = None
    if args.field_type == "lattice" and write_series and batch.block_series is not None:
        frames = batch.time_series_frames(batch.block_series)                           # <title>.csv of run.py:324-326
    for r, strings in enumerate(mine):
        title = title_of(names, strings)
        if frames is not None:
            frames[r].to_csv(os.path.join(outdir, title + ".csv"))
        d = dict([(key, [out[key][r]]) for key in out])                                 # parallel_single_run.py:113
        pd.DataFrame.from_dict(d).to_csv(os.path.join(outdir, title + "_mean.csv"))
        if profiles is not None:
            Z = int(profiles[2][r])
            pd.Series(profiles[0][r, :Z]).to_csv(os.path.join(outdir, title + "_profile.csv"))         # run.py:343-346
            pd.Series(profiles[1][r, :Z]).to_csv(os.path.join(outdir, title + "_profile_abs.csv"))
    return out


# ---- parallel_postprocessing.py ---------------------------------------------------------------------------
def postprocess(outdir="."):
    """Join the *_mean.csv files into one <column>.csv table per column (rows: var2 strings, columns: var1 strings), as
    parallel_postprocessing.py:8-36 does, including its way of taking the values from the file names."""
    import pandas as pd
    filenames = sorted(glob.glob(os.path.join(outdir, "*_mean.csv")))
    if not filenames:
        raise FileNotFoundError("no *_mean.csv in %s" % outdir)
    colnames = list(pd.read_csv(filenames[0], index_col=0, header=0).columns)
    data = dict([(c, defaultdict(lambda: defaultdict())) for c in colnames])
    for filename in filenames:
        filedata = pd.read_csv(filename, index_col=0, header=0)
        vs = os.path.basename(filename).split("_")
        filevar1, filevar2 = vs[1], vs[-2]                                              # :25-26
        for c in colnames:
            data[c][filevar1][filevar2] = filedata[c][0]
    for c in colnames:
        pd.DataFrame(data[c]).to_csv(os.path.join(outdir, c + ".csv"))
    return colnames


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m cylinder_b200.farm", description=__doc__.split("\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    pp = sub.add_parser("preprocess")
    pp.add_argument('-n', '--notes', dest='notes', type=str, required=True)
    pp.add_argument('--var1name', type=str, required=True)
    pp.add_argument('--var2name', type=str, required=True)
    pp.add_argument('--var1range', type=float, nargs=3, required=True)
    pp.add_argument('--var2range', type=float, nargs=3, required=True)
    pp.add_argument('--outdir', default=".")
    pr = sub.add_parser("run")
    pr.add_argument('-i', '--input', default="infile.txt", dest="infile")
    pr.add_argument('--varfile', default="varfile.txt")
    pr.add_argument('--outdir', default=".")
    pr.add_argument('--variant', default="rescaled_0thorder")
    pr.add_argument('--seed', type=int, default=0)
    pr.add_argument('--series', action="store_true", help="also write every line's time series <title>.csv")
    pq = sub.add_parser("postprocess")
    pq.add_argument('--outdir', default=".")
    a = ap.parse_args(argv)
    if a.cmd == "preprocess":
        n = preprocess(a.notes, a.var1name, a.var1range, a.var2name, a.var2range, a.outdir)
        print("varfile.txt: %d grid points" % n)
    elif a.cmd == "run":
        import torch
        import torch.distributed as dist
        if "RANK" in os.environ and not dist.is_initialized():
            local = int(os.environ.get("LOCAL_RANK", 0))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        run(a.infile, a.varfile, a.outdir, variant=a.variant, seed=a.seed, write_series=a.series)
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    else:
        print("wrote", ", ".join(c + ".csv" for c in postprocess(a.outdir)))


if __name__ == "__main__":
    main()
