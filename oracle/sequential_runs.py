"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package `cylinder_b200`).

Whole runs of the reference's SEQUENTIAL sampler for the ensemble tests: Z*T proposals per sweep at uniformly random sites
(Lattice.step_lattice, On_lattice_simple.py:593-596), per-proposal Robbins-Monro (:688-696), one measurement per sweep
(Lattice.measure / measure_avgs, :121-156) and optionally one engine-style amplitude move per sweep (run_lattice.py:72-90).
Every step is the pinned restatement of oracle/lattice_oracle.py; this module only loops over it, one chain per process
(picklable entry point for multiprocessing, spawn context: the parent may hold a CUDA context)."""
import random

import numpy as np

from . import checkerboard_oracle as cb
from . import lattice_oracle as lo


def run_chain(args):
    """args = dict(psi0 complex [Z,T], P dict of Cylinder/Lattice parameters, variant, amplitude, sigma, step_counter,
    nsweeps, burn, seed, amp_moves=False, sigma_amp=.005, amp_counter=0, adapt_amp=True).
    Returns a dict of time averages over the measured sweeps: psi2, abs_psi, acceptance, profile_abs [Z], profile [Z] complex,
    abs_a, amp_acceptance, final sigma."""
    P = args["P"]
    psi = np.array(args["psi0"], dtype=complex)
    Z, T = psi.shape
    st = lo.LatticeState(psi, P["wavenumber"], P["radius"], P["alpha"], P["u"], P["C"], P["n"], P["temperature"],
                         variant=args.get("variant", lo.SIMPLE))
    st.sigma = float(args["sigma"])
    st.step_counter = int(args.get("step_counter", 0))
    rng = random.Random(args["seed"])
    amp_moves = bool(args.get("amp_moves", False))
    am = cb.AmpState(amplitude=float(args["amplitude"]), sigma_amp=float(args.get("sigma_amp", .005)),
                     target=P.get("amp_target_acceptance", .3), bound=P.get("amp_bound", .8))
    am.counter = int(args.get("amp_counter", 0))
    nsw, burn = int(args["nsweeps"]), int(args.get("burn", 0))
    acc0 = prop0 = 0
    n = 0
    s_psi2 = s_abs = s_absa = 0.0
    prof_abs = np.zeros(Z)
    prof = np.zeros(Z, dtype=complex)
    for s in range(burn + nsw):
        if s == burn:
            acc0, prop0 = st.accepted, st.step_counter
            am.proposed = am.accepted = 0
        for _ in range(Z * T):
            lo.step_lattice_loc(st, am.a, rng.randrange(-1, Z - 1), rng.randrange(-1, T - 1), rng.gauss(0, 1), rng.gauss(0, 1),
                                rng.random)
        if s >= burn:
            n += 1
            s_psi2 += float(np.mean(st.psi2))
            ps, pa = lo.measure_profiles(st.psi)
            s_abs += float(pa.sum()) / (Z * T)
            s_absa += abs(am.a)
            prof_abs += pa / T
            prof += ps / T
        if amp_moves:
            cb.amplitude_move(st, am, P, args["seed"], 0x7FFFFFFF, s, adapt=bool(args.get("adapt_amp", True)))
    return dict(psi2=s_psi2 / n, abs_psi=s_abs / n, acceptance=(st.accepted - acc0) / max(1, st.step_counter - prop0),
                profile_abs=prof_abs / n, profile=prof / n, abs_a=s_absa / n,
                amp_acceptance=am.accepted / max(1, am.proposed), sigma=st.sigma, amplitude=am.a)


def run_chains(tasks, processes=None):
    """Map run_chain over tasks in a spawn-context pool (one chain per host core)."""
    import multiprocessing as mp
    import os
    processes = processes or min(len(tasks), os.cpu_count() or 1)
    with mp.get_context("spawn").Pool(processes) as pool:
        return pool.map(run_chain, tasks)
