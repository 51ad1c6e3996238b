#!/usr/bin/env python
"""bench.py -- Metropolis site-updates/s of the cylinder_b200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...      # the reference algorithm on the host CPUs
    torchrun --nproc-per-node N bench.py --gpus N ...                # one rank per GPU (weak scaling)

One "step" = SWEEPS_PER_STEP full sweeps (site moves + Robbins-Monro + measurement + one global amplitude move per
sweep) of the per-GPU share of config C5: 8192 independent replicas over C3's (alpha, C, k) grid x 2 seeds, 50 x 50
base lattice (Z = round(50/k) in [40, 100], T = 50), fp32 psi, lattices resident in shared memory.
Legs, all in one run and one JSON line:
  value      device-resident state, CUDA events around K steps, max over ranks
  e2e        same steps through ReplicaBatch with HOST (pinned) lattice + chain buffers: H2D of the state, the sweeps,
             D2H of lattice + observable records, every step
  roofline   config C4: one 4096 x 4096 fp32 lattice streamed from HBM by the TMA-tiled kernel, timed alone with
             CUDA events; achieved = 24 B x site-updates / time against MEASURED_PEAKS.json hbm_gbs
  roofline_value  the kernel the headline times (sweep_smem_kernel<float>): warp-instructions of one step (ncu
             smsp__inst_executed.sum of exactly this launch, kept with the hash of the kernel sources in
             profiles/r02_counters.json) / step time, against SMs x 4 issue slots x the SM clock sampled during the run
  sustained  the value leg repeated for >= --sustained-seconds with one CUDA event per step: rate of the first and of the
             last second (clock droop under sustained load would show here)
  cpu_baseline  the oracle's sequential restatement of the reference's per-site Python loop, one independent replica
             per host core for a bounded wall time (rank 0, N = 1 only)
  aux        N = 1 only: the same full path in fp64 (the reference's precision), config C1 through Lattice.run, config C3
             (4096 replicas x 1000 sweeps), config C2 (Fourier model)
--scaling strong: config C5's 65536 replicas in total, split over the ranks (default weak: 8192 per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "metropolis_site_updates_per_s"
UNIT = "site-updates/s"
REPLICAS_PER_GPU = 8192
SWEEPS_PER_STEP = 200
DIMS = (50, 50)
SEED = 0x5EED
WORKLOAD = ("C5 per-GPU share: %d replicas/GPU over C3's 16x16x16 (alpha, C, k) grid x seeds, dims (50,50) => Z=round(50/k) in "
            "[40,100] (odd Z: 3 colours), T=50")
PER_SWEEP = "site moves + Robbins-Monro + row moments/measure + 1 amplitude move"
CPU_SAMPLE = ("%d replicas of the same grid, one per host core, whole sweeps of the reference algorithm (Z*T random-site proposals "
              "with per-proposal Robbins-Monro + 1 amplitude move + 1 measurement per sweep)")
ALG_BYTES_PER_UPDATE_F32 = 24.0      # 2 reads + 1 write of an 8-byte site per full sweep
COUNTERS = os.path.join(ROOT, "profiles", "r02_counters.json")   # written by scripts/ncu_counters.py from the .ncu-rep files
STRONG_TOTAL = 65536


def profiler_counters():
    """ncu counters of the two hot kernels with the hash of the kernel sources they were measured on.  Returns
    (counters or None, note): counters of another source state are refused (the leg then reports no fraction)."""
    from cylinder_b200.build import source_sha
    try:
        with open(COUNTERS) as f:
            c = json.load(f)
    except (OSError, ValueError):
        return None, "profiles/r02_counters.json missing"
    sha = source_sha()
    if c.get("source_sha") != sha:
        return None, "profiles/r02_counters.json was measured on kernel sources %s, this build is %s: re-profile" % (c.get("source_sha"), sha)
    return c, "profiles/r02_counters.json (kernel sources %s)" % sha


def grid_params(n_replicas, replica0=0):
    """C3's grid: alpha in {-2..1.75} (16) x C in {.25..4} (16) x k in linspace(.5, 1.25, 16) = 4096 points; replica r
    uses grid point r % 4096 (the seed copies of C5 differ only in their Philox replica index)."""
    import numpy as np
    alphas = -2.0 + 0.25 * np.arange(16)
    Cs = 0.25 + 0.25 * np.arange(16)
    ks = np.linspace(0.5, 1.25, 16)
    idx = (replica0 + np.arange(n_replicas)) % 4096
    ia, ic, ik = idx // 256, (idx // 16) % 16, idx % 16
    return dict(wavenumber=ks[ik], radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=alphas[ia], u=1.0, C=Cs[ic],
                n=6.0, temperature=0.01, amp_bound=0.8, amp_target_acceptance=0.3)


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (sequential random-site Metropolis, per-proposal Robbins-Monro) as restated by
# the oracle -- the only place bench.py executes oracle/ code, and only as the thing timed for the CPU baseline.
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One host core = one replica of the SAME workload as the GPU arm (a point of C3's grid, Z = round(50/k) rows), advanced
    by whole sweeps of the reference algorithm: Z*T proposals at uniformly random sites with per-proposal Robbins-Monro
    (Lattice.step_lattice), one global amplitude move and one measurement per sweep (run_lattice.py:72-90)."""
    idx, seconds, warm_sweeps = args
    import random
    import numpy as np
    from oracle import checkerboard_oracle as cb
    from oracle import lattice_oracle as lo
    g = grid_params(1, (idx * 257) % 4096)                                  # spread the sample over the grid
    P = {k: float(np.atleast_1d(v)[0]) for k, v in g.items()}
    Z, T = lo.lattice_dims(DIMS, P["wavenumber"])
    psi = cb.random_initialize(Z, T, .1, SEED, idx)
    st = lo.LatticeState(psi, P["wavenumber"], P["radius"], P["alpha"], P["u"], P["C"], P["n"], P["temperature"])
    am = cb.AmpState(amplitude=0.0, sigma_amp=.005, target=P["amp_target_acceptance"], bound=P["amp_bound"])
    rng = random.Random(12345 + idx)

    def sweep(s):
        for _ in range(Z * T):
            lo.step_lattice_loc(st, am.a, rng.randrange(-1, Z - 1), rng.randrange(-1, T - 1), rng.gauss(0, 1), rng.gauss(0, 1),
                                rng.random)
        lo.measure_profiles(st.psi)
        cb.amplitude_move(st, am, P, SEED, idx, s)
    for s in range(warm_sweeps):
        sweep(s)
    n_done, s, t0 = 0, warm_sweeps, time.perf_counter()
    while True:
        sweep(s)
        s += 1
        n_done += Z * T
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n_done, dt


def cpu_reference_rate(seconds, cores=None, warm=1):
    """Aggregate site-updates/s of `cores` independent replicas (what scripts/taskarray.sh does with one process per
    grid point).  Returns (rate, cores, total updates)."""
    import multiprocessing as mp
    cores = cores or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(i, seconds, warm) for i in range(cores)])
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return total / wall, cores, total


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_rate(min(per_step, 2.0))
    rates, cores, total = [], 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, cores, n = cpu_reference_rate(per_step)
        rates.append(r)
        total += n
    wall = time.perf_counter() - t0
    value = sum(rates) / len(rates)
    sample = CPU_SAMPLE % cores + ", %.1f s per step" % per_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (STRONG_TOTAL // max(1, args.gpus) if args.scaling == "strong" else args.replicas_per_gpu), "per_sweep": PER_SWEEP,
                       "sample": "reference algorithm on the host cores: " + sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows]
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--replicas-per-gpu", type=int, default=REPLICAS_PER_GPU)
    ap.add_argument("--sweeps-per-step", type=int, default=SWEEPS_PER_STEP)
    ap.add_argument("--hbm-size", type=int, default=4096)
    ap.add_argument("--hbm-sweeps", type=int, default=200)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-hbm", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--sustained-seconds", type=float, default=10.0)
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from cylinder_b200 import _abi, ops
    from cylinder_b200.replicas import ReplicaBatch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own stream at high priority: the record gather runs beside the next step's sweeps (a grid of 8192 CTAs), and a
        # normal-priority kernel behind such a grid is only dispatched when the grid has been dispatched completely
        os.environ.setdefault("NCCL_MAX_CTAS", "4")     # the gather moves ~12 MB per rank per step: a few CTAs are plenty, and
        opts = dist.ProcessGroupNCCL.Options()          # every CTA it holds is one the sweeps running beside it do not get
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    _abi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from cylinder_b200.replicas import shard_range
    nsw, K, W = args.sweeps_per_step, args.steps, max(args.warmup, 0)
    if args.scaling == "strong":                                   # config C5 in total, whatever the number of GPUs
        replica0, hi = shard_range(STRONG_TOTAL, rank, world)
        Bg = hi - replica0
    else:
        Bg = args.replicas_per_gpu
        replica0 = rank * Bg
    batch = ReplicaBatch(grid_params(Bg, replica0), dims=DIMS, variant="simple", dtype=torch.complex64, device=dev, seed=SEED,
                         replica0=replica0, amplitude=0.0)
    sites_per_sweep = batch.sites
    rec_buf = None

    def step_device():
        batch.run(nsw, amp_moves=True, measure=True, adapt=True)
        if world > 1:                      # the path's one exchange: observable records of every replica on every rank --
            batch.gather_start()           # snapshot + all-gather on a side stream, overlapped by the next step's sweeps

    # ---- value leg ----
    sampler = ClockSampler(local) if rank == 0 else None      # started before warm-up so that samples exist from the first timed ms
    for _ in range(W):
        step_device()
    barrier()
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_device()
    if world > 1:
        batch.gather_finish()              # the last exchange is inside the timed region
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    # ---- sustained leg: the same steps for >= S seconds, one event per step ----
    sustained = None
    if not args.no_sustained and args.sustained_seconds > 0:
        n_sus = max(K, int(args.sustained_seconds * 1e3 / (ms_total / K)) + 1)
        if world > 1:
            t = torch.tensor([n_sus], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_sus = int(t.item())
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_sus + 1)]
        evs[0].record()
        for i in range(n_sus):
            step_device()
            evs[i + 1].record()
        barrier()
        t_wall1 = time.perf_counter()
        ts = [evs[0].elapsed_time(e) for e in evs[1:]]                     # ms since the start, per step
        per_step_local = float(sites_per_sweep) * nsw

        def window_rate(lo_ms, hi_ms):
            idx = [i for i, t in enumerate(ts) if lo_ms < t <= hi_ms]
            if not idx:
                return None
            t_first = ts[idx[0] - 1] if idx[0] > 0 else 0.0
            return per_step_local * len(idx) / ((ts[idx[-1]] - t_first) * 1e-3)
        total_ms = max_over_ranks(ts[-1])
        sustained = {"seconds": total_ms * 1e-3, "steps": n_sus, "rate_first_second_this_rank": window_rate(0.0, 1000.0),
                     "rate_last_second_this_rank": window_rate(ts[-1] - 1000.0, ts[-1]),
                     "min_step_ms": min(b - a for a, b in zip([0.0] + ts[:-1], ts)), "max_step_ms": max(b - a for a, b in zip([0.0] + ts[:-1], ts))}
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    updates_per_step_all = float(sites_per_sweep) * nsw
    if world > 1:
        t = torch.tensor([updates_per_step_all], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        updates_per_step_all = float(t.item())
    value = updates_per_step_all * K / (ms_total * 1e-3)
    if sustained is not None:
        sustained["rate"] = updates_per_step_all * sustained["steps"] / sustained["seconds"]
        sustained["unit"] = UNIT
        sustained["vs_value"] = sustained["rate"] / value
    acc = float(batch.observables()["site_acceptance"].mean().item())

    # ---- roofline of the kernel the headline times: instruction issue ----
    counters, counters_note = profiler_counters()
    roofline_value = None
    if rank == 0:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
        kc = (counters or {}).get("sweep_smem_kernel<float,extras>")
        same = kc is not None and kc.get("replicas") == Bg and kc.get("sweeps") == nsw and args.scaling == "weak"
        roofline_value = {"bound": "sm_issue", "unit": "warp-instructions/s", "kernel": "sweep_smem_kernel<float, extras>",
                          "peak": (sms * 4 * sm_mhz * 1e6) if sm_mhz else None,
                          "peak_source": "%d SMs x 4 issue slots x %.0f MHz (median SM clock sampled during the timed region)" % (sms, sm_mhz or 0),
                          "counters": counters_note}
        if same and sm_mhz:
            inst = float(kc["smsp__inst_executed.sum"])
            roofline_value.update({"achieved": inst / (ms_total / K * 1e-3), "frac": inst / (ms_total / K * 1e-3) / roofline_value["peak"],
                                   "warp_instructions_per_step": inst,
                                   "thread_instruction_slots_per_site_update": inst * 32.0 / (float(sites_per_sweep) * nsw),
                                   "ncu": {k: kc.get(k) for k in ("gpu__time_duration_ms", "issue_active_pct", "top_stalls", "site_loop_sass_instructions")},
                                   "note": "issue slots are not all usable by this mix: on this GPU FFMA with three register sources issues "
                                           "every 1.9 cycles per sub-partition, FFMA2 every 3.0, IMAD.WIDE / LOP3 / IMAD every 2.0-2.4, MUFU every 8 "
                                           "(scripts/micro/pipes.cu, profiles/r02_pipes.md)"})
        else:
            roofline_value.update({"achieved": None, "frac": None, "why": counters_note if kc is None else "counters are for %s replicas x %s sweeps, weak" % (kc.get("replicas"), kc.get("sweeps"))})

    # ---- e2e leg: host buffers in, host buffers out, every step ----
    e2e = None
    if not args.no_e2e:
        psi_h = batch.psi.cpu().pin_memory()
        chain_h = batch.chain.cpu().pin_memory()
        rec = batch.record()
        rec_h = torch.empty(rec.shape, dtype=rec.dtype).pin_memory()
        psi_out_h = torch.empty_like(psi_h).pin_memory()

        def step_e2e():
            # H2D: lattices + chain state; the sweeps; D2H: lattices (what `.lattice` hands to the caller) + records;
            # pipelined over 4 slices of the batch, returns when the host buffers are complete
            batch.step_host(psi_h, chain_h, psi_out_h, rec_h, nsw, nchunks=args.e2e_chunks)
            if world > 1:
                batch.gather_start()
        for _ in range(min(W, 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        if world > 1:
            batch.gather_finish()
        barrier()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
        h2d = psi_h.numel() * psi_h.element_size() + chain_h.numel() * 8
        d2h = rec_h.numel() * 8 + psi_out_h.numel() * psi_out_h.element_size()
        # the same with the lattices resident on the device (chain state in, records out): a production run between checkpoints
        def step_rec():
            batch.step_host(psi_h, chain_h, psi_out_h, rec_h, nsw, nchunks=args.e2e_chunks, lattices=False)
            if world > 1:
                batch.gather_start()
        step_rec()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_rec()
        if world > 1:
            batch.gather_finish()
        barrier()
        ms_rec = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": updates_per_step_all * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / K,
               "api": "ReplicaBatch.step_host: pinned host lattices + chain in, sweeps, lattices + records out, %d-slice stream pipeline, host-synchronised per step" % args.e2e_chunks,
               "lattices_resident": {"value": updates_per_step_all * K / (ms_rec * 1e-3), "ms_per_step": ms_rec / K,
                                     "h2d_bytes_per_step": int(chain_h.numel() * 8), "d2h_bytes_per_step": int(rec_h.numel() * 8),
                                     "note": "step_host(lattices=False): chain state in, records out, lattices stay on the device (not the headline e2e)"}}

    # ---- roofline leg: C4, one large lattice streamed from HBM ----
    roofline = None
    if not args.no_hbm:
        n = args.hbm_size
        par = ops.make_params(dev, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
        Zt = torch.tensor([n], dtype=torch.int32, device=dev)
        psi = torch.zeros(1, n, n, dtype=torch.complex64, device=dev)
        ops.lattice_init(psi, Zt, .1, SEED, 0)
        ws = ops.HbmWorkspace(n, n, torch.complex64, dev)
        ws.pack(psi[0])
        del psi
        chain = ops.new_chain(1, dev, amplitude=0.3)[0]
        ws.sweep(par, chain, SEED, 0, 10, 0, _abi.SWEEP_ADAPT_SIGMA)          # warm-up (also adapts sigma)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        ws.sweep(par, chain, SEED, 10, args.hbm_sweeps, 0, _abi.SWEEP_ADAPT_SIGMA)
        h1.record()
        barrier()
        ms_hbm = h0.elapsed_time(h1)
        # the same lattice with measurement, profiles and one amplitude move per sweep ("and with amplitude moves")
        obs_h = torch.zeros(16, dtype=torch.float64, device=dev)
        prof_h = torch.zeros(n, 3, dtype=torch.float64, device=dev)
        full = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
        nfull = max(1, args.hbm_sweeps // 4)
        ws.sweep(par, chain, SEED, 10 + args.hbm_sweeps, 4, 0, full, obs=obs_h, prof=prof_h)
        barrier()
        h0.record()
        ws.sweep(par, chain, SEED, 14 + args.hbm_sweeps, nfull, 0, full, obs=obs_h, prof=prof_h)
        h1.record()
        barrier()
        ms_full = h0.elapsed_time(h1)
        peak, peak_src = measured_peak_hbm()
        upd = float(n) * n * args.hbm_sweeps
        achieved = upd * ALG_BYTES_PER_UPDATE_F32 / (ms_hbm * 1e-3) / 1e9
        kh = (counters or {}).get("sweep_hbm_kernel<float,2,32,false>")
        traffic = (kh["dram__bytes_read.sum"] + kh["dram__bytes_write.sum"]) if (kh and kh.get("n") == n) else None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, " + counters_note + " "
                                      "(below the 24 B/update algorithmic figure: the kernel updates all colours per pass, 8.1 B read + "
                                      "8 B written per site; part of the writes is still dirty in the 126 MB L2 when the kernel ends)",
                    "algorithmic_bytes_per_launch": ALG_BYTES_PER_UPDATE_F32 * n * n,
                    "with_measurement_and_amplitude_move": {"us_per_sweep": 1e3 * ms_full / nfull, "sweeps": nfull,
                                                            "site_updates_per_s": float(n) * n * nfull / (ms_full * 1e-3),
                                                            "frac": float(n) * n * nfull * ALG_BYTES_PER_UPDATE_F32 / (ms_full * 1e-3) / 1e9 / peak},
                    "peak_source": peak_src, "kernel": "sweep_hbm_kernel<float,2,TH,false> (TMA-tiled, ping-pong)",
                    "workload": "C4: %dx%d fp32 lattice (%.0f MB > L2), a=0.3 fixed, %d sweeps" % (n, n, n * n * 8 / 1e6, args.hbm_sweeps),
                    "algorithmic_bytes_per_update": ALG_BYTES_PER_UPDATE_F32, "site_updates_per_s": upd / (ms_hbm * 1e-3),
                    "us_per_sweep": 1e3 * ms_hbm / args.hbm_sweeps,
                    "note": "time = whole call / sweeps: one sweep kernel per sweep (it carries the Robbins-Monro record itself) + one coefficient and one finalize kernel per call"}
        del ws

    # ---- config C2 (Fourier-mode model + adaptive engine), batched: reported beside the headline, not part of it ----
    aux = None
    if rank == 0 and world == 1 and not args.no_aux:
        import numpy as np
        from cylinder_b200.fourier_chains import FourierChains
        nch, nst = 4096, 40
        fc = FourierChains(dict(wavenumber=np.full(nch, .5), radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0,
                                C=1.0, n=6.0, temperature=.1, amp_bound=.99), num_field_coeffs=(2, 1), amplitude=0.0, seed=1, device=dev)
        fc.run(10, measure_every=10, fieldsteps_per_ampstep=10)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        fc.run(nst, measure_every=10, fieldsteps_per_ampstep=10)
        f1.record()
        torch.cuda.synchronize()
        moves = nch * nst * 10 * 11
        aux = {"fourier_c2": {"engine_moves_per_s": moves / (f0.elapsed_time(f1) * 1e-3), "chains": nch,
                              "config": "Cylinder2D (nz, nth) = (2, 1): 15 complex coefficients + amplitude, method sequential, "
                                        "measure_every 10, fieldsteps_per_ampstep 10, T = .1; one warp per chain",
                              "reference": "0.24 ms per calc_field_energy, 47 ms per calc_field_energy_amplitude_change on one "
                                           "CPU core (SURVEY.md section 6)"}}
        del fc

        def timed(fn):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e-3
        # the headline's full per-sweep path in the reference's own precision (complex128 psi, fp64 arithmetic throughout)
        B64, nsw64 = 2368, 100
        b64 = ReplicaBatch(grid_params(B64, 0), dims=DIMS, variant="simple", dtype=torch.complex128, device=dev, seed=SEED, amplitude=0.0)
        b64.run(20)
        t64 = timed(lambda: b64.run(nsw64))
        aux["fp64"] = {"site_updates_per_s": b64.sites * nsw64 / t64, "replicas": B64, "sweeps": nsw64, "dtype": "f64",
                       "config": "same grid and per-sweep path as the headline, psi complex128 (the reference's precision); 2 CTAs/SM"}
        del b64
        # config C3 as SURVEY 8d states it: 4096 replicas (one per grid point) x 1000 sweeps
        b3 = ReplicaBatch(grid_params(4096, 0), dims=DIMS, variant="simple", dtype=torch.complex64, device=dev, seed=SEED, amplitude=0.0)
        b3.run(50)
        t3 = timed(lambda: b3.run(1000))
        aux["c3"] = {"site_updates_per_s": b3.sites * 1000 / t3, "replicas": 4096, "sweeps": 1000, "seconds": t3, "dtype": "f32"}
        del b3
        # config C1: the reference's own single-replica case through the drop-in class (one 50 x 50 lattice = ONE CTA of the GPU)
        import random
        import cylinder_b200
        random.seed(12345)
        lat = cylinder_b200.Lattice(amplitude=0, wavenumber=1, radius=1, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1, C=1, n=6,
                                    temperature=.01, temperature_final=.01, dims=(50, 50), fieldsteps_per_ampstep=50)
        lat.run(10, 50)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        n_steps_c1, sps_c1 = 2000, 5
        lat.run(n_steps_c1, 50, sweeps_per_step=sps_c1)
        torch.cuda.synchronize()
        w1 = time.perf_counter() - w0
        aux["c1"] = {"seconds": w1, "sweeps": n_steps_c1 * sps_c1, "site_updates_per_s": n_steps_c1 * sps_c1 * 2500 / w1,
                     "config": "Lattice(dims=(50,50), T=.01, alpha=-1, u=1, C=1, n=6).run(2000 steps of 5 sweeps + 1 amplitude move + "
                               "measurement = 1e4 sweeps), wall clock incl. the class's own burn-in sweeps and per-step host round trips",
                     "reference": "~20 min for the same number of proposals at ~2e4 proposals/s (SURVEY.md section 6)"}
        del lat

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, cores, total = cpu_reference_rate(args.cpu_seconds)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "per_core": rate / max(1, cores), "kind": "port",
               "sample": CPU_SAMPLE % cores + ", %.0f s wall, %d proposals: oracle restatement of the reference's sequential "
                         "per-site Python loop" % (args.cpu_seconds, total)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD % Bg + ", fp32 psi resident in shared memory",
                           "sweeps_per_step": nsw, "per_sweep": PER_SWEEP,
                           "site_updates_per_step": updates_per_step_all, "variant": "simple", "temperature": 0.01,
                           "l2_policy": "state re-read from HBM once per step (%.0f MB/GPU > L2 not required: lattices live in "
                                        "shared memory during a step)" % (Bg * batch.Zmax * batch.T * 8 / 1e6),
                           "mean_site_acceptance": acc},
                "clocks": clocks, "e2e": e2e, "gpu_launches": K, "roofline": roofline, "roofline_value": roofline_value,
                "sustained": sustained, "cpu_baseline": cpu, "aux": aux}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
