"""Batched replicas: the (alpha, C, n, kappa, k, ...) parameter grid that the reference farms out as one OS process
per grid point (scripts/taskarray.sh -> scripts/cylinder_array.sh:10-15 -> parallel_single_run.py) runs here as
ONE launch: replica b of the batch = task b of the SGE array.  Replicas never interact, so across GPUs they shard
by contiguous index range, one process per GPU, with a single NCCL all-gather of the observable records standing
in for the reference's `glob("*_mean.csv")` join (parallel_postprocessing.py:8-29).
"""
import numpy as np
import torch

from . import _abi, ops
from ._abi import SWEEP_ADAPT_SIGMA, SWEEP_AMP_MOVES, SWEEP_MEASURE

OBS_NAMES = ("abs_amplitude", "amplitude", "amplitude_squared", "psi_squared", "abs_psi", "field_energy", "surface_energy",
             "field_energy_squared", "sampling_width")


def shard_range(n_total, rank, world):
    """Contiguous replica index range [lo, hi) owned by `rank` (balanced to within one replica)."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pipeline_slices(n_total, nchunks):
    """Slice boundaries [(lo, hi)] of a host round trip cut into `nchunks` slices: the first and last slices are a quarter, their
    neighbours half, of the ones in the middle.  The pipeline's fill is the upload of the first slice (nothing can run before it
    has landed) and its drain the download of the last one, so those are kept short; the middle slices stay large enough for a
    launch to fill the GPU."""
    nchunks = max(1, min(int(nchunks), int(n_total)))
    if nchunks < 5:
        return [shard_range(n_total, c, nchunks) for c in range(nchunks)]
    w = [4] * nchunks
    w[0] = w[-1] = 1
    w[1] = w[-2] = 2
    tot, acc, cuts = sum(w), 0, [0]
    for x in w[:-1]:
        acc += x
        cuts.append(max(cuts[-1], min(n_total, (n_total * acc + tot // 2) // tot)))
    cuts.append(n_total)
    return [(cuts[c], cuts[c + 1]) for c in range(nchunks)]


def parameter_grid(**axes):
    """Cartesian product of named 1-D axes -> dict of flat arrays (first axis slowest), like the reference's
    varfile.txt written by parallel_preprocessing.py:22-46."""
    names = list(axes)
    grids = np.meshgrid(*[np.asarray(axes[n], dtype=np.float64) for n in names], indexing="ij")
    return {n: g.ravel() for n, g in zip(names, grids)}


class ReplicaBatch:
    """B independent Lattice chains on one GPU, lattices resident in shared memory during a launch.

    params: dict name -> scalar or array[B] with the reference's constructor names (wavenumber, radius, gamma, kappa,
    intrinsic_curvature, alpha, u, C, n, temperature) plus amp_bound / amp_target_acceptance.
    `replica0` is the global index of local replica 0 (the Philox stream of a replica depends only on its global
    index and the seed, so results do not depend on how the batch is sharded)."""

    def __init__(self, params, dims=(50, 50), variant="simple", dtype=torch.complex64, device=None, seed=0, replica0=0,
                 amplitude=0.0, sigma_site=.005, sigma_amp=.005, init_magnitude=.1, initialize=True):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        _abi.lib()                                           # fail loudly if the CUDA library is missing
        self.params = ops.make_params(self.device, **params)
        self.B = self.params.shape[0]
        k = self.params[:, 0].cpu().numpy()
        zs = np.array([ops.lattice_dims(dims, kk)[0] for kk in k], dtype=np.int32)      # On_lattice_simple.py:41
        self.T = int(dims[1])
        self.Zmax = int(zs.max())
        self.Z_host = zs
        self.Z = torch.from_numpy(zs).to(self.device)
        self.variant = ops.variant_id(variant)
        self.dtype = dtype
        self.seed, self.replica0 = int(seed), int(replica0)
        self.psi = torch.zeros(self.B, self.Zmax, self.T, dtype=dtype, device=self.device)
        if initialize:
            ops.lattice_init(self.psi, self.Z, init_magnitude, self.seed, self.replica0)
        self.chain = ops.new_chain(self.B, self.device, amplitude, sigma_site, sigma_amp)
        self.obs, self.prof = ops.new_obs(self.B, self.Zmax, self.device)
        self.sweeps_done = 0

    # ---- hot path --------------------------------------------------------------------------------------
    @property
    def sites(self):
        """Site-updates per sweep summed over the batch."""
        return int(self.Z_host.astype(np.int64).sum()) * self.T

    def run(self, nsweeps, amp_moves=True, measure=True, adapt=True, series=False, index_order=False):
        """nsweeps sweeps of every replica in one launch.  series=True (with measure): also returns the per-sweep records
        [B, nsweeps, 8] (ops.SERIES_NAMES: amplitude, energies, <|psi|^2>, <|psi|>, width, accept flags) -- the engine's time
        series (me.df, run.py:324-326) for every replica at once."""
        flags = (SWEEP_ADAPT_SIGMA if adapt else 0) | (SWEEP_AMP_MOVES if amp_moves else 0) | (SWEEP_MEASURE if measure else 0)
        if index_order:                      # CTA b works on replica b (default: longest replicas first; same results)
            flags |= _abi.SWEEP_INDEX_ORDER
        ts = torch.zeros(self.B, int(nsweeps), len(ops.SERIES_NAMES), dtype=torch.float64, device=self.device) if (series and measure) else None
        ops.sweep_smem(self.psi, self.Z, self.params, self.chain, self.seed, self.sweeps_done, nsweeps, self.variant, flags,
                       self.replica0, self.obs, self.prof, series=ts)
        self.sweeps_done += int(nsweeps)
        return ts if ts is not None else self

    def time_series_frames(self, series, sweep0=0):
        """One pandas DataFrame per replica (columns ops.SERIES_NAMES + energy, index = sweep) from a run(..., series=True)
        result -- what me.save_time_series() leaves in me.df for a single run."""
        import pandas as pd
        arr = series.cpu().numpy()
        frames = []
        for b in range(arr.shape[0]):
            df = pd.DataFrame(arr[b], columns=list(ops.SERIES_NAMES), index=pd.RangeIndex(sweep0, sweep0 + arr.shape[1], name="sweep"))
            df["energy"] = df["field_energy"] + df["surface_energy"]
            frames.append(df)
        return frames

    def reset_observables(self):
        self.obs.zero_()
        self.prof.zero_()

    def run_blocks(self, nblocks, sweeps_per_block, amp_moves=True, adapt=True, stride=1, series=False):
        """nblocks blocks of sweeps_per_block measured sweeps (each preceded by stride-1 plain sweeps: the reference's
        fieldsteps_per_ampstep); returns the block means as a dict name -> tensor [nblocks, B] (device) -- the time series
        the equilibrium cut-off works on (stats.equilibrated) -- and keeps the per-block profile means in
        `self.block_profiles` [nblocks, B, Zmax, 3]."""
        blocks = {name: [] for name in OBS_NAMES}
        profs, records = [], []
        for _ in range(int(nblocks)):
            o0, p0 = self.obs.clone(), self.prof.clone()
            if stride <= 1:
                r = self.run(sweeps_per_block, amp_moves=amp_moves, measure=True, adapt=adapt, series=series)
                if series:
                    records.append(r)
            else:
                for _ in range(int(sweeps_per_block)):
                    self.run(stride - 1, amp_moves=False, measure=False, adapt=adapt)
                    r = self.run(1, amp_moves=amp_moves, measure=True, adapt=adapt, series=series)
                    if series:
                        records.append(r)
            d = self.obs - o0
            n = d[:, _abi.OB_COUNT].clamp(min=1.0)
            for i, name in enumerate(OBS_NAMES):
                blocks[name].append(d[:, 1 + i] / n)
            profs.append((self.prof - p0) / (n[:, None, None] * self.T))
        self.block_profiles = torch.stack(profs)
        self.block_series = torch.cat(records, dim=1) if records else None      # [B, measurements, 8] (series=True)
        return {name: torch.stack(v) for name, v in blocks.items()}

    def equilibrated_means(self, series, by="abs_amplitude", max_fraction=.5):
        """Means over the equilibrated part of a run_blocks series: every replica's series is cut where the MSER rule puts
        the end of the transient of `by`.  Returns (means, standard errors, cut index per replica); profiles are averaged
        over the same blocks."""
        from . import stats
        cut = stats.mser_cut(series[by], max_fraction)
        means, errs = {}, {}
        for name, x in series.items():
            means[name], errs[name] = stats.truncated_mean(x, cut)
        if getattr(self, "block_profiles", None) is not None and self.block_profiles.shape[0] == series[by].shape[0]:
            c = cut.reshape(-1, 1, 1)
            p, _ = stats.truncated_mean(self.block_profiles, c.expand(-1, self.Zmax, 3))
            means["profile"] = torch.complex(p[..., 0], p[..., 1])
            means["profile_abs"] = p[..., 2]
        return means, errs, cut

    # ---- observables -----------------------------------------------------------------------------------
    def observables(self):
        """Per-replica means over the measurements so far: dict name -> tensor[B] (device)."""
        n = self.obs[:, _abi.OB_COUNT].clamp(min=1.0)
        out = {name: self.obs[:, 1 + i] / n for i, name in enumerate(OBS_NAMES)}
        ch = self.chain
        out["site_acceptance"] = ch[:, _abi.CH_SITE_ACCEPTED] / ch[:, _abi.CH_SITE_PROPOSED].clamp(min=1.0)
        out["amplitude_acceptance"] = ch[:, _abi.CH_AMP_ACCEPTED] / ch[:, _abi.CH_AMP_PROPOSED].clamp(min=1.0)
        out["n_measurements"] = self.obs[:, _abi.OB_COUNT]
        return out

    def profiles(self):
        """(<Psi>(z), <|Psi|>(z)) per replica, theta- and time-averaged: Lattice.get_profiles_df (On_lattice_simple.py:109-118)."""
        n = self.obs[:, _abi.OB_COUNT].clamp(min=1.0)[:, None] * self.T
        return torch.complex(self.prof[..., 0], self.prof[..., 1]) / n, self.prof[..., 2] / n

    def record(self):
        """Fixed-size per-replica record [B, 16 + 16 + 3*Zmax] (obs | chain | profiles) -- the gather payload."""
        return torch.cat([self.obs, self.chain, self.prof.reshape(self.B, -1)], dim=1).contiguous()

    def gather_records(self, group=None):
        """All ranks' records, concatenated in replica order, on every rank (fp64 [sum B_r, 32 + 3 Zmax]).  Blocking form:
        gather_start() + gather_finish()."""
        self.gather_start(group)
        return self.gather_finish()

    def gather_start(self, group=None):
        """Start the exchange of the per-replica records and return at once.  The records are snapshotted on the compute
        stream (one small copy), the all-gather runs on a side stream into buffers allocated once -- the shapes were exchanged
        when the exchange was built -- so the next launch of sweeps overlaps it: the gather is never on the critical path
        (SURVEY.md 8e).  Profiles travel as fp32 (they are sums of O(10^4) terms of O(1): 1e-7 relative is far below their
        statistical error); obs | chain stay fp64."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            self._gather_local = True
            return self
        self._gather_local = False
        if getattr(self, "_exchange", None) is None or self._exchange.group is not group:
            self._exchange = RecordExchange(self.B, self.obs.shape[1] + self.chain.shape[1], 3 * self.Zmax, self.device, group)
        self._exchange.start(torch.cat([self.obs, self.chain], dim=1), self.prof.reshape(self.B, -1))
        return self

    def gather_finish(self):
        """Wait (on the current stream, not the host) for the exchange started by gather_start; returns the records."""
        if self._gather_local:
            return self.record()
        r64, r32 = self._exchange.finish()
        return torch.cat([r64, r32.to(torch.float64)], dim=1)

    # ---- host round trips ------------------------------------------------------------------------------
    @property
    def lattice(self):
        """numpy [B, Zmax, T] complex (rows >= Z[b] are zero)."""
        return self.psi.cpu().numpy()

    def load_host(self, psi_host, chain_host=None, non_blocking=True):
        """Upload lattices (and chain state) from host tensors (pinned for async copies)."""
        self.psi.copy_(psi_host, non_blocking=non_blocking)
        if chain_host is not None:
            self.chain.copy_(chain_host, non_blocking=non_blocking)

    def step_host(self, psi_host, chain_host, psi_out_host, rec_out_host, nsweeps, nchunks=16, amp_moves=True, measure=True,
                  adapt=True, wait=True, lattices=True):
        """One step with HOST buffers in and out (pinned tensors): upload lattices + chain state, run nsweeps, download
        lattices + per-replica records.  The batch is cut into `nchunks` slices (pipeline_slices: short first and last slices),
        each on its own stream, so the H2D copy of one slice, the sweeps of another and the D2H copy of a third overlap (PCIe is
        full duplex; replicas are independent).  Returns after everything has landed in the host buffers (wait=True).  lattices=False keeps the lattices resident
        on the device (chain state in, records out): what a production run does between checkpoints -- on a host with several GPUs
        the pinned copies of the lattices are what limits the step (profiles/r02_e2e_limits.md)."""
        flags = (SWEEP_ADAPT_SIGMA if adapt else 0) | (SWEEP_AMP_MOVES if amp_moves else 0) | (SWEEP_MEASURE if measure else 0)
        if not hasattr(self, "_streams") or len(self._streams) != nchunks:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(nchunks)]
        main = torch.cuda.current_stream(self.device)
        bounds = pipeline_slices(self.B, nchunks)
        for st, (b0, b1) in zip(self._streams, bounds):
            if b1 <= b0:
                continue
            st.wait_stream(main)
            with torch.cuda.stream(st):
                if lattices:
                    self.psi[b0:b1].copy_(psi_host[b0:b1], non_blocking=True)
                self.chain[b0:b1].copy_(chain_host[b0:b1], non_blocking=True)
                ops.sweep_smem(self.psi[b0:b1], self.Z[b0:b1], self.params[b0:b1], self.chain[b0:b1], self.seed, self.sweeps_done,
                               nsweeps, self.variant, flags, self.replica0 + b0, self.obs[b0:b1], self.prof[b0:b1])
                if lattices:
                    psi_out_host[b0:b1].copy_(self.psi[b0:b1], non_blocking=True)
                rec = torch.cat([self.obs[b0:b1], self.chain[b0:b1], self.prof[b0:b1].reshape(b1 - b0, -1)], dim=1)
                rec_out_host[b0:b1].copy_(rec, non_blocking=True)
        for st in self._streams:
            main.wait_stream(st)
        self.sweeps_done += int(nsweeps)
        if wait:                                 # wait=False: the caller synchronises before touching the host buffers
            main.synchronize()

    def state_dict(self):
        return dict(psi=self.psi.cpu(), chain=self.chain.cpu(), obs=self.obs.cpu(), prof=self.prof.cpu(), seed=self.seed,
                    replica0=self.replica0, sweeps_done=self.sweeps_done, variant=self.variant)

    def load_state_dict(self, sd):
        """Exact resume: the counter-based RNG has no state beyond (seed, replica, sweep index)."""
        self.psi.copy_(sd["psi"])
        self.chain.copy_(sd["chain"])
        self.obs.copy_(sd["obs"])
        self.prof.copy_(sd["prof"])
        self.seed, self.replica0, self.sweeps_done = int(sd["seed"]), int(sd["replica0"]), int(sd["sweeps_done"])


class RecordExchange:
    """All-gather of fixed-size per-replica records, [n_r, w64] float64 + [n_r, w32] float32 per rank, built for repetition:
    the ranks exchange their (n_r, w64, w32) ONCE here, the padded send / receive buffers are allocated once, and every
    start() is a snapshot copy on the caller's stream plus two all_gather_into_tensor calls on a side stream (CUDA; on CPU
    tensors -- the gloo tests -- it runs inline).  finish() makes the caller's stream wait and returns the records of all
    ranks concatenated in rank (= replica) order, trimmed to each rank's own shape."""

    def __init__(self, n_local, w64, w32, device, group=None):
        import torch.distributed as dist
        self.group, self.device = group, torch.device(device)
        self.world = dist.get_world_size(group)
        mine = torch.tensor([n_local, w64, w32], dtype=torch.int64, device=self.device)
        shapes = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(shapes, mine, group=group)
        self.shapes = [tuple(int(v) for v in sh.tolist()) for sh in shapes]          # the only host sync, once
        self.n, self.w64, self.w32 = n_local, w64, w32
        nmax = max(sh[0] for sh in self.shapes)
        m64, m32 = max(sh[1] for sh in self.shapes), max(sh[2] for sh in self.shapes)
        # two sets of send buffers, used alternately: the snapshot of one step never waits for the previous step's gather
        self.send64 = [torch.zeros(nmax, m64, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.send32 = [torch.zeros(nmax, m32, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.turn = 0
        self.recv64 = torch.empty(self.world, nmax, m64, dtype=torch.float64, device=self.device)
        self.recv32 = torch.empty(self.world, nmax, m32, dtype=torch.float32, device=self.device)
        self.cuda = self.device.type == "cuda"
        # high priority: the gather's few CTAs are placed as soon as CTAs of the concurrently running sweeps retire (a normal-
        # priority kernel launched behind a grid of 8192 CTAs only starts when that grid has been dispatched completely)
        self.stream = torch.cuda.Stream(device=self.device, priority=-1) if self.cuda else None
        self.done = [None, None]                             # per send-buffer set: event "its last gather has read it"
        self.pending = False

    def start(self, rec64, rec32):
        import torch.distributed as dist
        t = self.turn
        self.turn ^= 1
        s64, s32 = self.send64[t], self.send32[t]
        if self.cuda:
            main = torch.cuda.current_stream(self.device)
            if self.done[t] is not None:
                main.wait_event(self.done[t])                # the gather that last read this set (two steps ago)
        s64[: self.n, : self.w64].copy_(rec64)
        s32[: self.n, : self.w32].copy_(rec32)
        if self.cuda:
            self.stream.wait_stream(main)
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.recv64.view(-1), s64.view(-1), group=self.group)
                dist.all_gather_into_tensor(self.recv32.view(-1), s32.view(-1), group=self.group)
                self.done[t] = torch.cuda.Event()
                self.done[t].record(self.stream)
        else:
            dist.all_gather(list(self.recv64.unbind(0)), s64, group=self.group)
            dist.all_gather(list(self.recv32.unbind(0)), s32, group=self.group)
        self.pending = True

    def finish(self):
        assert self.pending, "RecordExchange.finish() without start()"
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self.pending = False
        m32 = self.recv32.shape[2]
        r64 = torch.cat([self.recv64[r, :sh[0]] for r, sh in enumerate(self.shapes)], dim=0)
        r32 = torch.cat([self.recv32[r, :sh[0]] for r, sh in enumerate(self.shapes)], dim=0)
        return r64, r32


def gather_ragged(rec, group=None):
    """all_gather of [B_r, W] tensors whose B_r (and W) may differ by rank: pad to the max, gather, trim."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    shape = torch.tensor(list(rec.shape), dtype=torch.int64, device=rec.device)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    bmax = max(int(s[0]) for s in shapes)
    wmax = max(int(s[1]) for s in shapes)
    pad = torch.zeros(bmax, wmax, dtype=rec.dtype, device=rec.device)
    pad[: rec.shape[0], : rec.shape[1]] = rec
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: int(s[0])] for o, s in zip(outs, shapes)], dim=0)
