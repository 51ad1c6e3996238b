"""Parity of the batched shared-memory kernel ON THE SHAPES THE BENCHMARK TIMES (config C3 / C5: T = 50, Z = round(50/k)
in [40, 100], even and odd Z, amplitude move + measurement every sweep, T = 0.01), through the C-ABI, against the numpy
oracle on the same Philox streams.  At T = 50 a thread of the column-mapped site loops walks 4-10 rows per colour
(k_sweep_smem.cu: row groups), odd Z takes the 3-colour mapped branch, and the energy pass strides rows with the periodic
wrap upwards -- none of which the small-lattice cases of test_gpu_parity.py reach.  Reference semantics:
On_lattice_simple.py:598-686 (site move), :185-213 (field energy), :121-131 (profiles)."""
import numpy as np
import pytest

from oracle import checkerboard_oracle as cb
from oracle import lattice_oracle as lo

pytestmark = pytest.mark.gpu

ALPHAS = -2.0 + 0.25 * np.arange(16)          # C3's grid (bench.py:grid_params)
CS = 0.25 + 0.25 * np.arange(16)
KS = np.linspace(0.5, 1.25, 16)
BENCH_T = 50


def _k_for_rows(Z):
    """A wavenumber with round(50 / k) == Z (the benchmark's own k values where they exist)."""
    for k in KS:
        if int(round(50 / k)) == Z:
            return float(k)
    return 50.0 / Z


def _oracle_run(psi0, P, variant, a0, sigma_site, sigma_amp, seed, rep, sweep0, nsw, amp_moves=True):
    """nsw sweeps of the oracle in the kernel's order: colour passes, measurement, amplitude move.  Returns the state, the
    amplitude record and the accumulated observables."""
    st = lo.LatticeState(psi0.copy(), P["wavenumber"], P["radius"], P["alpha"], P["u"], P["C"], P["n"], P["temperature"], variant=variant)
    st.sigma = sigma_site
    am = cb.AmpState(amplitude=a0, sigma_amp=sigma_amp, target=P.get("amp_target_acceptance", .3), bound=P.get("amp_bound", .8))
    Z = psi0.shape[0]
    acc = dict(abs_a=0.0, e_field=0.0, e_surf=0.0, psi2=0.0, abs_psi=0.0, prof=np.zeros((Z, 3)))
    for s in range(sweep0, sweep0 + nsw):
        cb.sweep(st, am.a, seed, rep, s)
        acc["abs_a"] += abs(am.a)
        acc["e_field"] += lo.surface_field_energy(st, am.a, am.a)
        acc["e_surf"] += lo.surface_energy(st.k, st.r, P["gamma"], P["kappa"], P["intrinsic_curvature"], am.a)
        acc["psi2"] += float(np.mean(st.psi2))
        acc["abs_psi"] += float(np.mean(np.abs(st.psi)))
        ps, pa = lo.measure_profiles(st.psi)
        acc["prof"] += np.stack([ps.real, ps.imag, pa], 1)
        if amp_moves:
            cb.amplitude_move(st, am, P, seed, rep, s)
    return st, am, acc


def _check_against_oracle(_abi, psi_gpu, chain, obs, prof, st, am, acc, nsw):
    ch = chain.cpu().numpy()
    assert ch[_abi.CH_AMP_ACCEPTED] == am.accepted and ch[_abi.CH_AMP_PROPOSED] == am.proposed
    assert ch[_abi.CH_SITE_ACCEPTED] == st.accepted and ch[_abi.CH_STEP_COUNTER] == st.step_counter
    assert ch[_abi.CH_AMPLITUDE] == pytest.approx(am.a, rel=1e-12, abs=1e-15)
    assert ch[_abi.CH_SIGMA_AMP] == pytest.approx(am.sigma, rel=1e-12)
    assert ch[_abi.CH_SIGMA_SITE] == pytest.approx(st.sigma, rel=1e-12)
    Z = st.psi.shape[0]
    np.testing.assert_allclose(psi_gpu.cpu().numpy()[:Z], st.psi, rtol=0, atol=1e-11)
    o = obs.cpu().numpy()
    assert o[_abi.OB_COUNT] == nsw
    assert o[_abi.OB_ABS_A] == pytest.approx(acc["abs_a"], rel=1e-12, abs=1e-15)
    assert o[_abi.OB_E_FIELD] == pytest.approx(acc["e_field"], rel=1e-10)
    assert o[_abi.OB_E_SURF] == pytest.approx(acc["e_surf"], rel=1e-10)
    assert o[_abi.OB_PSI2] == pytest.approx(acc["psi2"], rel=1e-11)
    assert o[_abi.OB_ABSPSI] == pytest.approx(acc["abs_psi"], rel=1e-11)
    np.testing.assert_allclose(prof.cpu().numpy()[:Z], acc["prof"], rtol=1e-10, atol=1e-12)


# (Z, variant, benchmark temperature?)  -- all six row counts at the benchmark's temperature with the benchmark's variant, and
# all four variants at one odd and one even row count at a temperature where amplitude moves are accepted within a few sweeps
SHAPE_CASES = [(Z, lo.SIMPLE, True) for Z in (100, 91, 62, 59, 43, 40)]
SHAPE_CASES += [(Z, v, False) for Z in (91, 62) for v in (lo.SIMPLE, lo.RESCALED_0TH, lo.RESCALED_1ST, lo.DECOUPLED)]
SHAPE_CASES += [(59, lo.RESCALED_0TH, True), (100, lo.RESCALED_1ST, True)]


@pytest.mark.parametrize("Z,variant,bench_temp", SHAPE_CASES)
def test_smem_f64_bench_shapes_match_oracle(cuda, Z, variant, bench_temp):
    """fp64 sweep_smem with ADAPT | AMP_MOVES | MEASURE at T = 50, Z in {100, 91, 62, 59, 43, 40}: lattice to 1e-11, identical
    site / amplitude accept counts, widths to 1e-12, accumulated energies to 1e-10, profiles to 1e-10 -- same asserts as the
    small-lattice test (test_gpu_parity.py::test_smem_sweep_with_amplitude_moves_matches_oracle)."""
    import torch
    from cylinder_b200 import ops, _abi
    k = _k_for_rows(Z)
    assert ops.lattice_dims((50, 50), k)[0] == Z
    if bench_temp:          # a point of C3's grid
        P = dict(wavenumber=k, radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0, n=6, temperature=.01,
                 amp_bound=.8, amp_target_acceptance=.3)
        a0, sig, sig_a, mag = 0.0, .005, .005, .1
    else:
        P = dict(wavenumber=k, radius=1.1, gamma=1.0, kappa=0.3, intrinsic_curvature=0.2, alpha=-.8, u=1.2, C=.9, n=3, temperature=.3,
                 amp_bound=.8, amp_target_acceptance=.3)
        a0, sig, sig_a, mag = 0.1, .05, .05, .4
    T = BENCH_T
    seed, rep, sweep0, nsw = 0x5EED + Z, 4096 + Z, 3, 8
    psi0 = cb.random_initialize(Z, T, mag, seed, rep)
    st, am, acc = _oracle_run(psi0, P, variant, a0, sig, sig_a, seed, rep, sweep0, nsw)
    par = ops.make_params(cuda, **P)
    psi = torch.tensor(psi0, dtype=torch.complex128, device=cuda)[None].contiguous()
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    chain = ops.new_chain(1, cuda, amplitude=a0, sigma_site=sig, sigma_amp=sig_a)
    obs, prof = ops.new_obs(1, Z, cuda)
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
    ops.sweep_smem(psi, Zt, par, chain, seed, sweep0, nsw, variant, flags, replica0=rep, obs=obs, prof=prof)
    _check_against_oracle(_abi, psi[0], chain[0], obs[0], prof[0], st, am, acc, nsw)
    if not bench_temp:
        assert am.accepted >= 1, "case should exercise an accepted amplitude move"


def test_smem_f64_c3_grid_slice_matches_oracle(cuda):
    """One full k-slice of C3's grid (all 16 wavenumbers => 16 different row counts in [40, 100], ragged batch, even and odd) in
    ONE launch, exactly as bench.py lays the batch out, fp64, against the oracle replica by replica."""
    import torch
    from cylinder_b200 import ops, _abi
    B, T = 16, BENCH_T
    ia, ic = 5, 9
    P = dict(wavenumber=KS, radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=float(ALPHAS[ia]), u=1.0, C=float(CS[ic]), n=6.0,
             temperature=.01, amp_bound=.8, amp_target_acceptance=.3)
    Zs = [ops.lattice_dims((50, 50), k)[0] for k in KS]
    assert min(Zs) == 40 and max(Zs) == 100 and any(z & 1 for z in Zs) and any(not (z & 1) for z in Zs)
    Zmax = max(Zs)
    seed, rep0, nsw = 0x5EED, 256 * ia + 16 * ic, 6
    par = ops.make_params(cuda, **P)
    Zt = torch.tensor(Zs, dtype=torch.int32, device=cuda)
    psi = torch.zeros(B, Zmax, T, dtype=torch.complex128, device=cuda)
    ops.lattice_init(psi, Zt, .1, seed, rep0)
    start = psi.cpu().numpy()
    chain = ops.new_chain(B, cuda)
    obs, prof = ops.new_obs(B, Zmax, cuda)
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
    ops.sweep_smem(psi, Zt, par, chain, seed, 0, nsw, lo.SIMPLE, flags, replica0=rep0, obs=obs, prof=prof)
    n_amp_acc = 0
    for b in range(B):
        Pb = dict(P, wavenumber=float(KS[b]))
        st, am, acc = _oracle_run(start[b, :Zs[b]], Pb, lo.SIMPLE, 0.0, .005, .005, seed, rep0 + b, 0, nsw)
        _check_against_oracle(_abi, psi[b], chain[b], obs[b], prof[b], st, am, acc, nsw)
        assert torch.all(psi[b, Zs[b]:] == 0)
        n_amp_acc += am.accepted
    assert n_amp_acc >= 1


@pytest.mark.parametrize("Z,T", [(100, BENCH_T), (62, BENCH_T), (40, BENCH_T), (96, 260)])
def test_f32_smem_and_hbm_agree_with_amplitude_moves(cuda, Z, T):
    """The two fp32 kernels run the same arithmetic on the same Philox streams -- also for the amplitude move: identical
    lattices, chain state and accept counts after sweeps with ADAPT | AMP_MOVES | MEASURE on the benchmark's shapes (the HBM path
    needs both lengths even).  96 x 260 is the largest lattice the shared-memory kernel holds that has a tile away from every edge
    on the HBM path (3 x 3 tiles of 32 x 128, the last column cut): there the HBM path takes the energy of the move from its pass
    over the finished tile, with its own summation order -- the decisions still agree (a flip needs T ln u + dE below the fp32
    rounding of the sums, ~1e-6 relative)."""
    import torch
    from cylinder_b200 import ops, _abi
    k = _k_for_rows(Z)
    P = dict(wavenumber=k, radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0, n=6, temperature=.01)
    par = ops.make_params(cuda, **P)
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    psi = torch.zeros(1, Z, T, dtype=torch.complex64, device=cuda)
    ops.lattice_init(psi, Zt, .1, seed=5)
    ws = ops.HbmWorkspace(Z, T, torch.complex64, cuda)
    ws.pack(psi[0])
    c1, c2 = ops.new_chain(1, cuda, amplitude=.1), ops.new_chain(1, cuda, amplitude=.1)
    o1, p1 = ops.new_obs(1, Z, cuda)
    o2, p2 = ops.new_obs(1, Z, cuda)
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
    nsw = 40
    ops.sweep_smem(psi, Zt, par, c1, 42, 0, nsw, 0, flags, obs=o1, prof=p1)
    ws.sweep(par, c2[0], 42, 0, nsw, 0, flags, obs=o2[0], prof=p2[0])
    assert c1[0, _abi.CH_AMP_ACCEPTED].item() >= 1
    assert c2[0, _abi.CH_AMP_ACCEPTED].item() == c1[0, _abi.CH_AMP_ACCEPTED].item()
    assert c2[0, _abi.CH_SITE_ACCEPTED].item() == c1[0, _abi.CH_SITE_ACCEPTED].item()
    assert c2[0, _abi.CH_AMPLITUDE].item() == c1[0, _abi.CH_AMPLITUDE].item()
    assert torch.equal(ws.unpack(), psi[0])
    # the measured sums are fp32 partial sums in a different order on the two paths
    np.testing.assert_allclose(o2.cpu().numpy()[0, :10], o1.cpu().numpy()[0, :10], rtol=2e-5)
    np.testing.assert_allclose(p2.cpu().numpy()[..., 2], p1.cpu().numpy()[..., 2], rtol=1e-5)
    np.testing.assert_allclose(p2.cpu().numpy()[..., :2], p1.cpu().numpy()[..., :2], rtol=0, atol=2e-3)


@pytest.mark.parametrize("Z,variant", [(100, lo.SIMPLE), (91, lo.SIMPLE), (59, lo.RESCALED_0TH), (40, lo.SIMPLE)])
def test_f32_smem_bench_shapes_track_oracle(cuda, Z, variant):
    """The fp32 instantiation (the one bench.py times) against the fp64 oracle on the benchmark's shapes WITH amplitude moves
    and measurement, over one sweep from the same start: every site whose accept decision is not borderline ends within 2e-6 of
    the oracle's value, at most 0.5 % of the sites may differ (borderline exp comparisons, MUFU transcendentals), and the
    measured energy of the swept lattice agrees to 2e-5 relative.  Tolerances: fp32 psi has 2^-24 relative rounding per stored
    value; the field energy sums Z*T terms of mixed sign."""
    import torch
    from cylinder_b200 import ops, _abi
    T = BENCH_T
    k = _k_for_rows(Z)
    P = dict(wavenumber=k, radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0, n=6, temperature=.01,
             amp_bound=.8, amp_target_acceptance=.3)
    seed, rep = 77, 900 + Z
    par = ops.make_params(cuda, **P)
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    psi32 = torch.zeros(1, Z, T, dtype=torch.complex64, device=cuda)
    ops.lattice_init(psi32, Zt, .1, seed, rep)
    # a few fp32 sweeps first so that the start is not the small-amplitude initial state; the oracle starts from the SAME fp32 field
    c0 = ops.new_chain(1, cuda, amplitude=.2, sigma_site=.02)
    ops.sweep_smem(psi32, Zt, par, c0, seed, 0, 20, variant, _abi.SWEEP_ADAPT_SIGMA, replica0=rep)
    start = psi32[0].cpu().numpy().astype(np.complex128)
    sig = float(c0[0, _abi.CH_SIGMA_SITE].item())
    st, am, acc = _oracle_run(start, P, variant, .2, sig, .005, seed, rep, 20, 1)
    chain = ops.new_chain(1, cuda, amplitude=.2, sigma_site=sig, sigma_amp=.005)
    obs, prof = ops.new_obs(1, Z, cuda)
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
    ops.sweep_smem(psi32, Zt, par, chain, seed, 20, 1, variant, flags, replica0=rep, obs=obs, prof=prof)
    # the oracle's amplitude move changed nothing in psi; compare the swept lattice
    diff = np.abs(psi32[0].cpu().numpy().astype(np.complex128) - st.psi)
    mism = float((diff > 2e-6).mean())
    assert mism < 5e-3, mism
    assert abs(chain[0, _abi.CH_SITE_ACCEPTED].item() - st.accepted) <= 0.005 * Z * T
    if mism == 0.0:
        o = obs[0].cpu().numpy()
        assert o[_abi.OB_E_FIELD] == pytest.approx(acc["e_field"], rel=2e-5)
        assert o[_abi.OB_PSI2] == pytest.approx(acc["psi2"], rel=2e-6)
        np.testing.assert_allclose(prof[0].cpu().numpy()[:, 2], acc["prof"][:, 2], rtol=2e-6)


@pytest.mark.parametrize("dtype_name", ["complex64", "complex128"])
def test_smem_one_launch_equals_many_launches(cuda, dtype_name):
    """Exact resume on the benchmark's shapes: one launch of n sweeps, n launches of one sweep, and a last sweep taken by the
    plain-sweep instance (no amplitude machinery, compile-time row length) give bit-identical lattices, replica by replica --
    every state a launch carries in shared memory (row tables after an accepted amplitude move, widths, draws) is a function
    of the chain record alone.  Even and odd Z (2 and 3 colours), amplitude moves accepted in the first sweeps."""
    import sys, os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from cylinder_b200 import _abi
    from cylinder_b200.replicas import ReplicaBatch
    dt = getattr(torch, dtype_name)
    B, n = 32, 4
    kw = dict(dims=(50, 50), dtype=dt, device=cuda, seed=11)
    one = ReplicaBatch(bench.grid_params(B, 0), **kw)
    one.run(n, measure=False, index_order=True)
    many = ReplicaBatch(bench.grid_params(B, 0), **kw)
    for _ in range(n):
        many.run(1, measure=False, index_order=True)
    mixed = ReplicaBatch(bench.grid_params(B, 0), **kw)
    mixed.run(n - 1, measure=False, index_order=True)
    mixed.run(1, amp_moves=False, measure=False, index_order=True)      # site moves of the last sweep through the plain instance
    assert one.chain[:, _abi.CH_AMP_ACCEPTED].sum().item() >= B           # the launches did rebuild their tables
    assert torch.equal(one.psi, many.psi) and torch.equal(one.chain[:, :3], many.chain[:, :3])
    assert torch.equal(one.psi, mixed.psi)
