"""Ensemble parity: the parallel (checkerboard, Philox) sampler on the GPU against the reference's sequential
random-site sampler as restated by the oracle -- <|psi|^2>, acceptance rate and <|a|> agree within statistical error."""
import math
import random

import numpy as np
import pytest

from oracle import lattice_oracle as lo

pytestmark = pytest.mark.gpu


def _batch_err(x, nb=10):
    b = x[: len(x) // nb * nb].reshape(nb, -1).mean(1)
    return b.std(ddof=1) / math.sqrt(nb)


@pytest.mark.parametrize("variant,dtype,Z,T", [("simple", "complex64", 6, 5), ("rescaled_0thorder", "complex64", 6, 5),
                                               ("simple", "complex128", 6, 5), ("simple", "complex64", 7, 6), ("simple", "complex64", 6, 6),
                                               ("rescaled_1storder", "complex64", 7, 6), ("decoupled", "complex64", 6, 6)])
def test_ensemble_matches_sequential_oracle(cuda, variant, dtype, Z, T):
    """(6, 5): odd theta ring, generic 3-colour loop; (7, 6): column-mapped 3-colour loop; (6, 6): column-mapped 2-colour loop."""
    import torch
    from cylinder_b200.replicas import ReplicaBatch
    P = dict(wavenumber=1.0, radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0, n=2.0, temperature=.3)
    a = 0.3
    # oracle: one long sequential chain (reference ordering, per-proposal Robbins-Monro)
    rng = random.Random(2)
    psi0 = np.array([[complex(rng.gauss(0, .3), rng.gauss(0, .3)) for _ in range(T)] for _ in range(Z)])
    st = lo.LatticeState(psi0, 1.0, 1.0, -1.0, 1.0, 1.0, 2.0, .3, variant=lo.VARIANTS[variant])
    st.sigma = .3
    seq = []
    for s in range(1800):
        for _ in range(Z * T):
            lo.step_lattice_loc(st, a, rng.randrange(-1, Z - 1), rng.randrange(-1, T - 1), rng.gauss(0, 1), rng.gauss(0, 1), rng.random)
        seq.append(float(st.psi2.mean()))
    seq = np.array(seq[300:])
    # GPU: 256 independent replicas of the same system
    B = 256
    batch = ReplicaBatch(dict(P, wavenumber=np.full(B, 1.0)), dims=(Z, T), variant=variant, dtype=getattr(torch, dtype), device=cuda,
                         seed=5, amplitude=a, sigma_site=.3)
    batch.run(400, amp_moves=False, measure=False)
    batch.run(600, amp_moves=False, measure=True)
    obs = batch.observables()
    gpu = obs["psi_squared"].cpu().numpy()
    gpu_mean, gpu_err = gpu.mean(), gpu.std(ddof=1) / math.sqrt(B)
    err = math.hypot(_batch_err(seq), gpu_err)
    assert abs(gpu_mean - seq.mean()) < 5 * err, (gpu_mean, seq.mean(), err)
    assert obs["site_acceptance"].mean().item() == pytest.approx(.5, abs=.05)
    assert st.accepted / st.step_counter == pytest.approx(.5, abs=.05)
    # profiles: <|psi|>(z) is periodic in z with the surface modulation; sanity: positive, finite, right shape
    prof, prof_abs = batch.profiles()
    assert prof_abs.shape == (B, Z) and torch.isfinite(prof_abs).all() and (prof_abs > 0).all()


def test_amplitude_distribution_matches_oracle(cuda):
    """With amplitude moves on: <|a|> and <a^2> of the GPU ensemble vs a sequential oracle chain that interleaves the
    reference's site proposals with engine-style amplitude moves (same energies, same accept rule)."""
    import torch
    from oracle import checkerboard_oracle as cb
    from cylinder_b200.replicas import ReplicaBatch
    P = dict(wavenumber=1.0, radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0, n=2.0, temperature=.5)
    Z, T = 6, 6
    rng = random.Random(9)
    psi0 = np.array([[complex(rng.gauss(0, .3), rng.gauss(0, .3)) for _ in range(T)] for _ in range(Z)])
    st = lo.LatticeState(psi0, 1.0, 1.0, -1.0, 1.0, 1.0, 2.0, .5)
    st.sigma = .3
    am = cb.AmpState(amplitude=0.0, sigma_amp=.1, target=.3, bound=.8)
    abs_a = []
    for s in range(2500):
        for _ in range(Z * T):
            lo.step_lattice_loc(st, am.a, rng.randrange(-1, Z - 1), rng.randrange(-1, T - 1), rng.gauss(0, 1), rng.gauss(0, 1), rng.random)
        cb.amplitude_move(st, am, P, seed=123, replica=0, sweep_index=s)
        abs_a.append(abs(am.a))
    abs_a = np.array(abs_a[500:])
    B = 256
    batch = ReplicaBatch(dict(P, wavenumber=np.full(B, 1.0)), dims=(Z, T), dtype=torch.complex64, device=cuda, seed=8, amplitude=0.0,
                         sigma_site=.3, sigma_amp=.1)
    batch.run(500, measure=False)
    batch.run(1000)
    obs = batch.observables()
    g = obs["abs_amplitude"].cpu().numpy()
    err = math.hypot(_batch_err(abs_a), g.std(ddof=1) / math.sqrt(B))
    assert abs(g.mean() - abs_a.mean()) < 5 * err, (g.mean(), abs_a.mean(), err)
    assert obs["amplitude_acceptance"].mean().item() == pytest.approx(.3, abs=.06)
    assert (batch.chain[:, 0].abs() < .8).all()


# ------------------------------------------------------------------------------------------------------
# BASELINE config C1 / one point of C3's grid at full size: 50 x 50 base lattice, T = 0.01, alpha = -1, u = 1, C = 1, n = 6
# (infile.txt:1-19), fp32 storage -- the kernel instantiation and the shapes bench.py times.
#
# At this temperature the slow modes of the field need > 10^4 sweeps to equilibrate (the sequential reference does ~5 x 10^4
# proposals/s/core: 20 sweeps a second), so "run both to equilibrium and compare" is not affordable for the CPU side.  Two
# tests instead:
#   (A) paired continuation.  The GPU equilibrates 512 replicas (2 x 10^4 sweeps); the final lattices of 16 of them are
#       handed to 16 sequential reference chains (oracle, one per host core, random-site order, per-proposal Robbins-Monro),
#       and BOTH samplers continue from these states for the same number of sweeps.  If the stationary distributions agree,
#       the start states are draws from it for both and every expectation agrees; the slow modes are shared by each pair, so
#       the paired differences are sharp.  A bias of the fp32 / checkerboard / 23-bit-draw sampler in any mode that relaxes
#       within the window (|psi|^2, acceptance, the z-profile <|psi|>(z)) shows up as a non-zero mean difference.
#   (B) independent chains.  8 reference chains and 512 GPU replicas from the same distribution of random starts
#       (On_lattice_simple.py:159-182), same schedule; ensemble means within 5 sigma of the (larger) reference scatter.
# ------------------------------------------------------------------------------------------------------
C1 = dict(radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01,
          amp_bound=.8, amp_target_acceptance=.3)


def _paired(o, g):
    """mean and standard error of the paired differences"""
    d = np.asarray(o) - np.asarray(g)
    return d.mean(0), d.std(0, ddof=1) / math.sqrt(d.shape[0])


@pytest.mark.parametrize("k,variant,amp_moves", [(1.0, "simple", False), (0.55, "simple", False), (1.0, "rescaled_0thorder", False),
                                                 (1.0, "simple", True)])
def test_c1_paired_continuation_matches_sequential_reference(cuda, k, variant, amp_moves):
    import torch
    from cylinder_b200 import _abi
    from cylinder_b200.replicas import ReplicaBatch
    from oracle import sequential_runs as sr
    B, NP, NSW = 512, 16, 120
    P = dict(C1, wavenumber=k)
    batch = ReplicaBatch(dict(P, wavenumber=np.full(B, k)), dims=(50, 50), variant=variant, dtype=torch.complex64, device=cuda,
                         seed=0xC1, amplitude=0.0 if amp_moves else 0.3)
    Z, T = int(batch.Z_host[0]), batch.T
    assert (Z, T) == ((50, 50) if k == 1.0 else (91, 50))
    batch.run(20000, amp_moves=amp_moves, measure=False)
    ch0 = batch.chain[:NP].cpu().numpy().copy()
    starts = batch.psi[:NP, :Z].cpu().numpy().astype(np.complex128)
    batch.reset_observables()
    batch.run(NSW, amp_moves=amp_moves, measure=True)
    ch1 = batch.chain[:NP].cpu().numpy()
    obs = batch.obs[:NP].cpu().numpy()
    g_psi2, g_abs = obs[:, _abi.OB_PSI2] / NSW, obs[:, _abi.OB_ABSPSI] / NSW
    g_acc = (ch1[:, _abi.CH_SITE_ACCEPTED] - ch0[:, _abi.CH_SITE_ACCEPTED]) / (NSW * Z * T)
    g_prof = batch.prof[:NP, :Z, 2].cpu().numpy() / (NSW * T)
    tasks = [dict(psi0=starts[i], P=P, variant=lo.VARIANTS[variant], amplitude=ch0[i, _abi.CH_AMPLITUDE],
                  sigma=ch0[i, _abi.CH_SIGMA_SITE], step_counter=int(ch0[i, _abi.CH_STEP_COUNTER]), nsweeps=NSW, seed=1000 + i,
                  amp_moves=amp_moves, sigma_amp=ch0[i, _abi.CH_SIGMA_AMP], amp_counter=int(ch0[i, _abi.CH_AMP_COUNTER]))
             for i in range(NP)]
    res = sr.run_chains(tasks)
    for name, o, g in (("psi2", [r["psi2"] for r in res], g_psi2), ("abs_psi", [r["abs_psi"] for r in res], g_abs),
                       ("acceptance", [r["acceptance"] for r in res], g_acc)):
        m, se = _paired(o, g)
        assert abs(m) < 5 * se, (name, m, se, np.mean(o), np.mean(g))
    # field profile <|psi|>(z), row by row: pooled standard error over the rows (the rows' scatters are alike)
    m, se = _paired(np.stack([r["profile_abs"] for r in res]), g_prof)
    pooled = math.sqrt(float(np.mean(se ** 2)))
    assert np.all(np.abs(m) < 5 * pooled), (float(np.abs(m).max()), pooled)
    if not amp_moves:                           # a = 0.3: the order parameter is suppressed where n A_theta is large
        assert g_prof.max() > 1.3 * g_prof.min(), "the profile should be non-trivial at this amplitude"
    # the paired differences are small compared with the observable itself (the test has power)
    assert pooled < .05 * float(g_prof.mean())
    if amp_moves:
        g_amp_acc = (ch1[:, _abi.CH_AMP_ACCEPTED] - ch0[:, _abi.CH_AMP_ACCEPTED]) / NSW
        m, se = _paired([r["amp_acceptance"] for r in res], g_amp_acc)
        assert abs(m) < 5 * se, ("amp_acceptance", m, se)
        m, se = _paired([r["abs_a"] for r in res], obs[:, _abi.OB_ABS_A] / NSW)
        assert abs(m) < 5 * se, ("abs_a", m, se)
        # (the run-long ensemble acceptance is NOT asserted against the Robbins-Monro target .3: the width recursion moves by
        # O(1/n) per proposal and proposals beyond amp_bound count as rejections, so after 2 x 10^4 sweeps the mean still sits
        # below the target -- for the sequential reference chains just as for the kernel, which is what the pairs check)


def test_c1_independent_chains_match_sequential_reference(cuda):
    import torch
    from oracle import checkerboard_oracle as cb
    from oracle import sequential_runs as sr
    from cylinder_b200.replicas import ReplicaBatch
    B, NC, BURN, NSW, a = 512, 8, 80, 120, 0.3
    P = dict(C1, wavenumber=1.0)
    batch = ReplicaBatch(dict(P, wavenumber=np.full(B, 1.0)), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=77, amplitude=a)
    batch.run(BURN, amp_moves=False, measure=False)
    batch.run(NSW, amp_moves=False, measure=True)
    g = batch.observables()
    g_psi2 = g["psi_squared"].cpu().numpy()
    g_prof = (batch.prof[:, :, 2] / (NSW * 50)).cpu().numpy()
    tasks = [dict(psi0=cb.random_initialize(50, 50, .1, 77, 100000 + i), P=P, amplitude=a, sigma=.005, nsweeps=NSW, burn=BURN, seed=500 + i)
             for i in range(NC)]
    res = sr.run_chains(tasks)
    o_psi2 = np.array([r["psi2"] for r in res])
    err = math.hypot(o_psi2.std(ddof=1) / math.sqrt(NC), g_psi2.std(ddof=1) / math.sqrt(B))
    assert abs(o_psi2.mean() - g_psi2.mean()) < 5 * err, (o_psi2.mean(), g_psi2.mean(), err)
    o_acc = np.array([r["acceptance"] for r in res])
    assert o_acc.mean() == pytest.approx(g["site_acceptance"].mean().item(), abs=.01)
    o_prof = np.stack([r["profile_abs"] for r in res])
    perr = np.hypot(o_prof.std(0, ddof=1) / math.sqrt(NC), g_prof.std(0, ddof=1) / math.sqrt(B))
    pooled = math.sqrt(float(np.mean(perr ** 2)))
    assert np.all(np.abs(o_prof.mean(0) - g_prof.mean(0)) < 5 * pooled), (float(np.abs(o_prof.mean(0) - g_prof.mean(0)).max()), pooled)
