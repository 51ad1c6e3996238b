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
