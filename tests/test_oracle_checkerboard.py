"""CPU checks of the parallel-schedule oracle (oracle/checkerboard_oracle.py): Philox known answers, colouring
validity incl. odd periodic rings, vectorised dE == the pinned scalar dE, and statistical equivalence of the
checkerboard ordering with the reference's sequential random-site ordering."""
import math
import random

import numpy as np
import pytest

from oracle import checkerboard_oracle as cb
from oracle import lattice_oracle as lo


def test_philox_known_answers():
    for ctr, key, exp in cb.PHILOX_KAT:
        out = cb.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(x[0]) for x in out) == exp
    for ctr, key, exp in cb.PHILOX2_KAT:
        out = cb.philox2x32_10(np.array([ctr[0]]), np.array([ctr[1]]), key)
        assert tuple(int(x[0]) for x in out) == exp


def test_site_draw_bit_fields_are_independent_and_uniform():
    x0, x1 = cb.philox2x32_10(np.arange(200000), 7, 0xABCDEF)
    zre, zim, u = cb.site_draws(x0, x1)
    assert abs(zre.mean()) < 0.01 and abs(zim.mean()) < 0.01
    assert zre.std() == pytest.approx(1.0, abs=0.01) and zim.std() == pytest.approx(1.0, abs=0.01)
    assert abs(np.mean(zre * zim)) < 0.01 and abs(np.corrcoef(zre ** 2 + zim ** 2, u)[0, 1]) < 0.01
    assert u.min() > 0 and u.max() < 1 and u.mean() == pytest.approx(.5, abs=.005)


@pytest.mark.parametrize("Z,T", [(4, 4), (6, 8), (5, 4), (4, 7), (5, 7), (71, 25), (91, 50), (3, 3), (2, 2)])
def test_colouring_is_proper(Z, T):
    cmap = cb.colour_map(Z, T)
    nc = cb.num_colours(Z, T)
    assert cmap.min() == 0 and cmap.max() <= nc - 1
    if Z > 2:
        assert np.all(cmap != np.roll(cmap, 1, axis=0))
    if T > 2:
        assert np.all(cmap != np.roll(cmap, 1, axis=1))
    assert (nc == 2) == (Z % 2 == 0 and T % 2 == 0)


def _state(Z, T, variant, seed=3, **kw):
    rng = np.random.RandomState(seed)
    psi = rng.normal(0, .4, (Z, T)) + 1j * rng.normal(0, .4, (Z, T))
    P = dict(k=.8, r=1.1, alpha=-.7, u=1.3, C=.9, n=3, temp=.05)
    P.update(kw)
    return lo.LatticeState(psi, P["k"], P["r"], P["alpha"], P["u"], P["C"], P["n"], P["temp"], variant=variant)


ALL_VARIANTS = [lo.SIMPLE, lo.RESCALED_0TH, lo.RESCALED_1ST, lo.DECOUPLED]


@pytest.mark.parametrize("variant", ALL_VARIANTS)
@pytest.mark.parametrize("a", [0.0, 0.35, -0.7])
def test_vectorised_dE_equals_scalar(variant, a):
    st = _state(9, 6, variant)
    rng = np.random.RandomState(1)
    I, J = np.meshgrid(np.arange(st.Z), np.arange(st.T_len), indexing="ij")
    I, J = I.ravel(), J.ravel()
    new = st.psi[I, J] + rng.normal(0, .05, I.size) + 1j * rng.normal(0, .05, I.size)
    dE_vec, _ = cb.delta_energy_vec(st, a, I, J, new)
    for q in range(I.size):
        iz = int(I[q]) if I[q] < st.Z - 1 else -1          # reference index convention: [-1, len-2]
        ith = int(J[q]) if J[q] < st.T_len - 1 else -1
        dE, _ = lo.delta_energy(st, a, iz, ith, complex(new[q]))
        assert dE_vec[q] == pytest.approx(dE, rel=1e-11, abs=1e-15)


def test_sweep_touches_every_site_once_and_is_reproducible():
    st1, st2 = _state(7, 5, lo.SIMPLE), _state(7, 5, lo.SIMPLE)
    n1, d1 = cb.sweep(st1, .2, seed=99, replica=4, sweep_index=11, return_details=True)
    n2, d2 = cb.sweep(st2, .2, seed=99, replica=4, sweep_index=11, return_details=True)
    assert n1 == n2 and np.array_equal(st1.psi, st2.psi)
    assert np.all(d1["u"] > 0)                 # every site drew
    st3 = _state(7, 5, lo.SIMPLE)
    cb.sweep(st3, .2, seed=99, replica=5, sweep_index=11)
    assert not np.array_equal(st1.psi, st3.psi)


def _sequential_chain(st, a, nsweeps, rng):
    """The reference's ordering: uniformly random site, sigma adapted after every proposal."""
    n = st.Z * st.T_len
    out = []
    for s in range(nsweeps):
        for _ in range(n):
            iz, ith = rng.randrange(-1, st.Z - 1), rng.randrange(-1, st.T_len - 1)
            lo.step_lattice_loc(st, a, iz, ith, rng.gauss(0, 1), rng.gauss(0, 1), rng.random)
        out.append(float(np.mean(st.psi2)))
    return np.array(out)


def _batch_err(x, nb=10):
    b = x[: len(x) // nb * nb].reshape(nb, -1).mean(1)
    return b.std(ddof=1) / math.sqrt(nb)


@pytest.mark.parametrize("variant", ALL_VARIANTS)
def test_checkerboard_matches_sequential_statistics(variant):
    """<|psi|^2> of the checkerboard ordering agrees with the sequential ordering within statistical error
    (6x5 lattice incl. an odd ring, a = 0.3)."""
    kw = dict(k=1.0, r=1.0, alpha=-1.0, u=1.0, C=1.0, n=2, temp=.3)
    a, nsw, burn = 0.3, 1500, 300
    st_seq = _state(6, 5, variant, seed=5, **kw)
    st_cb = _state(6, 5, variant, seed=6, **kw)
    st_seq.sigma = st_cb.sigma = .3
    seq = _sequential_chain(st_seq, a, nsw, random.Random(42))[burn:]
    chk = []
    for s in range(nsw):
        cb.sweep(st_cb, a, seed=7, replica=0, sweep_index=s)
        chk.append(float(np.mean(st_cb.psi2)))
    chk = np.array(chk[burn:])
    err = math.hypot(_batch_err(seq), _batch_err(chk))
    assert abs(seq.mean() - chk.mean()) < 5 * err, (seq.mean(), chk.mean(), err)
    # both adapt towards 50 % acceptance
    assert 0.4 < st_cb.accepted / st_cb.step_counter < 0.6
    assert 0.4 < st_seq.accepted / st_seq.step_counter < 0.6
