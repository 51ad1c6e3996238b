"""Whole-lattice proposals (On_lattice_simple.py:247-382): the oracle's restatement of step_lattice_all against sequences
recorded from the unmodified reference, and the invariants of the adopted wavevector / row-twist semantics."""
import math
import os

import numpy as np
import pytest

from oracle import lattice_oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))


def _state(g, tag):
    P = g[tag + "_params"].tolist()
    st = lo.LatticeState(g[tag + "_psi0"], P[0], P[1], P[5], P[6], P[7], P[8], P[9], variant=lo.SIMPLE, sigma=.02)
    return st, float(g[tag + "_amp"])


@pytest.mark.parametrize("tag", ["warm", "cold"])
def test_step_lattice_all_reproduces_the_reference(tag):
    g = np.load(os.path.join(HERE, "golden", "collective.npz"))
    st, a = _state(g, tag)
    energy = lo.surface_field_energy(st, a, a)
    assert energy == pytest.approx(float(g[tag + "_e0"]), rel=1e-13)
    for i in range(len(g[tag + "_zre"])):
        assert energy == pytest.approx(float(g[tag + "_e_old"][i]), rel=1e-12)
        u = float(g[tag + "_u"][i])
        acc, e_new, u_used = lo.step_lattice_all(st, a, energy, float(g[tag + "_zre"][i]), float(g[tag + "_zim"][i]), u)
        assert e_new == pytest.approx(float(g[tag + "_e_new"][i]), rel=1e-12)
        assert acc == bool(g[tag + "_accept"][i])
        assert (u_used is None) == bool(np.isnan(u))                       # a uniform is drawn exactly when the reference draws one
        assert st.sigma == pytest.approx(float(g[tag + "_sigma"][i]), rel=1e-14)
        if acc:
            energy = e_new
    np.testing.assert_allclose(st.psi, g[tag + "_psi_final"], rtol=0, atol=1e-15)


def _small(variant=lo.SIMPLE, Z=9, T=8, seed=3, temp=.2):
    rng = np.random.RandomState(seed)
    psi = rng.normal(0, .4, (Z, T)) + 1j * rng.normal(0, .4, (Z, T))
    return lo.LatticeState(psi, .9, 1.1, -.8, 1.2, .9, 3, temp, variant=variant, sigma=.05)


def test_wave_array_and_wavevector_move():
    st = _small()
    w = lo.wave_array(.3 - .1j, (2 * math.pi, -4 * math.pi), st.Z, st.T_len)
    assert w[0, 0] == pytest.approx(.3 - .1j) and w.shape == (st.Z, st.T_len)
    assert w[2, 3] == pytest.approx((.3 - .1j) * np.exp(1j * (2 * 2 * math.pi / st.Z - 3 * 4 * math.pi / st.T_len)))
    e0 = lo.surface_field_energy(st, .2, .2)
    psi0 = st.psi.copy()
    acc, e_new, _ = lo.step_lattice_wavevector(st, .2, e0, (2 * math.pi, 2 * math.pi), .7, -.4, 0.999)
    expect = psi0 + lo.wave_array((.7 - .4j) * .005, (2 * math.pi, 2 * math.pi), st.Z, st.T_len)
    np.testing.assert_allclose(st.psi if acc else expect, expect, atol=1e-15)
    assert lo.surface_field_energy(st, .2, .2) == pytest.approx(e_new if acc else e0, rel=1e-13)
    assert st.bump_step_counter == 1 and st.bump_sigma != .005


@pytest.mark.parametrize("variant", [lo.SIMPLE, lo.RESCALED_0TH])
def test_row_twist_changes_the_winding_number_of_exactly_the_block(variant):
    st = _small(variant, Z=9, T=8, temp=0.0)
    st.psi = np.abs(st.psi) + .5 + 0j                                        # positive real field: every row winds 0 times
    st.rebuild_caches()
    before = lo.winding_numbers(st.psi)
    assert np.all(before == 0)
    st.temperature = 1e9                                                     # accept anything: isolates the proposal itself
    acc, e_old, e_new, _ = lo.step_row_rotate(st, .3, n_rot=1, phase=.7, num_rows=3, row_index=7, uniform=0.0)
    assert acc and e_new > e_old                                             # a twist costs gradient energy
    after = lo.winding_numbers(st.psi)
    assert after.tolist() == [1 if i in (7, 8, 0) else 0 for i in range(9)]  # the block wraps around the seam
    np.testing.assert_allclose(np.abs(st.psi), np.abs(st.psi), rtol=0)
    # undoing the twist restores the energy
    acc, e2_old, e2_new, _ = lo.step_row_rotate(st, .3, n_rot=-1, phase=-.7, num_rows=3, row_index=7, uniform=0.0)
    assert e2_new == pytest.approx(e_old, rel=1e-12) and np.all(lo.winding_numbers(st.psi) == 0)
