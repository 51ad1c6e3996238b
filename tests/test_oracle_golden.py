"""The oracle (oracle/lattice_oracle.py) against golden vectors recorded from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import glob
import math
import os

import numpy as np
import pytest

from oracle import lattice_oracle as lo

LATTICE_CASES = sorted(os.path.basename(p)[len("lattice_"):-4]
                       for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lattice_*.npz")))


def load_case(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "lattice_%s.npz" % tag))
    P = dict(zip(g["params_names"].tolist(), g["params"].tolist()))
    variant = lo.VARIANTS[str(g["variant"])[len("On_lattice_"):]]
    st = lo.LatticeState(g["psi0"], P["wavenumber"], P["radius"], P["alpha"], P["u"], P["C"], P["n"],
                         P["temperature"], variant=variant, sigma=lo.INIT_SIGMA[variant])
    if "sigma0" in g.files:
        assert float(g["sigma0"]) == lo.INIT_SIGMA[variant]
    return g, P, st


def test_geometry(golden_dir):
    t = np.load(os.path.join(golden_dir, "geometry.npz"))["table"]
    for k, r, a, z, rr, sgt, gt, sgz, gz, ath in t:
        assert lo.radius_rescaled(r, a) == rr
        assert lo.sqrt_g_theta(k, r, a, z) == sgt
        assert lo.g_theta(k, r, a, z) == gt
        assert lo.sqrt_g_z(k, r, a, z) == sgz
        assert lo.g_z(k, r, a, z) == gz
        assert lo.A_theta(k, r, a, z) == ath


def test_surface_energy_tables(golden_dir):
    """surfenergy_H0.csv == evaluate_A_integral_0 (r=1); curvenergy_H0.csv == calc_bending_energy (kappa=1,H0=0)."""
    g = np.load(os.path.join(golden_dir, "surface_energy.npz"))
    for i, a in enumerate(g["a_vals"]):
        for j, k in enumerate(g["k_vals"]):
            assert lo.area_integral_0(k, 1.0, a) == pytest.approx(g["surfenergy_H0"][i, j], rel=1e-12)
            assert lo.bending_energy(k, 1.0, a, 0.0) == pytest.approx(g["curvenergy_H0"][i, j], rel=1e-10)


def test_surface_energy_direct(golden_dir):
    g = np.load(os.path.join(golden_dir, "surface_energy.npz"))
    for k, r, gam, kap, H0, a, A0, bend, E in g["direct"]:
        assert lo.area_integral_0(k, r, a) == A0
        assert lo.bending_energy(k, r, a, H0) == bend
        assert lo.surface_energy(k, r, gam, kap, H0, a) == E
        # the quadrature scheme the GPU uses reaches the 1e-10 bar
        assert lo.surface_energy_quadrature(k, r, gam, kap, H0, a) == pytest.approx(E, rel=1e-10)


@pytest.mark.parametrize("tag", LATTICE_CASES)
def test_lattice_replay(golden_dir, tag):
    g, P, st = load_case(golden_dir, tag)
    assert st.lz == float(g["z_pixel_len"]) and st.lth == float(g["th_pixel_len"])
    amp = float(g["amp"])
    # initial energies
    for a2, ar, e in g["field_energy_initial"]:
        assert lo.surface_field_energy(st, a2, ar) == pytest.approx(e, rel=1e-13)
    acc, dE, sig = lo.replay_tape(st, amp, g["iz"], g["ith"], g["zre"], g["zim"], g["u"])
    assert np.array_equal(acc, g["accept"])
    assert np.array_equal(dE, g["dE"])              # same operations in the same order: bit-exact
    assert np.array_equal(sig, g["sigma"])
    assert np.array_equal(st.psi, g["psi_final"])
    assert np.array_equal(st.psi2, g["psi2_final"])
    assert np.array_equal(st.dz, g["dz_final"])
    assert np.array_equal(st.dth, g["dth_final"])
    assert st.step_counter == int(g["step_counter"]) and st.accepted == int(g["acceptance_counter"])
    for a2, ar, e in g["field_energy_final"]:
        assert lo.surface_field_energy(st, a2, ar) == pytest.approx(e, rel=1e-13)
    for a2, e in g["surface_energy"]:
        assert lo.surface_energy(P["wavenumber"], P["radius"], P["gamma"], P["kappa"],
                                 P["intrinsic_curvature"], a2) == e
    ps, pa = lo.measure_profiles(st.psi)
    assert np.allclose(ps, g["profile_sum"], rtol=1e-13, atol=1e-15)
    assert np.allclose(pa, g["profile_abs_sum"], rtol=1e-13)


@pytest.mark.parametrize("tag", LATTICE_CASES)
def test_cache_invariant(golden_dir, tag):
    """dz / dth / psi_squared caches are pure functions of psi: the kernels need not store them."""
    g, P, st = load_case(golden_dir, tag)
    st.psi = g["psi_final"].copy()
    st.rebuild_caches()
    assert np.allclose(st.dz, g["dz_final"], rtol=0, atol=1e-13)
    assert np.allclose(st.dth, g["dth_final"], rtol=0, atol=1e-13)
    assert np.allclose(st.psi2, g["psi2_final"], rtol=1e-15)


def test_known_answer_uniform_field(golden_dir):
    """tests/test_On_lattice.py:75-88: psi==1, alpha=-1, u=1, a=0  =>  E = -2 pi^2 on both lattices."""
    g = np.load(os.path.join(golden_dir, "known_answers.npz"))
    for name, dims in (("fine", (500, 250)), ("coarse", (50, 25))):
        Z, T = lo.lattice_dims(dims, 1.0)
        st = lo.LatticeState(np.ones((Z, T), complex), 1.0, 1.0, -1, 1, 1, 1, 0.01)
        e = lo.surface_field_energy(st, 0.0)
        assert e == pytest.approx(-.5 * 4 * math.pi ** 2, abs=1e-7)
        assert e == pytest.approx(float(g[name]), rel=1e-13)
        assert lo.surface_field_energy(st, 0.4) == pytest.approx(float(g[name + "_a04"]), rel=1e-12)


def test_metropolis_decision_known_answers():
    """Recovered from tests/__pycache__/test_metropolis_engine*.pyc."""
    never = lambda: (_ for _ in ()).throw(AssertionError("no draw expected"))  # noqa: E731
    assert lo.metropolis_decision(1e-3, 12.45, -1451, never)[0] is True
    assert lo.metropolis_decision(0, -145.4, -145.3, never)[0] is False
    assert lo.metropolis_decision(0, -145.4, -145.5, never)[0] is True
    assert lo.metropolis_decision(0, 0, 0, never)[0] is True
    assert lo.metropolis_decision(1.5, 14.3, 14.3, never)[0] is True


def test_rm_ratio():
    assert lo.rm_ratio(.5, 1) == 4.0


def test_row_moment_form_matches_energy(golden_dir):
    """E(a') is linear in a-independent row sums (simple variant)."""
    g, P, st = load_case(golden_dir, "simple_c1_a03")
    S = lo.row_moments(st)
    for a in (0.0, 0.3, -0.55):
        e = 0.0
        for i in range(st.Z):
            z, zm = i * st.lz, (i - .5) * st.lz
            zs = lo.sqrt_g_z(st.k, st.r, a, zm)
            w0 = zs * lo.sqrt_g_theta(st.k, st.r, a, z)
            wir = lo.sqrt_g_z(st.k, st.r, a, z) / lo.sqrt_g_theta(st.k, st.r, a, zm)
            A = lo.A_theta(st.k, st.r, a, zm)
            e += ((st.alpha * S[i, 0] + st.u / 2 * S[i, 1]) * w0 + st.C * S[i, 2] * w0 / zs ** 2
                  + st.C * S[i, 3] * wir + 2 * st.C * st.n * A * S[i, 4] * wir + st.Cn2 * A * A * S[i, 0] * wir)
        assert e * st.lz * st.lth == pytest.approx(lo.surface_field_energy(st, a), rel=1e-12)
