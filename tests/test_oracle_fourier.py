"""The Fourier-model oracle (oracle/fourier_oracle.py) against golden tensors and energies recorded from the
reference's Cylinder2D / Cylinder1D (tests/golden/fourier.npz).  CPU only."""
import os

import numpy as np
import pytest

from oracle import fourier_oracle as fo

CASES = ["k05_n6_21", "k1_n1_11", "k08_n2_30", "k12_n3_02"]
NAMES = ("wavenumber", "radius", "alpha", "C", "u", "n", "kappa", "gamma", "intrinsic_curvature")


def load(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "fourier.npz"))
    P = dict(zip(NAMES, g[tag + "_params"].tolist()))
    return g, P, tuple(int(x) for x in g[tag + "_nfc"])


@pytest.mark.parametrize("tag", CASES)
def test_tensor_form_matches_reference(golden_dir, tag):
    g, P, (nz, nth) = load(golden_dir, tag)
    c = g[tag + "_coeffs"]
    for ia in (0, 1, 3):                      # a = 0 (radius branch), 0.3, 0.9
        a = float(g[tag + "_amps"][ia])
        E, Aint, B = fo.field_energy_tensor(P, nz, nth, a, c)
        np.testing.assert_allclose(Aint, g[tag + "_Aint"][ia], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(B, g[tag + "_B"][ia], rtol=1e-12, atol=1e-12)
        assert E == pytest.approx(float(g[tag + "_E_change"][ia]), rel=1e-12)
        assert float(g[tag + "_E_stored"][ia]) == pytest.approx(float(g[tag + "_E_change"][ia]), rel=1e-14)


@pytest.mark.parametrize("tag", CASES)
def test_realspace_form_matches_reference(golden_dir, tag):
    """The single-integral form the GPU evaluates reproduces the reference's tensor contraction to <= 1e-10."""
    g, P, (nz, nth) = load(golden_dir, tag)
    c = g[tag + "_coeffs"]
    for ia, a in enumerate(g[tag + "_amps"]):
        assert fo.field_energy_realspace(P, nz, nth, float(a), c) == pytest.approx(float(g[tag + "_E_change"][ia]), rel=1e-10)


@pytest.mark.parametrize("tag", CASES)
def test_surface_energy_1d(golden_dir, tag):
    g, P, _ = load(golden_dir, tag)
    for ia, a in enumerate(g[tag + "_amps"]):
        e = g[tag + "_E_surf"][ia]
        assert abs(e.imag) < 1e-9
        assert fo.surface_energy_1d(P, float(a)) == pytest.approx(e.real, rel=1e-12)
