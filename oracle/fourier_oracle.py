"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package `cylinder_b200`).

CPU restatement of the reference's Fourier-mode field model (config 2):
  Cylinder2D.evaluate_A_integrals / evaluate_B_integrals / calc_field_energy_amplitude_change
      /root/reference/surfaces_and_fields/system_cylinder2D.py:64-198
  Cylinder1D.calc_surface_energy   system_cylinder1D.py:189-194
in the reference's own tensor form (scipy.integrate.quad + einsum).  Parity status: PINNED against
tests/golden/fourier.npz (A integrals, B matrices and energies recorded from the unmodified reference).
`field_energy_realspace` is the single-integral form the CUDA kernel evaluates, restated in numpy.
"""
import math

import numpy as np

from . import lattice_oracle as lo


def _quad(f, k):
    import scipy.integrate as integrate
    return integrate.quad(f, 0, 2 * math.pi / k)[0]


def A_integrals(k, r, a, nz):
    """tmp_A_integrals[d + 4 nz], d in [-4nz, 4nz]      system_cylinder2D.py:103-109, system_cylinder1D.py:37-59"""
    out = np.zeros(8 * nz + 1, dtype=complex)
    for d in range(-4 * nz, 4 * nz + 1):
        if a == 0:
            w = lambda z: r                                                   # noqa: E731
        else:
            w = lambda z: lo.sqrt_g_theta(k, r, a, z) * lo.sqrt_g_z(k, r, a, z)   # noqa: E731
        im = _quad(lambda z: math.sin(d * k * z) * w(z), k)
        re = _quad(lambda z: math.cos(d * k * z) * w(z), k)
        out[d + 4 * nz] = complex(re, im)
    return out


def B_matrix(k, r, n, a, nz, nth):
    """tmp_B_matrix[i+nz, j+nz, b+nth, b+nth]            system_cylinder2D.py:64-100,134-147"""
    lz, lth = 2 * nz + 1, 2 * nth + 1
    B = np.zeros((lz, lz, lth, lth), dtype=complex)
    for i in range(-nz, nz + 1):
        for j in range(-nz, nz + 1):
            for b in range(-nth, nth + 1):
                def body(z):
                    gz, gt = lo.sqrt_g_z(k, r, a, z), lo.sqrt_g_theta(k, r, a, z)
                    return k ** 2 * i * j * gz * gt + (b - n * lo.A_theta(k, r, a, z)) ** 2 * gz / gt
                im = _quad(lambda z: math.sin((i - j) * k * z) * body(z), k)
                re = _quad(lambda z: math.cos((i - j) * k * z) * body(z), k)
                B[i + nz, j + nz, b + nth, b + nth] = complex(re, im)
    return B


def field_energy_tensor(P, nz, nth, a, c):
    """calc_field_energy_amplitude_change(a, c)   system_cylinder2D.py:176-198.  P: dict wavenumber, radius, alpha, C, u, n;
    c: [2 nth + 1, 2 nz + 1] complex."""
    k, r = P["wavenumber"], P["radius"]
    lz, lth = 2 * nz + 1, 2 * nth + 1
    Aint = A_integrals(k, r, a, nz)
    A = np.zeros((lz, lz), dtype=complex)
    D = np.zeros((lz, lz, lz, lz), dtype=complex)
    for i in range(lz):                                                        # :114-118
        A[i, ::-1] = Aint[i + 2 * nz:i + 4 * nz + 1]
        for j in range(lz):
            for kk in range(lz):
                D[i, j, kk, ::-1] = Aint[i + j - kk + 2 * nz:i + j - kk + 4 * nz + 1]
    B = B_matrix(k, r, P["n"], a, nz, nth)
    Ae = sum(np.einsum("ij, i, j -> ", A, c[b], c[b].conjugate()) for b in range(lth))
    De = 0j
    for i1 in range(lth):                                                      # quartic selection rule :47-56
        for i2 in range(lth):
            for i3 in range(lth):
                for i4 in range(lth):
                    if i1 + i2 == i3 + i4:
                        De += np.einsum("ijkl, i, j, k, l -> ", D, c[i1], c[i2], c[i3].conjugate(), c[i4].conjugate())
    Be = np.einsum("ijab, ai, bj -> ", B, c, c.conjugate())
    return P["alpha"] * Ae.real + P["C"] * Be.real + 0.5 * P["u"] * De.real, Aint, B


def field_energy_realspace(P, nz, nth, a, c, nodes=256):
    """Same energy as one surface integral -- the scheme of csrc/k_fourier.cu."""
    k, r, n = P["wavenumber"], P["radius"], P["n"]
    ra = lo.radius_rescaled(r, a)
    L = 2 * math.pi / k
    z = L * np.arange(nodes) / nodes
    s, co = np.sin(k * z), np.cos(k * z)
    gth = ra * (1 + a * s)
    gz = np.sqrt(1 + (ra * a * k * co) ** 2)
    ath = k * ra * a * co / gz
    js = np.arange(-nz, nz + 1)
    bs = np.arange(-nth, nth + 1)
    ph = np.exp(1j * np.outer(js, k * z))                       # [lz, nodes]
    psi = c @ ph                                                # [lth, nodes]
    dpsi = (c * (1j * js * k)[None, :]) @ ph
    nthq = 4 * nth + 2
    th = 2 * math.pi * np.arange(nthq) / nthq
    Psi = np.einsum("bn,bm->nm", psi, np.exp(1j * np.outer(bs, th)))     # [nodes, nthq]
    quart = (np.abs(Psi) ** 4).mean(axis=1)
    p2 = (np.abs(psi) ** 2)
    dens = gth * gz * (P["alpha"] * p2.sum(0) + 0.5 * P["u"] * quart) + P["C"] * (
        gth * gz * (np.abs(dpsi) ** 2).sum(0) + gz / gth * (((bs[:, None] - n * ath[None, :]) ** 2) * p2).sum(0))
    return float(dens.sum() * L / nodes)


def surface_energy_1d(P, a):
    """Cylinder1D.calc_surface_energy  system_cylinder1D.py:189-194: gamma_eff * int Gth Gz dz + kappa * bending."""
    k, r = P["wavenumber"], P["radius"]
    H0 = P.get("intrinsic_curvature", 0.0)
    gamma_eff = P["gamma"] + P["kappa"] / 2 * H0 ** 2
    if a == 0:
        area = _quad(lambda z: r, k)
    else:
        area = _quad(lambda z: lo.sqrt_g_theta(k, r, a, z) * lo.sqrt_g_z(k, r, a, z), k)
    return gamma_eff * area + P["kappa"] * lo.bending_energy(k, r, a, H0)
