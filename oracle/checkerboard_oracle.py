"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package `cylinder_b200`).

CPU restatement (numpy, fp64) of the PARALLEL schedule the CUDA sweep kernels run: a checkerboard ordering of
the reference's single-site move, with the random draws of each proposal taken from Philox4x32-10 at counter
(site, replica, sweep, tag).  The energy difference of every proposal is the verbatim reference expression
(oracle/lattice_oracle.py:delta_energy, pinned bit-exactly to golden vectors recorded from the reference) merely
evaluated for many sites at once; tests/test_oracle_checkerboard.py checks the vectorised form against the
scalar one.  What is NOT in the reference (and therefore parity-unpinned against it) is the
ordering itself: the reference proposes uniformly random sites one at a time and adapts sigma after each.  The
checkerboard is valid because the dE stencil is nearest-neighbour; tests compare its ensemble observables with
the sequential oracle's.

Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11; known-answer
vectors from the Random123 distribution (kat_vectors).
"""
import math

import numpy as np

from . import lattice_oracle as lo

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
TAG_SITE, TAG_AMP, TAG_INIT = 0, 1, 2

# Random123 kat_vectors, philox4x32-10:  (counter, key, expected)
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x).astype(np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


M2 = np.uint64(0xD256D193)
# Random123 kat_vectors, philox2x32-10:  (counter, key, expected)
PHILOX2_KAT = [
    ((0x00000000, 0x00000000), 0x00000000, (0xff1dae59, 0x6cd10df2)),
    ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
    ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12)),
]


def philox2x32_10(c0, c1, key):
    """Vectorised Philox2x32-10 (the site-move generator, cyl_common.cuh).  Returns 2 uint32 arrays."""
    c0, c1 = (np.asarray(x).astype(np.uint64) & MASK for x in np.broadcast_arrays(c0, c1))
    key = int(key) & 0xFFFFFFFF
    for _ in range(10):
        p = M2 * c0
        c0, c1 = (p >> np.uint64(32)) ^ np.uint64(key) ^ c1, p & MASK
        key = (key + W0) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32)


def site_stream(seed, replica):
    """(key, off) of a replica's site-move stream: two words of Philox4x32-10(replica; seed)  (cyl_common.cuh:site_stream)."""
    x0, x1, _, _ = philox4x32_10(replica & 0xFFFFFFFF, replica >> 32, 0, philox_c3(TAG_SITE, 0), seed & 0xFFFFFFFF, seed >> 32)
    return int(x0), int(x1)


def site_draws(x0, x1):
    """Bit fields of one site draw: x0[31:9] radius, (x0[7:0]:x1[7:0]) angle, x1[31:9] acceptance -> (z_re, z_im, u)."""
    u_rad = ((x0 >> np.uint32(9)).astype(np.float64) + 0.5) / 8388608.0
    ang = ((((x0 & np.uint32(0xFF)).astype(np.uint32) << np.uint32(8)) | (x1 & np.uint32(0xFF))).astype(np.float64) + 0.5) / 65536.0
    u_acc = ((x1 >> np.uint32(9)).astype(np.float64) + 0.5) / 8388608.0
    rad = np.sqrt(-2.0 * np.log(u_rad))
    return rad * np.cos(2.0 * np.pi * ang), rad * np.sin(2.0 * np.pi * ang), u_acc


def philox_c3(tag, sweep):
    return (tag << 24) | ((int(sweep) >> 32) & 0xFFFFFF)


def u01(x):
    """uint32 -> (0,1), 32 bits, midpoint (cyl_common.cuh:u01_f64)."""
    return (x.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def box_muller(x0, x1):
    rad = np.sqrt(-2.0 * np.log(u01(x0)))
    ang = 2.0 * np.pi * u01(x1)
    return rad * np.cos(ang), rad * np.sin(ang)


def axis_colour(i, n):
    i = np.asarray(i)
    c = i & 1
    if n & 1:
        c = np.where(i == n - 1, 2, c)
    return c


def num_colours(Z, T):
    return 3 if ((Z | T) & 1) else 2


def colour_map(Z, T):
    nc = num_colours(Z, T)
    return (axis_colour(np.arange(Z), Z)[:, None] + axis_colour(np.arange(T), T)[None, :]) % nc


def random_initialize(Z, T, mag_max, seed, replica):
    """cyl_lattice_init: psi = rect(U(0,mag_max), U(0,2pi)) at counter (site, replica, 0, INIT)."""
    site = np.arange(Z * T, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(site, replica & 0xFFFFFFFF, 0, philox_c3(TAG_INIT, 0) ^ (replica >> 32), seed & 0xFFFFFFFF,
                                 seed >> 32)
    mag = mag_max * u01(x0)
    ph = 2.0 * u01(x1)
    return (mag * np.cos(np.pi * ph) + 1j * mag * np.sin(np.pi * ph)).reshape(Z, T)


# --------------------------------------------------------------------------------------
# vectorised verbatim dE (same expression tree as lattice_oracle.delta_energy)
# --------------------------------------------------------------------------------------
def _geo(st, a, z):
    ra = st.r / math.sqrt(1 + a ** 2 / 2.0)
    s, c = np.sin(st.k * z), np.cos(st.k * z)
    gth = ra * (1 + a * s)
    gz = np.sqrt(1 + (ra * a * st.k * c) ** 2)
    ath = st.k * ra * a * c / gz
    return gth, gz, ath


def _sq(c):
    return c.real * c.real + c.imag * c.imag


def delta_energy_vec(st, a, I, J, new_values):
    """dE for the moves psi[I,J] -> new_values, all evaluated against the CURRENT psi (valid for a set of
    mutually non-adjacent sites).  I, J in [0, len)."""
    Z, T = st.Z, st.T_len
    lz, lth = st.lz, st.lth
    psi = st.psi
    gth0, gz0, ath0 = _geo(st, a, lz * I)
    gthm, gzm, athm = _geo(st, a, lz * (I - .5))
    gthp, gzp, athp = _geo(st, a, lz * (I + .5))
    if st.variant != lo.RESCALED_0TH:
        A, R, A_nb, R_nb = athm, 1 / gthm ** 2, athp, 1 / gthp ** 2
    else:
        A, R = ath0, 1 / gth0 ** 2
        A_nb, R_nb = A, R
    sg = gth0 * gz0
    v = psi[I, J]
    left_z, right_z = psi[(I - 1) % Z, J], psi[(I + 1) % Z, J]
    left_th, right_th = psi[I, (J - 1) % T], psi[I, (J + 1) % T]
    new_p2, old_p2 = _sq(new_values), _sq(v)
    dE = (st.alpha * new_p2 + st.u / 2 * new_p2 ** 2) - (st.alpha * old_p2 + st.u / 2 * old_p2 ** 2)
    dz_c, dth_c = (v - left_z) / lz, (v - left_th) / lth
    dz_n, dth_n = (right_z - v) / lz, (right_th - v) / lth
    ndz, ndth = (new_values - left_z) / lz, (new_values - left_th) / lth
    nnz, nnth = (right_z - new_values) / lz, (right_th - new_values) / lth
    if st.variant == lo.SIMPLE:
        d_z = (_sq(ndz) - _sq(dz_c)) / gzm ** 2
        d_th = (_sq(ndth) - _sq(dth_c)) * R / gthm ** 2
        dn_z = (_sq(nnz) - _sq(dz_n)) / gzp ** 2
        dn_th = (_sq(nnth) - _sq(dth_n)) * R_nb / gthp ** 2
    elif st.variant == lo.RESCALED_1ST:
        d_z = (_sq(ndz) - _sq(dz_c)) / gzm ** 2
        d_th = (_sq(ndth) - _sq(dth_c)) * R
        dn_z = (_sq(nnz) - _sq(dz_n)) / gzp ** 2
        dn_th = (_sq(nnth) - _sq(dth_n)) * R_nb
    elif st.variant == lo.DECOUPLED:
        d_z = (_sq(ndz) - _sq(dz_c))
        d_th = (_sq(ndth) - _sq(dth_c)) * R
        dn_z = (_sq(nnz) - _sq(dz_n))
        dn_th = (_sq(nnth) - _sq(dth_n)) * R_nb
    else:
        d_z = (_sq(ndz) - _sq(dz_c)) / gzm ** 2 * (gthm * gzm) / sg
        d_th = (_sq(ndth) - _sq(dth_c)) * R
        dn_z = (_sq(nnz) - _sq(dz_n)) / gzp ** 2 * (gthp * gzp) / sg
        dn_th = (_sq(nnth) - _sq(dth_n)) * R
    dE = dE + st.C * (d_z + d_th)
    dE = dE + st.C * (dn_z + dn_th)
    if st.variant == lo.RESCALED_1ST:            # cross terms carry 1/sqrt_g_theta of the interstitial row (1st:190-203)
        dth_c, ndth, dth_n, nnth = dth_c / gthm, ndth / gthm, dth_n / gthp, nnth / gthp
    old_cross = (A * v.conj() * dth_c).imag
    new_cross = (A * new_values.conj() * ndth).imag
    dE = dE + st.C * 2 * st.n * R * (new_cross - old_cross)
    old_nb = (A_nb * right_th.conj() * dth_n).imag
    new_nb = (A_nb * right_th.conj() * nnth).imag
    dE = dE + st.C * 2 * st.n * R_nb * (new_nb - old_nb)
    dE = dE + st.Cn2 * R * A * A * (new_p2 - old_p2)
    dE = dE * sg
    dE = dE * (lz * lth)
    return dE, sg


def rm_batch_update(sigma, step_counter, n, k):
    """cyl_site.cuh:rm_batch_update -- the reference's per-proposal Robbins-Monro rule (On_lattice_simple.py:688-696,
    m = 1, p* = .5, ratio = 4) applied as a product with the step factor frozen at mid-batch."""
    f = max(step_counter + 0.5 * n, 200.0)
    return sigma * math.exp(k * math.log1p(2.0 / f) + (n - k) * math.log1p(-2.0 / f))


def sweep(st, a, seed, replica, sweep_index, adapt=True, return_details=False):
    """One checkerboard sweep of `st` (lattice_oracle.LatticeState) in place.  Returns the accept count."""
    Z, T = st.Z, st.T_len
    cmap = colour_map(Z, T)
    key, off = site_stream(seed, replica)
    nacc = 0
    details = {"dE": np.zeros((Z, T)), "accept": np.zeros((Z, T), bool), "u": np.zeros((Z, T))}
    for colour in range(num_colours(Z, T)):
        I, J = np.nonzero(cmap == colour)
        site = (I * T + J).astype(np.uint64)
        x0, x1 = philox2x32_10(site, (sweep_index + off) & 0xFFFFFFFF, key)
        zre, zim, u = site_draws(x0, x1)
        v = st.psi[I, J]
        if st.variant == lo.SIMPLE:
            w = st.sigma
        else:
            gth0, gz0, _ = _geo(st, a, st.lz * I)
            w = st.sigma * (gth0 * gz0)
        new_values = (v.real + w * zre) + 1j * (v.imag + w * zim)
        dE, _ = delta_energy_vec(st, a, I, J, new_values)
        if st.temperature > 0:
            with np.errstate(over="ignore"):
                acc = (dE <= 0) | (u <= np.exp(-dE / st.temperature))
        else:
            acc = dE <= 0
        st.psi[I[acc], J[acc]] = new_values[acc]
        nacc += int(acc.sum())
        details["dE"][I, J] = dE
        details["accept"][I, J] = acc
        details["u"][I, J] = u
    st.rebuild_caches()
    n = Z * T
    if adapt:
        st.sigma = rm_batch_update(st.sigma, st.step_counter, n, nacc)
    st.step_counter += n
    st.accepted += nacc
    return (nacc, details) if return_details else nacc


class AmpState:
    """Scalars of the global amplitude move (engine.step_real_group restated; SURVEY.md B.2/B.3, parity unpinned)."""

    def __init__(self, amplitude=0.0, sigma_amp=.005, target=.3, bound=.8):
        self.a, self.sigma, self.target, self.bound = amplitude, sigma_amp, target, bound
        self.counter, self.proposed, self.accepted = 0, 0, 0


def amplitude_move(st, am, P, seed, replica, sweep_index, adapt=True):
    """One global move a' = a + sigma_amp * N(0,1) on E_surf(a') + E_field(a', a_ref=a); P = dict of Cylinder params."""
    x0, x1, x2, _ = philox4x32_10(0, replica & 0xFFFFFFFF, sweep_index & 0xFFFFFFFF, philox_c3(TAG_AMP, sweep_index),
                                  seed & 0xFFFFFFFF, seed >> 32)
    zre, _ = box_muller(x0, x1)
    a_new = am.a + am.sigma * float(zre)
    accept = False
    if abs(a_new) < am.bound:
        e_cur = lo.surface_field_energy(st, am.a, am.a) + lo.surface_energy(st.k, st.r, P["gamma"], P["kappa"],
                                                                           P["intrinsic_curvature"], am.a)
        e_new = lo.surface_field_energy(st, a_new, am.a) + lo.surface_energy(st.k, st.r, P["gamma"], P["kappa"],
                                                                            P["intrinsic_curvature"], a_new)
        accept, _ = lo.metropolis_decision(st.temperature, e_cur, e_new, lambda: float(u01(x2)))
    am.counter += 1
    am.proposed += 1
    if adapt:
        ratio = 1.0 / (am.target * (1 - am.target))
        f = max(am.counter, 200.0)
        c = am.sigma * ratio
        if accept:
            am.sigma += c * (1 - am.target) / f
        else:
            am.sigma -= c * am.target / f
    if accept:
        am.a = a_new
        am.accepted += 1
    return accept
