"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy / scalar Python, fp64) of the
reference's lattice hot path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this; the product package `cylinder_b200` never does.

Parity status: PINNED.  Every function here is checked against golden vectors generated from the
unmodified reference (tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py)
and, when /root/reference is present, against the reference itself (tests/test_oracle_vs_reference.py).
Exception: the Robbins-Monro / accept rule of the un-vendored `metropolisengine` package is pinned only
by the inline copy in On_lattice_simple.py:95-100,688-696 and by five known answers.

Reference files cited as  <file>:<line>  relative to /root/reference/surfaces_and_fields/ unless noted.
"""
import math

import numpy as np

SIMPLE = 0             # On_lattice_simple.py
RESCALED_0TH = 1       # On_lattice_rescaled_0thorder.py
RESCALED_1ST = 2       # On_lattice_rescaled_1storder.py   ("1st:" in the citations below)
DECOUPLED = 3          # On_lattice_decoupled.py           ("dec:")
VARIANTS = {"simple": SIMPLE, "rescaled_0thorder": RESCALED_0TH, "rescaled_1storder": RESCALED_1ST, "decoupled": DECOUPLED}
# defaults that differ by variant: initial |psi| bound (:164 / dec:157) and initial sampling width (:101 / dec:99)
INIT_MAGNITUDE = {SIMPLE: .1, RESCALED_0TH: .1, RESCALED_1ST: .1, DECOUPLED: .001}
INIT_SIGMA = {SIMPLE: .005, RESCALED_0TH: .005, RESCALED_1ST: .005, DECOUPLED: .0001}


# --------------------------------------------------------------------------------------
# surface geometry  (system_cylinder.py:25-41)
# --------------------------------------------------------------------------------------
def radius_rescaled(r, a):                                   # system_cylinder.py:37
    return r / math.sqrt(1 + a ** 2 / 2.0)


def sqrt_g_theta(k, r, a, z):                                # system_cylinder.py:28
    return radius_rescaled(r, a) * (1 + a * math.sin(k * z))


def g_theta(k, r, a, z):                                     # system_cylinder.py:25
    return radius_rescaled(r, a) ** 2 * (1 + a * math.sin(k * z)) ** 2


def g_z(k, r, a, z):                                         # system_cylinder.py:31
    return 1 + (radius_rescaled(r, a) * a * k * math.cos(k * z)) ** 2


def sqrt_g_z(k, r, a, z):                                    # system_cylinder.py:34
    return math.sqrt(1 + (radius_rescaled(r, a) * a * k * math.cos(k * z)) ** 2)


def A_theta(k, r, a, z):                                     # system_cylinder.py:40
    return k * radius_rescaled(r, a) * a * math.cos(k * z) / sqrt_g_z(k, r, a, z)


# --------------------------------------------------------------------------------------
# bare surface energy  (system_cylinder.py:91-124)
# --------------------------------------------------------------------------------------
def area_integral_0(k, r, a):
    """4 r(a)/k * E(m = -(a r(a) k)^2)   system_cylinder.py:91-103 (scipy ellipe takes parameter m)."""
    import scipy.special
    ra = radius_rescaled(r, a)
    return 4.0 * ra / k * scipy.special.ellipe(-(a * ra * k) ** 2)


def bending_energy(k, r, a, H0):
    """system_cylinder.py:108-117 -- two adaptive quad integrals + two closed forms."""
    import scipy.integrate as integrate
    ra = radius_rescaled(r, a)

    def kzz(z):                                              # system_cylinder.py:52-55
        return (ra * a * k ** 2 * math.sin(k * z)) ** 2 * sqrt_g_theta(k, r, a, z) / sqrt_g_z(k, r, a, z) ** 5

    def kthth(z):                                            # system_cylinder.py:46-50
        return 1 / (sqrt_g_theta(k, r, a, z) * sqrt_g_z(k, r, a, z))

    Kzz, _ = integrate.quad(kzz, 0, 2 * math.pi / k)
    Kthth, _ = integrate.quad(kthth, 0, 2 * math.pi / k)
    Kzz_lin = -2 * math.pi / k * (-1 + math.sqrt(1 + (a * k * ra) ** 2))      # :115
    Kthth_lin = -2 * math.pi / k                                              # :116
    return Kzz + Kthth - 2 * H0 * (Kzz_lin + Kthth_lin)


def surface_energy(k, r, gamma, kappa, H0, a):
    """Cylinder.calc_surface_energy  system_cylinder.py:119-124, effective_gamma :22."""
    gamma_eff = gamma + kappa / 2 * H0 ** 2
    return 2 * math.pi * (gamma_eff * area_integral_0(k, r, a) + .5 * kappa * bending_energy(k, r, a, H0))


def surface_energy_quadrature(k, r, gamma, kappa, H0, a, nodes=256):
    """Same number by the scheme the GPU uses: E(m) by AGM, integrals by periodic
    trapezoid.  Independent restatement used to bound the quadrature error in tests."""
    ra = radius_rescaled(r, a)
    m = -(a * ra * k) ** 2
    # complete elliptic integral of the 2nd kind, parameter m<=0, by arithmetic-geometric mean
    a0, b0 = 1.0, math.sqrt(1.0 - m)
    c2sum, pw = 0.5 * (a0 * a0 - b0 * b0), 0.5        # sum 2^(n-1) c_n^2 with c_0^2 = a0^2-b0^2 = m
    for _ in range(40):
        an, bn = 0.5 * (a0 + b0), math.sqrt(a0 * b0)
        cn = 0.5 * (a0 - b0)
        pw *= 2.0
        c2sum += pw * cn * cn
        a0, b0 = an, bn
        if abs(cn) < 1e-17 * abs(an):
            break
    ellipe = math.pi / (2.0 * a0) * (1.0 - c2sum)
    area = 4.0 * ra / k * ellipe
    L = 2 * math.pi / k
    s_zz = s_tt = 0.0
    for i in range(nodes):
        z = L * i / nodes
        gt, gz = sqrt_g_theta(k, r, a, z), sqrt_g_z(k, r, a, z)
        s_zz += (ra * a * k ** 2 * math.sin(k * z)) ** 2 * gt / gz ** 5
        s_tt += 1.0 / (gt * gz)
    Kzz, Kthth = s_zz * L / nodes, s_tt * L / nodes
    Kzz_lin = -2 * math.pi / k * (-1 + math.sqrt(1 + (a * k * ra) ** 2))
    Kthth_lin = -2 * math.pi / k
    bend = Kzz + Kthth - 2 * H0 * (Kzz_lin + Kthth_lin)
    gamma_eff = gamma + kappa / 2 * H0 ** 2
    return 2 * math.pi * (gamma_eff * area + .5 * kappa * bend)


# --------------------------------------------------------------------------------------
# lattice state
# --------------------------------------------------------------------------------------
def squared(c):                                              # On_lattice_simple.py:105-106
    return abs(c * c.conjugate())


class LatticeState:
    """Arrays + scalars of reference `Lattice` (On_lattice_simple.py:18-102) that the hot path touches."""

    def __init__(self, psi, k, r, alpha, u, C, n, T, variant=SIMPLE, dims0=None,
                 sigma=.005, step_counter=0):
        psi = np.array(psi, dtype=complex)
        self.Z, self.T_len = psi.shape
        self.k, self.r, self.alpha, self.u, self.C, self.n = k, r, alpha, u, C, n
        self.Cn2 = C * n ** 2                                # :32
        self.temperature = T
        self.variant = variant
        self.lz = (2 * math.pi / k) / self.Z                 # :38,:47
        self.lth = (2 * math.pi * r) / self.T_len            # :39,:49
        self.psi = psi
        self.rebuild_caches()
        # Robbins-Monro constants :95-101
        self.target = .5
        self.m = 1
        self.ratio = rm_ratio(self.target, self.m)
        self.sigma = sigma
        self.step_counter = step_counter
        self.accepted = 0
        self.bump_sigma = .005                                 # :102 bump_sampling_width
        self.bump_step_counter = 0

    def rebuild_caches(self):
        """psi_squared, dz, dth as in random_initialize :159-182 (left differences, periodic)."""
        p = self.psi
        # |psi|^2 as re*re + im*im in separate roundings == abs(c*c.conjugate()) of scalar Python complex
        # (:105-106,:167); numpy's vectorised complex multiply may fuse and differ in the last bit.
        self.psi2 = p.real * p.real + p.imag * p.imag
        self.dz = (p - np.roll(p, 1, axis=0)) / self.lz
        self.dth = (p - np.roll(p, 1, axis=1)) / self.lth


def lattice_dims(dims, k):
    """z_len = int(round(dims[0]/k)), th_len = dims[1]   On_lattice_simple.py:41-42."""
    return int(round(dims[0] / k)), dims[1]


def rm_ratio(target, m):
    """On_lattice_simple.py:96-100 (same rule as Garthwaite, Fan & Sisson arXiv:1006.3690)."""
    import scipy.stats
    ppf_alpha = -1 * scipy.stats.norm.ppf(target / 2)
    return ((1 - (1 / m)) * math.sqrt(2 * math.pi) * math.exp(ppf_alpha ** 2 / 2) / 2 * ppf_alpha
            + 1 / (m * target * (1 - target)))


def update_sigma(st, accept):
    """On_lattice_simple.py:688-696."""
    st.step_counter += 1
    f = max(st.step_counter / st.m, 200)
    c = st.sigma * st.ratio
    if accept:
        st.sigma += c * (1 - st.target) / f
    else:
        st.sigma -= c * st.target / f
    assert st.sigma > 0


def metropolis_decision(T, old, new, draw_uniform):
    """Engine accept rule (predecessor metropolis_engine.py:226-246).
    draw_uniform() is called only when diff>0 and T!=0.  Returns (accept, u_or_None)."""
    diff = new - old
    if diff <= 0:
        return True, None
    if T == 0:
        return False, None
    u = draw_uniform()
    return u <= math.exp(-diff / T), u


# --------------------------------------------------------------------------------------
# single-site proposal: local energy difference
# --------------------------------------------------------------------------------------
def proposal_sigma(st, a, iz):
    """Std-dev of the Gaussian increment: sigma (simple :613) or sigma*sqrt_g(z_i) (0th:135-136, 1st:153-154, dec:274-275)."""
    if st.variant == SIMPLE:
        return st.sigma
    z = st.lz * iz
    return st.sigma * (sqrt_g_theta(st.k, st.r, a, z) * sqrt_g_z(st.k, st.r, a, z))


def delta_energy(st, a, iz, ith, new_value):
    """dE of psi[iz,ith] -> new_value, exactly the quantity the reference passes to
    metropolis_decision(0, dE).  simple: On_lattice_simple.py:598-669;
    rescaled_0thorder: On_lattice_rescaled_0thorder.py:116-192.
    iz, ith follow the reference's convention: any index in [-1, len-2] (negative wraps)."""
    k, r = st.k, st.r
    lz, lth = st.lz, st.lth
    psi, dz, dth = st.psi, st.dz, st.dth
    z_loc = lz * iz
    z_m = lz * (iz - .5)
    z_p = lz * (iz + .5)
    if st.variant in (RESCALED_1ST, DECOUPLED):
        return _delta_energy_interstitial(st, a, iz, ith, new_value)
    if st.variant == SIMPLE:
        A = A_theta(k, r, a, z_m)                                            # :603
        R = 1 / g_theta(k, r, a, z_m)                                        # :604
        A_nb = A_theta(k, r, a, z_p)                                         # :605
        R_nb = 1 / g_theta(k, r, a, z_p)                                     # :606
    else:
        A = A_theta(k, r, a, z_loc)                                          # 0th:122
        R = 1 / g_theta(k, r, a, z_loc)                                      # 0th:123
        A_nb, R_nb = A, R                                                    # 0th:164,184-188 use same-row values
    sg = sqrt_g_theta(k, r, a, z_loc) * sqrt_g_z(k, r, a, z_loc)             # :608 / 0th:127
    value = psi[iz, ith]
    new_p2 = squared(new_value)                                              # :618
    old_p2 = st.psi2[iz, ith]                                                # :620
    dE = (st.alpha * new_p2 + st.u / 2 * new_p2 ** 2) - (st.alpha * old_p2 + st.u / 2 * old_p2 ** 2)   # :621-623
    left_z, left_th = psi[iz - 1, ith], psi[iz, ith - 1]                     # :628-629
    ndz = (new_value - left_z) / lz                                          # :630
    ndth = (new_value - left_th) / lth                                       # :631
    if st.variant == SIMPLE:
        d_z = (squared(ndz) - squared(dz[iz, ith])) / sqrt_g_z(k, r, a, z_m) ** 2                     # :633
        d_th = (squared(ndth) - squared(dth[iz, ith])) * R / sqrt_g_theta(k, r, a, z_m) ** 2          # :634
    else:
        sg_m = sqrt_g_theta(k, r, a, z_m) * sqrt_g_z(k, r, a, z_m)                                    # 0th:128
        d_z = (squared(ndz) - squared(dz[iz, ith])) / (sqrt_g_z(k, r, a, z_m)) ** 2 * sg_m / sg       # 0th:153
        d_th = (squared(ndth) - squared(dth[iz, ith])) * R                                            # 0th:154
    dE += st.C * (d_z + d_th)                                                # :636
    right_z, right_th = psi[iz + 1, ith], psi[iz, ith + 1]                   # :639-640
    nnz = (right_z - new_value) / lz                                         # :641
    nnth = (right_th - new_value) / lth                                      # :642
    if st.variant == SIMPLE:
        dn_z = (squared(nnz) - squared(dz[iz + 1, ith])) / sqrt_g_z(k, r, a, z_p) ** 2                # :643
        dn_th = (squared(nnth) - squared(dth[iz, ith + 1])) * R_nb / sqrt_g_theta(k, r, a, z_p) ** 2  # :644
    else:
        sg_p = sqrt_g_theta(k, r, a, z_p) * sqrt_g_z(k, r, a, z_p)                                    # 0th:130
        dn_z = (squared(nnz) - squared(dz[iz + 1, ith])) / sqrt_g_z(k, r, a, z_p) ** 2 * sg_p / sg    # 0th:163
        dn_th = (squared(nnth) - squared(dth[iz, ith + 1])) * R                                       # 0th:164
    dE += st.C * (dn_z + dn_th)                                              # :646
    old_cross = (A * value.conjugate() * dth[iz, ith]).imag                  # :651-652 (A real => conj no-op)
    new_cross = (A * new_value.conjugate() * ndth).imag                      # :654-655
    dE += st.C * 2 * st.n * R * (new_cross - old_cross)                      # :656
    if st.variant == SIMPLE:
        old_nb = (A_nb * right_th.conjugate() * dth[iz, ith + 1]).imag       # :659-660
        new_nb = (A_nb * right_th.conjugate() * nnth).imag                   # :661-662
        dE += st.C * 2 * st.n * R_nb * (new_nb - old_nb)                     # :663
        dE += st.Cn2 * R * squared(complex(A)) * (squared(new_value) - squared(value))   # :665-666
    else:
        dE += st.Cn2 * R * squared(complex(A)) * (squared(new_value) - squared(value))   # 0th:179-180
        old_nb = (A * right_th.conjugate() * dth[iz, ith + 1]).imag          # 0th:184-185
        new_nb = (A * right_th.conjugate() * nnth).imag                      # 0th:186-187
        dE += st.C * 2 * st.n * R * (new_nb - old_nb)                        # 0th:188
    dE *= sg                                                                 # :668
    dE *= lz * lth                                                           # :669
    return dE, (new_p2, ndz, nnz, ndth, nnth)


def _delta_energy_interstitial(st, a, iz, ith, new_value):
    """rescaled_1storder: On_lattice_rescaled_1storder.py:138-210; decoupled: On_lattice_decoupled.py:259-330.
    Both take A_theta and the index raise at the interstitial rows z -+ lz/2 (like `simple`); they differ from it in
    the metric factors of the gradient and cross terms."""
    k, r = st.k, st.r
    lz, lth = st.lz, st.lth
    psi, dz, dth = st.psi, st.dz, st.dth
    first = st.variant == RESCALED_1ST
    z_loc = lz * iz                                                          # 1st:139 / dec:260
    z_m = lz * (iz - .5)                                                     # 1st:140 / dec:261
    z_p = lz * (iz + .5)                                                     # 1st:141 / dec:262
    A = A_theta(k, r, a, z_m)                                                # 1st:143 / dec:264
    R = 1 / g_theta(k, r, a, z_m)                                            # 1st:144 / dec:265
    A_nb = A_theta(k, r, a, z_p)                                             # 1st:145 / dec:266
    R_nb = 1 / g_theta(k, r, a, z_p)                                         # 1st:146 / dec:267
    sg = sqrt_g_theta(k, r, a, z_loc) * sqrt_g_z(k, r, a, z_loc)             # 1st:148 / dec:269
    value = psi[iz, ith]
    new_p2 = squared(new_value)                                              # 1st:158 / dec:279
    old_p2 = st.psi2[iz, ith]                                                # 1st:160 / dec:281
    dE = (st.alpha * new_p2 + st.u / 2 * new_p2 ** 2) - (st.alpha * old_p2 + st.u / 2 * old_p2 ** 2)   # 1st:161-163 / dec:282-284
    left_z, left_th = psi[iz - 1, ith], psi[iz, ith - 1]                     # 1st:166-167 / dec:289-290
    ndz = (new_value - left_z) / lz                                          # 1st:168 / dec:291
    ndth = (new_value - left_th) / lth                                       # 1st:169 / dec:292
    if first:
        d_z = (squared(ndz) - squared(dz[iz, ith])) / (sqrt_g_z(k, r, a, z_m)) ** 2               # 1st:171
    else:
        d_z = (squared(ndz) - squared(dz[iz, ith]))                                               # dec:294
    d_th = (squared(ndth) - squared(dth[iz, ith])) * R                       # 1st:172 / dec:295
    dE += st.C * (d_z + d_th)                                                # 1st:174 / dec:297
    right_z, right_th = psi[iz + 1, ith], psi[iz, ith + 1]                   # 1st:177-178 / dec:300-301
    nnz = (right_z - new_value) / lz                                         # 1st:179 / dec:302
    nnth = (right_th - new_value) / lth                                      # 1st:180 / dec:303
    if first:
        dn_z = (squared(nnz) - squared(dz[iz + 1, ith])) / sqrt_g_z(k, r, a, z_p) ** 2            # 1st:181
    else:
        dn_z = (squared(nnz) - squared(dz[iz + 1, ith]))                                          # dec:304
    dn_th = (squared(nnth) - squared(dth[iz, ith + 1])) * R_nb               # 1st:182 / dec:305
    dE += st.C * (dn_z + dn_th)                                              # 1st:184 / dec:307
    if first:
        gth_m, gth_p = sqrt_g_theta(k, r, a, z_m), sqrt_g_theta(k, r, a, z_p)
        old_cross = (A * value.conjugate() * dth[iz, ith] / gth_m).imag      # 1st:190-191
        new_cross = (A * new_value.conjugate() * ndth / gth_m).imag          # 1st:194-195
    else:
        old_cross = (A * value.conjugate() * dth[iz, ith]).imag              # dec:312-313
        new_cross = (A * new_value.conjugate() * ndth).imag                  # dec:315-316
    dE += st.C * 2 * st.n * R * (new_cross - old_cross)                      # 1st:196 / dec:317
    if first:
        old_nb = (A_nb * right_th.conjugate() * dth[iz, ith + 1] / gth_p).imag   # 1st:199-200
        new_nb = (A_nb * right_th.conjugate() * nnth / gth_p).imag               # 1st:201-202
    else:
        old_nb = (A_nb * right_th.conjugate() * dth[iz, ith + 1]).imag       # dec:320-321
        new_nb = (A_nb * right_th.conjugate() * nnth).imag                   # dec:322-323
    dE += st.C * 2 * st.n * R_nb * (new_nb - old_nb)                         # 1st:203 / dec:324
    dE += st.Cn2 * R * squared(complex(A)) * (squared(new_value) - squared(value))   # 1st:205-206 / dec:326-327
    dE *= sg                                                                 # 1st:209 / dec:329
    dE *= lz * lth                                                           # 1st:210 / dec:330
    return dE, (new_p2, ndz, nnz, ndth, nnth)


def commit(st, iz, ith, new_value, derived):
    """On accept: On_lattice_simple.py:671-685."""
    new_p2, ndz, nnz, ndth, nnth = derived
    st.psi[iz, ith] = new_value
    st.psi2[iz, ith] = new_p2
    st.dz[iz, ith] = ndz
    st.dz[iz + 1, ith] = nnz
    st.dth[iz, ith] = ndth
    st.dth[iz, ith + 1] = nnth
    st.accepted += 1


def step_lattice_loc(st, a, iz, ith, z_re, z_im, uniform):
    """One proposal with supplied unit normals (z_re drawn first, :617) and a uniform used only
    if dE>0 and T!=0.  `uniform` is a float or a zero-arg callable.  Returns (accept, dE, u_used)."""
    s = proposal_sigma(st, a, iz)
    v = st.psi[iz, ith]
    new_value = complex(v.real + (0.0 + z_re * s), v.imag + (0.0 + z_im * s))    # :617 with gauss = mu + z*sigma
    dE, derived = delta_energy(st, a, iz, ith, new_value)
    draw = uniform if callable(uniform) else (lambda: uniform)
    accept, u_used = metropolis_decision(st.temperature, 0, dE, draw)            # :670
    if accept:
        commit(st, iz, ith, new_value, derived)
    update_sigma(st, accept)                                                     # :686
    return accept, dE, u_used


def replay_tape(st, a, iz, ith, zre, zim, u):
    """Replay n proposals; u[i] is NaN where the reference drew no uniform.  Returns arrays
    (accept, dE, sigma_after)."""
    n = len(iz)
    acc = np.zeros(n, dtype=np.uint8)
    dE = np.zeros(n)
    sig = np.zeros(n)
    for i in range(n):
        a_i, d_i, _ = step_lattice_loc(st, a, int(iz[i]), int(ith[i]), float(zre[i]), float(zim[i]), float(u[i]))
        acc[i], dE[i], sig[i] = a_i, d_i, st.sigma
    return acc, dE, sig


# --------------------------------------------------------------------------------------
# whole-lattice field energy
# --------------------------------------------------------------------------------------
def background_energy(st, cell_dims):
    """On_lattice_rescaled_0thorder.py:66-92 (ordered_reference=True)."""
    area = cell_dims[0] * cell_dims[1]
    ans = -st.temperature
    ordered = .5 * st.alpha ** 2 / st.u if st.alpha < 0 else 0
    return ans / area + ordered


def background_energy_1st(st, cell_dims):
    """On_lattice_rescaled_1storder.py:66-114 (ordered_reference=True): the 0th-order background plus an O(u)
    fluctuation correction; the ValueError branches of the reference (log of a non-positive number) are
    restated as explicit sign tests -- math.log raises exactly for arguments <= 0."""
    area = cell_dims[0] * cell_dims[1]                                        # 1st:69
    ans = -st.temperature                                                     # 1st:86
    area_factor = 2 / math.sqrt(math.pi)                                      # 1st:34
    upper = 1 * area_factor                                                   # 1st:89
    lower = (1 / max((st.Z, st.T_len))) * area_factor                         # 1st:90
    alpha_cell = st.alpha * area                                              # 1st:92
    hi = st.C * upper ** 2 + alpha_cell
    lo = st.C * lower ** 2 + alpha_cell
    if hi > 0 and lo > 0:
        integral = math.pi / st.C * (math.log(hi) - math.log(lo))             # 1st:96
    elif hi > 0:                                                              # 1st:98-102
        lower = math.sqrt(abs(alpha_cell) / st.C) + .0000001
        integral = math.pi / st.C * (math.log(st.C * upper ** 2 + alpha_cell) - math.log(st.C * lower ** 2 + alpha_cell))
    else:
        integral = 0                                                          # 1st:103-105
    correction = st.u * st.temperature ** 3 * 4 * integral ** 2               # 1st:106
    correction /= st.Z * st.T_len                                             # 1st:107
    ans += correction                                                         # 1st:108
    ordered = .5 * st.alpha ** 2 / st.u if st.alpha < 0 else 0                # 1st:110-113
    return ans / area + ordered                                               # 1st:114


def _decoupled_field_energy(st, a, psi):
    """On_lattice_decoupled.py:179-202: only the C n^2 |A_theta psi|^2 coupling enters an amplitude move."""
    k, r = st.k, st.r
    energy = 0
    for i in range(st.Z):
        z_loc = i * st.lz                                                     # dec:188
        z_m = (i - .5) * st.lz                                                # dec:189
        col_raise = sqrt_g_z(k, r, a, z_loc) / sqrt_g_theta(k, r, a, z_m)     # dec:190
        col_A = A_theta(k, r, a, z_m)                                         # dec:191
        e = st.Cn2 * sum(np.abs(psi[i] * psi[i].conj())) * squared(complex(col_A)) * col_raise   # dec:198
        energy += e
    return energy * st.lz * st.lth                                            # dec:202


def surface_field_energy(st, a, a_ref=None, psi=None, psi2=None, dz=None, dth=None):
    """simple: On_lattice_simple.py:185-213 (ignores a_ref);
    rescaled_0thorder: On_lattice_rescaled_0thorder.py:247-278 (mixes a_ref and a in the metric);
    rescaled_1storder: On_lattice_rescaled_1storder.py:265-296 (as 0th order, A_theta at z - lz/2, 1st-order background);
    decoupled: On_lattice_decoupled.py:179-202 (takes the amplitude only)."""
    if st.variant == DECOUPLED:
        return _decoupled_field_energy(st, a, st.psi if psi is None else psi)
    psi = st.psi if psi is None else psi
    psi2 = st.psi2 if psi2 is None else psi2
    dz = st.dz if dz is None else dz
    dth = st.dth if dth is None else dth
    if a_ref is None:
        a_ref = a
    k, r = st.k, st.r
    energy = 0
    for i in range(st.Z):
        z_loc = i * st.lz
        z_m = (i - .5) * st.lz
        if st.variant == SIMPLE:
            z_spacing = sqrt_g_z(k, r, a, z_m)                                # :191
            th_spacing = sqrt_g_theta(k, r, a, z_loc)                         # :192
            col_sqrtg = z_spacing * th_spacing                                # :193
            col_raise = sqrt_g_z(k, r, a, z_loc) / sqrt_g_theta(k, r, a, z_m)  # :194
            col_A = A_theta(k, r, a, z_m)                                     # :195
        else:
            z_spacing = sqrt_g_z(k, r, a_ref, z_loc)                          # 0th:253
            th_spacing = sqrt_g_theta(k, r, a_ref, z_loc)                     # 0th:254
            col_sqrtg = sqrt_g_z(k, r, a_ref, z_loc) * sqrt_g_theta(k, r, a, z_loc)   # 0th:255
            col_raise = col_sqrtg / th_spacing ** 2                           # 0th:256
            col_A = A_theta(k, r, a, z_loc if st.variant == RESCALED_0TH else z_m)   # 0th:257 / 1st:275
        p2 = psi2[i]
        e = (st.alpha * sum(p2) + st.u / 2 * sum(p2 ** 2)) * col_sqrtg        # :199
        dz_col = dz[i] / z_spacing                                            # :200
        e += st.C * sum(np.abs(dz_col * dz_col.conj())) * col_sqrtg           # :202
        dth_col = dth[i]                                                      # :203
        e += st.C * sum(np.abs(dth_col * dth_col.conj())) * col_raise         # :206
        e += st.C * st.n * 2 * sum((col_A * psi[i].conj() * dth_col).imag) * col_raise    # :208
        e += st.Cn2 * sum(np.abs(psi[i] * psi[i].conj())) * squared(complex(col_A)) * col_raise   # :210
        if st.variant != SIMPLE:
            cell = (st.lz * sqrt_g_z(k, r, a, z_loc), st.lth * sqrt_g_theta(k, r, a, z_loc))       # 0th:261 / 1st:279
            bg = background_energy(st, cell) if st.variant == RESCALED_0TH else background_energy_1st(st, cell)
            e += st.T_len * bg * col_sqrtg                                                         # 0th:275 / 1st:293
        energy += e
    return energy * st.lz * st.lth                                            # :213


def row_moments(st):
    """S1..S5 of SURVEY.md A.3 per z-row: sum|psi|^2, sum|psi|^4, sum|dz|^2, sum|dth|^2, sum Im(conj(psi) dth)."""
    p = st.psi
    p2 = np.abs(p * p.conj())
    S = np.zeros((st.Z, 5))
    S[:, 0] = p2.sum(1)
    S[:, 1] = (p2 ** 2).sum(1)
    S[:, 2] = np.abs(st.dz * st.dz.conj()).sum(1)
    S[:, 3] = np.abs(st.dth * st.dth.conj()).sum(1)
    S[:, 4] = (p.conj() * st.dth).imag.sum(1)
    return S


# --------------------------------------------------------------------------------------
# observables  (On_lattice_simple.py:121-131)
# --------------------------------------------------------------------------------------
def measure_profiles(psi):
    """(sum_theta psi, sum_theta |psi|) per z-row  -- Lattice.measure :123,:128."""
    return psi.sum(axis=1), np.abs(psi).sum(axis=1)


# --------------------------------------------------------------------------------------
# whole-lattice proposals (On_lattice_simple.py:247-382)
# --------------------------------------------------------------------------------------
def update_bump_sigma(st, accept):
    """On_lattice_simple.py:700-708 -- note the step factor uses step_counter, not bump_step_counter (as committed)."""
    st.bump_step_counter += 1
    f = max(st.step_counter / st.m, 200)
    c = st.bump_sigma * st.ratio
    if accept:
        st.bump_sigma += c * (1 - st.target) / f
    else:
        st.bump_sigma -= c * st.target / f
    assert st.bump_sigma > 0


def _decide_and_commit(st, old_energy, new_energy, new_psi, uniform):
    draw = uniform if callable(uniform) else (lambda: uniform)
    accept, u_used = metropolis_decision(st.temperature, old_energy, new_energy, draw)
    if accept:
        st.psi = new_psi
        st.rebuild_caches()
    return accept, u_used


def step_lattice_all(st, a, old_energy, z_re, z_im, uniform):
    """Lattice.step_lattice_all :247-259: the same complex number added to every site.  dz / dth do not change (:251); the
    reference's psi_squared of the proposal is the complex product psi*conj(psi), whose real part is what is restated
    here (its imaginary part is rounding noise of order 1e-19).  Returns (accept, new_energy, u_used)."""
    addition = (0.0 + z_re * st.sigma) + (0.0 + z_im * st.sigma) * 1j                      # :248 (gauss = mu + z*sigma)
    new = st.psi + np.full((st.Z, st.T_len), addition)                                      # :249-250
    new_p2 = new.real * new.real + new.imag * new.imag                                      # :251 (real part)
    new_energy = surface_field_energy(st, a, a, new, new_p2, st.dz, st.dth)                 # :252
    accept, u_used = _decide_and_commit(st, old_energy, new_energy, new, uniform)           # :254-260
    update_sigma(st, accept)                                                                # :261
    return accept, new_energy, u_used


def wave_array(base, wavevector, Z, T):
    """Lattice.wave_array :267-271."""
    qz, qth = wavevector
    i, j = np.meshgrid(np.arange(Z, dtype=float), np.arange(T, dtype=float), indexing="ij")
    return base * np.exp(1j * (i * qz / Z + j * qth / T))


def _energy_of(st, a, psi):
    """surface_field_energy of a whole proposed field with its own derived arrays."""
    p2 = psi.real * psi.real + psi.imag * psi.imag
    dz = (psi - np.roll(psi, 1, axis=0)) / st.lz
    dth = (psi - np.roll(psi, 1, axis=1)) / st.lth
    return surface_field_energy(st, a, a, psi, p2, dz, dth)


def step_lattice_wavevector(st, a, old_energy, wavevector, z_re, z_im, uniform):
    """Lattice.step_lattice_wavevector :273-300 with COMPLEX derivatives of the proposal (adopted semantics: the reference
    allocates new_dz / new_dth as real arrays, :279-280, which silently drops their imaginary parts)."""
    addition = (0.0 + z_re * st.bump_sigma) + (0.0 + z_im * st.bump_sigma) * 1j            # :274
    new = st.psi + wave_array(addition, wavevector, st.Z, st.T_len)                         # :276-277
    new_energy = _energy_of(st, a, new)                                                     # :279-290
    accept, u_used = _decide_and_commit(st, old_energy, new_energy, new, uniform)           # :291-299
    update_bump_sigma(st, accept)                                                           # :300
    return accept, new_energy, u_used


def step_row_rotate(st, a, n_rot, phase, num_rows, row_index, uniform):
    """Lattice.step_row_rotate :303-382, adopted semantics: rows row_index .. row_index+num_rows-1 (mod Z) are multiplied by
    exp(i (2 pi n n_rot / T + phase)), n = theta index (:307,:319); the decision compares the energies of the affected rows
    (:372-374), which is the difference of the whole-lattice energies because nothing else changes.  (As committed the
    reference copies dth / dz of the FIRST row of the block into every row, :323-324, and its sublattice energy omits the
    metric's row offset for wrapped blocks.)  Returns (accept, old_energy, new_energy, u_used)."""
    angle_per_cell = math.pi * 2 / st.T_len                                                 # :307
    new = st.psi.copy()
    n = np.arange(st.T_len)
    for row in range(row_index, row_index + num_rows):
        new[row % st.Z] = st.psi[row % st.Z] * np.exp(1j * (angle_per_cell * n * n_rot + phase))   # :319
    old_energy = _energy_of(st, a, st.psi)
    new_energy = _energy_of(st, a, new)
    accept, u_used = _decide_and_commit(st, old_energy, new_energy, new, uniform)           # :374-381
    return accept, old_energy, new_energy, u_used


def winding_numbers(psi):
    """Winding number of the phase around each z-row: sum of the wrapped phase differences / 2 pi."""
    d = np.angle(psi * np.conj(np.roll(psi, 1, axis=1)))
    return np.rint(d.sum(axis=1) / (2 * math.pi)).astype(int)
