"""GPU parity tests: the CUDA path (through the C-ABI) against the golden vectors recorded from the reference
and against the oracle on the same seeded inputs.  Bars: accept/reject sequences exact; fp64 energies within
1e-10 relative (BASELINE.json north_star); the fp32 kernels within the tolerances written at each test."""
import glob
import math
import os

import numpy as np
import pytest

from oracle import checkerboard_oracle as cb
from oracle import lattice_oracle as lo

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
LATTICE_CASES = sorted(os.path.basename(p)[len("lattice_"):-4] for p in glob.glob(os.path.join(HERE, "golden", "lattice_*.npz")))


def _load_case(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "lattice_%s.npz" % tag))
    P = dict(zip(g["params_names"].tolist(), g["params"].tolist()))
    variant = lo.VARIANTS[str(g["variant"])[len("On_lattice_"):]]
    return g, P, variant


def _params(ops, dev, P, **over):
    kw = dict(P)
    kw.update(over)
    return ops.make_params(dev, **kw)


def test_geometry_matches_reference(cuda, golden_dir):
    import torch
    from cylinder_b200 import ops
    t = np.load(os.path.join(golden_dir, "geometry.npz"))["table"]
    for k, r in sorted(set(map(tuple, t[:, :2]))):
        rows = t[(t[:, 0] == k) & (t[:, 1] == r)]
        out = ops.geometry_eval(k, r, torch.tensor(rows[:, 2], device=cuda), torch.tensor(rows[:, 3], device=cuda)).cpu().numpy()
        np.testing.assert_allclose(out, rows[:, 4:10], rtol=2e-15, atol=1e-16)


def test_surface_energy_matches_reference(cuda, golden_dir):
    """Cylinder.calc_surface_energy incl. the reference's own tables surfenergy_H0.csv / curvenergy_H0.csv."""
    import torch
    from cylinder_b200 import ops
    g = np.load(os.path.join(golden_dir, "surface_energy.npz"))
    d = g["direct"]
    par = ops.make_params(cuda, wavenumber=d[:, 0], radius=d[:, 1], gamma=d[:, 2], kappa=d[:, 3], intrinsic_curvature=d[:, 4],
                          alpha=-1, u=1, C=1, n=1, temperature=.01)
    E, terms = ops.surface_energy(par, torch.tensor(d[:, 5], device=cuda), return_terms=True)
    np.testing.assert_allclose(terms[:, 0].cpu().numpy(), d[:, 6], rtol=1e-12)
    np.testing.assert_allclose(terms[:, 1].cpu().numpy(), d[:, 7], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(E.cpu().numpy(), d[:, 8], rtol=1e-10)
    A, K = np.meshgrid(g["a_vals"], g["k_vals"], indexing="ij")
    par = ops.make_params(cuda, wavenumber=K.ravel(), radius=1.0, gamma=1.0, kappa=1.0, intrinsic_curvature=0.0, alpha=-1, u=1,
                          C=1, n=1, temperature=.01)
    _, terms = ops.surface_energy(par, torch.tensor(A.ravel(), device=cuda), return_terms=True)
    np.testing.assert_allclose(terms[:, 0].cpu().numpy().reshape(A.shape), g["surfenergy_H0"], rtol=1e-12)
    np.testing.assert_allclose(terms[:, 1].cpu().numpy().reshape(A.shape), g["curvenergy_H0"], rtol=1e-10)


def test_philox_known_answers(cuda):
    import torch
    from cylinder_b200 import ops
    ctr = torch.from_numpy(np.array([k[0] for k in cb.PHILOX_KAT], dtype=np.uint32).view(np.int32)).to(cuda)
    key = torch.from_numpy(np.array([k[1] for k in cb.PHILOX_KAT], dtype=np.uint32).view(np.int32)).to(cuda)
    out = ops.philox4x32_10(ctr.contiguous(), key.contiguous()).cpu().numpy().view(np.uint32)
    assert out.tolist() == [list(k[2]) for k in cb.PHILOX_KAT]
    ctr = torch.from_numpy(np.array([k[0] for k in cb.PHILOX2_KAT], dtype=np.uint32).view(np.int32)).to(cuda)
    key = torch.from_numpy(np.array([k[1] for k in cb.PHILOX2_KAT], dtype=np.uint32).view(np.int32)).to(cuda)
    out = ops.philox2x32_10(ctr.contiguous(), key.contiguous()).cpu().numpy().view(np.uint32)
    assert out.tolist() == [list(k[2]) for k in cb.PHILOX2_KAT]


@pytest.mark.parametrize("tag", LATTICE_CASES)
def test_replay_reproduces_reference_chain(cuda, golden_dir, tag):
    """A chain fed the reference's recorded random stream reproduces its accept/reject sequence exactly; dE within
    1e-10 relative; sigma trajectory and final lattice to rounding."""
    import torch
    from cylinder_b200 import ops
    g, P, variant = _load_case(golden_dir, tag)
    par = _params(ops, cuda, P)
    psi = torch.tensor(g["psi0"], dtype=torch.complex128, device=cuda).contiguous()
    rm = torch.tensor([lo.INIT_SIGMA[variant], 0.0], dtype=torch.float64, device=cuda)       # :101 / dec:99
    u = np.where(np.isfinite(g["u"]), g["u"], 2.0)          # never consulted where the reference drew none
    acc, dE, sig = ops.replay(psi, par, float(g["amp"]), variant, torch.tensor(g["iz"], device=cuda), torch.tensor(g["ith"], device=cuda),
                              torch.tensor(g["zre"], device=cuda), torch.tensor(g["zim"], device=cuda), torch.tensor(u, device=cuda), rm)
    assert np.array_equal(acc.cpu().numpy(), g["accept"])
    np.testing.assert_allclose(dE.cpu().numpy(), g["dE"], rtol=1e-10, atol=1e-18)
    np.testing.assert_allclose(sig.cpu().numpy(), g["sigma"], rtol=1e-12)
    np.testing.assert_allclose(psi.cpu().numpy(), g["psi_final"], rtol=0, atol=1e-14)
    assert rm[1].item() == int(g["step_counter"])
    # the drawn uniforms were consulted exactly where the reference drew them
    drew = np.isfinite(g["u"])
    assert np.all((g["dE"] > 0)[drew]) and (P["temperature"] == 0 or np.all(drew[g["dE"] > 0]))


@pytest.mark.parametrize("tag", LATTICE_CASES)
def test_field_energy_matches_reference(cuda, golden_dir, tag):
    """surface_field_energy(a', a_ref) of the reference's initial and final configurations, fp64, <= 1e-10 relative;
    profiles of Lattice.measure."""
    import torch
    from cylinder_b200 import ops
    g, P, variant = _load_case(golden_dir, tag)
    par = _params(ops, cuda, P)
    for key_psi, key_e in (("psi0", "field_energy_initial"), ("psi_final", "field_energy_final")):
        psi = torch.tensor(g[key_psi], dtype=torch.complex128, device=cuda)[None].contiguous()
        Z = torch.tensor([psi.shape[1]], dtype=torch.int32, device=cuda)
        S = ops.row_moments(psi, Z, par)
        for a2, ar, e in g[key_e]:
            E = ops.field_energy(S, Z, psi.shape[2], par, torch.tensor([a2], device=cuda), torch.tensor([ar], device=cuda), variant)
            assert E.item() == pytest.approx(e, rel=1e-10)
    np.testing.assert_allclose(S[0, :, 6].cpu().numpy() + 1j * S[0, :, 7].cpu().numpy(), g["profile_sum"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(S[0, :, 5].cpu().numpy(), g["profile_abs_sum"], rtol=1e-12)
    # fp32 storage: same energies within 2e-6 relative (fp32 rounding of psi, fp64 accumulation)
    psi32 = psi.to(torch.complex64)
    S32 = ops.row_moments(psi32, Z, par)
    a2, ar, e = g["field_energy_final"][1]
    E32 = ops.field_energy(S32, Z, psi.shape[2], par, torch.tensor([a2], device=cuda), torch.tensor([ar], device=cuda), variant)
    assert E32.item() == pytest.approx(e, rel=2e-6)


def test_known_answer_uniform_field(cuda):
    """tests/test_On_lattice.py:75-88 of the reference: psi == 1, alpha=-1, u=1, a=0  =>  E = -2 pi^2 on (500,250) and (50,25)."""
    import torch
    from cylinder_b200 import ops
    par = ops.make_params(cuda, wavenumber=1, radius=1, alpha=-1, u=1, C=1, n=1, temperature=.01)
    for dims in ((500, 250), (50, 25)):
        Zl, T = ops.lattice_dims(dims, 1.0)
        psi = torch.ones(1, Zl, T, dtype=torch.complex128, device=cuda)
        Z = torch.tensor([Zl], dtype=torch.int32, device=cuda)
        E = ops.field_energy(ops.row_moments(psi, Z, par), Z, T, par, torch.zeros(1, device=cuda), None, 0)
        assert E.item() == pytest.approx(-.5 * 4 * math.pi ** 2, rel=1e-12)


def _oracle_state(psi, P, variant, sigma):
    st = lo.LatticeState(psi, P["wavenumber"], P["radius"], P["alpha"], P["u"], P["C"], P["n"], P["temperature"], variant=variant)
    st.sigma = sigma
    return st


CHECKER_CASES = [  # (Z, T, variant, a, temperature)
    (12, 10, lo.SIMPLE, 0.3, .05), (12, 10, lo.RESCALED_0TH, -0.45, .05), (11, 10, lo.SIMPLE, 0.3, .05),
    (12, 9, lo.RESCALED_0TH, 0.2, .05), (7, 5, lo.SIMPLE, -0.6, .2), (10, 8, lo.SIMPLE, 0.5, 0.0),
    (12, 10, lo.RESCALED_1ST, 0.3, .05), (11, 9, lo.RESCALED_1ST, -0.5, .1), (12, 10, lo.DECOUPLED, 0.3, .05),
    (9, 10, lo.DECOUPLED, -0.45, 0.0),
]


@pytest.mark.parametrize("Z,T,variant,a,temp", CHECKER_CASES)
def test_smem_sweep_f64_matches_oracle(cuda, Z, T, variant, a, temp):
    """fp64 shared-memory checkerboard kernel vs the numpy restatement, same Philox stream, several sweeps incl. odd
    rings (3 colours): lattice to 1e-11, identical accept counts and Robbins-Monro sigma."""
    import torch
    from cylinder_b200 import ops, _abi
    P = dict(wavenumber=.9, radius=1.1, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-.8, u=1.2, C=.9, n=3, temperature=temp)
    rng = np.random.RandomState(Z * 100 + T)
    psi0 = rng.normal(0, .3, (Z, T)) + 1j * rng.normal(0, .3, (Z, T))
    st = _oracle_state(psi0.copy(), P, variant, .05)
    par = _params(ops, cuda, P)
    psi = torch.tensor(psi0, dtype=torch.complex128, device=cuda)[None].contiguous()
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    chain = ops.new_chain(1, cuda, amplitude=a, sigma_site=.05)
    seed, rep, nsw = 0xC0FFEE1234, 17, 4
    ops.sweep_smem(psi, Zt, par, chain, seed, 100, nsw, variant, _abi.SWEEP_ADAPT_SIGMA, replica0=rep)
    for s in range(nsw):
        cb.sweep(st, a, seed, rep, 100 + s)
    np.testing.assert_allclose(psi[0].cpu().numpy(), st.psi, rtol=0, atol=1e-11)
    ch = chain[0].cpu().numpy()
    assert ch[_abi.CH_SITE_ACCEPTED] == st.accepted and ch[_abi.CH_STEP_COUNTER] == st.step_counter
    assert ch[_abi.CH_SIGMA_SITE] == pytest.approx(st.sigma, rel=1e-12)


@pytest.mark.parametrize("variant,Z,T", [(lo.SIMPLE, 12, 10), (lo.RESCALED_0TH, 12, 10), (lo.RESCALED_1ST, 12, 10), (lo.DECOUPLED, 12, 10),
                                         (lo.SIMPLE, 11, 10), (lo.RESCALED_0TH, 13, 8), (lo.SIMPLE, 12, 9), (lo.RESCALED_1ST, 9, 7)])
def test_smem_sweep_with_amplitude_moves_matches_oracle(cuda, variant, Z, T):
    """Sweeps + global amplitude moves + measurement, fp64, vs the oracle: same amplitude trajectory end point,
    same accept counts, energies within 1e-10.  Sizes cover the column-mapped 2- and 3-colour site loops (T even) and the
    generic ones (T odd), and with them the two forms of the energy pass that follows the colour passes (column pairs / a warp per row)."""
    import torch
    from cylinder_b200 import ops, _abi
    P = dict(wavenumber=.9, radius=1.1, gamma=1.0, kappa=0.3, intrinsic_curvature=0.2, alpha=-.8, u=1.2, C=.9, n=3, temperature=.3)
    rng = np.random.RandomState(5)
    psi0 = rng.normal(0, .3, (Z, T)) + 1j * rng.normal(0, .3, (Z, T))
    st = _oracle_state(psi0.copy(), P, variant, .05)
    am = cb.AmpState(amplitude=.1, sigma_amp=.05, target=.3, bound=.8)
    par = _params(ops, cuda, P, amp_bound=.8, amp_target_acceptance=.3)
    psi = torch.tensor(psi0, dtype=torch.complex128, device=cuda)[None].contiguous()
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    chain = ops.new_chain(1, cuda, amplitude=.1, sigma_site=.05, sigma_amp=.05)
    obs, prof = ops.new_obs(1, Z, cuda)
    seed, rep, nsw = 77, 3, 12
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
    ops.sweep_smem(psi, Zt, par, chain, seed, 0, nsw, variant, flags, replica0=rep, obs=obs, prof=prof)
    sum_abs_a = sum_ef = sum_abs_psi = 0.0
    prof_o = np.zeros((Z, 3))
    for s in range(nsw):
        cb.sweep(st, am.a, seed, rep, s)
        sum_abs_a += abs(am.a)
        sum_ef += lo.surface_field_energy(st, am.a, am.a)
        sum_abs_psi += float(np.mean(np.abs(st.psi)))
        ps, pa = lo.measure_profiles(st.psi)
        prof_o += np.stack([ps.real, ps.imag, pa], 1)
        cb.amplitude_move(st, am, P, seed, rep, s)
    ch = chain[0].cpu().numpy()
    assert ch[_abi.CH_AMP_ACCEPTED] == am.accepted and ch[_abi.CH_AMP_PROPOSED] == am.proposed
    assert am.accepted >= 1, "test should exercise an accepted amplitude move"
    assert ch[_abi.CH_AMPLITUDE] == pytest.approx(am.a, rel=1e-12)
    assert ch[_abi.CH_SIGMA_AMP] == pytest.approx(am.sigma, rel=1e-12)
    np.testing.assert_allclose(psi[0].cpu().numpy(), st.psi, rtol=0, atol=1e-11)
    o = obs[0].cpu().numpy()
    assert o[_abi.OB_COUNT] == nsw
    assert o[_abi.OB_ABS_A] == pytest.approx(sum_abs_a, rel=1e-12)
    assert o[_abi.OB_E_FIELD] == pytest.approx(sum_ef, rel=1e-10)
    assert o[_abi.OB_ABSPSI] == pytest.approx(sum_abs_psi, rel=1e-11)        # assembled from the profile pass's row sums
    np.testing.assert_allclose(prof[0].cpu().numpy(), prof_o, rtol=1e-10, atol=1e-12)
    # without a profile buffer the site loop accumulates sum |psi| itself
    chain2 = ops.new_chain(1, cuda, amplitude=.1, sigma_site=.05, sigma_amp=.05)
    psi2 = torch.tensor(psi0, dtype=torch.complex128, device=cuda)[None].contiguous()
    obs2, _ = ops.new_obs(1, Z, cuda)
    ops.sweep_smem(psi2, Zt, par, chain2, seed, 0, nsw, variant, flags, replica0=rep, obs=obs2, prof=None)
    assert obs2[0, _abi.OB_ABSPSI].item() == pytest.approx(sum_abs_psi, rel=1e-11)


@pytest.mark.parametrize("Z,T,variant", [(64, 128, lo.SIMPLE), (40, 24, lo.RESCALED_0TH), (71, 25, lo.SIMPLE), (70, 131, lo.SIMPLE),
                                         (130, 260, lo.SIMPLE), (40, 24, lo.RESCALED_1ST), (33, 130, lo.DECOUPLED)])
def test_hbm_sweep_f64_matches_oracle(cuda, Z, T, variant):
    """TMA-tiled HBM kernel (fp64) vs the oracle: tiles with halos, partial tiles, multi-tile lattices and odd rings
    reproduce the oracle's checkerboard sweep on the whole lattice."""
    import torch
    from cylinder_b200 import ops, _abi
    a = 0.3
    P = dict(wavenumber=.9, radius=1.1, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-.8, u=1.2, C=.9, n=3, temperature=.05)
    rng = np.random.RandomState(Z + T)
    psi0 = rng.normal(0, .3, (Z, T)) + 1j * rng.normal(0, .3, (Z, T))
    st = _oracle_state(psi0.copy(), P, variant, .05)
    par = _params(ops, cuda, P)
    ws = ops.HbmWorkspace(Z, T, torch.complex128, cuda)
    ws.pack(torch.tensor(psi0, dtype=torch.complex128, device=cuda))
    np.testing.assert_array_equal(ws.unpack().cpu().numpy(), psi0)
    chain = ops.new_chain(1, cuda, amplitude=a, sigma_site=.05)
    seed, rep, nsw = 991, 2, 3
    ws.sweep(par, chain[0], seed, 5, nsw, variant, _abi.SWEEP_ADAPT_SIGMA, replica=rep)
    for s in range(nsw):
        cb.sweep(st, a, seed, rep, 5 + s)
    np.testing.assert_allclose(ws.unpack().cpu().numpy(), st.psi, rtol=0, atol=1e-11)
    ch = chain[0].cpu().numpy()
    assert ch[_abi.CH_SITE_ACCEPTED] == st.accepted
    assert ch[_abi.CH_SIGMA_SITE] == pytest.approx(st.sigma, rel=1e-12)
    # row moments of the plane-split layout == packed-layout kernel
    S1 = ws.row_moments(par).cpu().numpy()
    S2 = ops.row_moments(ws.unpack()[None].contiguous(), torch.tensor([Z], dtype=torch.int32, device=cuda), par)[0].cpu().numpy()
    np.testing.assert_allclose(S1, S2, rtol=1e-12, atol=1e-13)


def test_hbm_and_smem_paths_agree_f32(cuda):
    """The two fp32 kernels run the same arithmetic on the same Philox stream: identical lattices after a sweep."""
    import torch
    from cylinder_b200 import ops, _abi
    Z, T = 64, 48
    P = dict(wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6, temperature=.01)
    par = ops.make_params(cuda, **P)
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    psi = torch.zeros(1, Z, T, dtype=torch.complex64, device=cuda)
    ops.lattice_init(psi, Zt, .1, seed=5)
    ws = ops.HbmWorkspace(Z, T, torch.complex64, cuda)
    ws.pack(psi[0])
    c1, c2 = ops.new_chain(1, cuda, amplitude=.3), ops.new_chain(1, cuda, amplitude=.3)
    ops.sweep_smem(psi, Zt, par, c1, 42, 0, 3, 0, _abi.SWEEP_ADAPT_SIGMA)
    ws.sweep(par, c2[0], 42, 0, 3, 0, _abi.SWEEP_ADAPT_SIGMA)
    assert torch.equal(ws.unpack(), psi[0])
    assert torch.equal(c1, c2)


def test_f32_sweep_tracks_f64(cuda):
    """fp32 kernel vs fp64 kernel on the same stream: after one sweep from the same start, <= 0.5 % of the sites may
    differ in their accept decision (borderline exp comparisons); the others agree to 2e-6 absolute."""
    import torch
    from cylinder_b200 import ops, _abi
    Z, T = 60, 50
    P = dict(wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6, temperature=.01)
    par = ops.make_params(cuda, **P)
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    psi64 = torch.zeros(1, Z, T, dtype=torch.complex128, device=cuda)
    ops.lattice_init(psi64, Zt, .1, seed=9)
    psi32 = psi64.to(torch.complex64)
    start = psi32.to(torch.complex128)
    psi64 = start.clone()
    c32, c64 = ops.new_chain(1, cuda, amplitude=.3), ops.new_chain(1, cuda, amplitude=.3)
    ops.sweep_smem(psi32, Zt, par, c32, 1, 0, 1, 0, 0)
    ops.sweep_smem(psi64, Zt, par, c64, 1, 0, 1, 0, 0)
    diff = (psi32.to(torch.complex128) - psi64).abs()[0]
    mism = (diff > 2e-6).float().mean().item()
    assert mism < 5e-3, mism
    assert abs(c32[0, _abi.CH_SITE_ACCEPTED].item() - c64[0, _abi.CH_SITE_ACCEPTED].item()) <= 0.005 * Z * T


def test_lattice_init_matches_oracle(cuda):
    import torch
    from cylinder_b200 import ops
    Zs = [7, 12]
    psi = torch.zeros(2, 12, 10, dtype=torch.complex128, device=cuda)
    ops.lattice_init(psi, torch.tensor(Zs, dtype=torch.int32, device=cuda), .1, seed=123456789012, replica0=40)
    for b, Z in enumerate(Zs):
        ref = cb.random_initialize(12, 10, .1, 123456789012, 40 + b)
        np.testing.assert_allclose(psi[b, :Z].cpu().numpy(), ref[:Z], rtol=0, atol=1e-16)
        assert torch.all(psi[b, Z:] == 0)
    assert psi.abs().max().item() <= .1


def test_batched_replicas_are_independent_and_ragged(cuda):
    """B replicas with different Z (ragged, incl. odd) and parameters in one launch == each run alone."""
    import torch
    from cylinder_b200 import ops, _abi
    ks = np.array([.5, .55, .8, 1.0, 1.25])
    Zs = [ops.lattice_dims((50, 50), k)[0] for k in ks]
    assert any(z % 2 for z in Zs)
    B, Zmax, T = len(ks), max(Zs), 50
    par = ops.make_params(cuda, wavenumber=ks, radius=1.0, alpha=[-2, -1, -.5, 0.5, 1.0], u=1.0, C=[.5, 1, 2, 1, 1], n=6, temperature=.01)
    Zt = torch.tensor(Zs, dtype=torch.int32, device=cuda)
    psi = torch.zeros(B, Zmax, T, dtype=torch.complex64, device=cuda)
    ops.lattice_init(psi, Zt, .1, seed=3)
    start = psi.clone()
    chain = ops.new_chain(B, cuda, amplitude=.2)
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES
    ops.sweep_smem(psi, Zt, par, chain, 11, 0, 5, 0, flags)
    for b in range(B):
        p1 = start[b:b + 1].clone()
        c1 = ops.new_chain(1, cuda, amplitude=.2)
        ops.sweep_smem(p1, Zt[b:b + 1].contiguous(), par[b:b + 1].contiguous(), c1, 11, 0, 5, 0, flags, replica0=b)
        assert torch.equal(p1[0], psi[b])
        assert torch.equal(c1[0], chain[b])
    assert torch.all(psi[0, Zs[0]:] == 0) or Zs[0] == Zmax


def test_invalid_arguments_fail_loudly(cuda):
    import torch
    from cylinder_b200 import ops, _abi
    par = ops.make_params(cuda, wavenumber=1, radius=1, alpha=-1, u=1, C=1, n=1, temperature=.01)
    psi = torch.zeros(1, 8, 8, dtype=torch.complex64, device=cuda)
    Zt = torch.tensor([8], dtype=torch.int32, device=cuda)
    with pytest.raises(_abi.CylinderB200Error):
        ops.sweep_smem(psi, Zt, par, ops.new_chain(1, cuda), 0, 0, 1, variant=7)
    with pytest.raises(_abi.CylinderB200Error):
        ops.sweep_smem(psi.cpu(), Zt, par, ops.new_chain(1, cuda), 0, 0, 1)
    big = torch.zeros(1, 400, 400, dtype=torch.complex64, device=cuda)
    with pytest.raises(_abi.CylinderB200Error, match="shared memory"):
        ops.sweep_smem(big, torch.tensor([400], dtype=torch.int32, device=cuda), par, ops.new_chain(1, cuda), 0, 0, 1)
    with pytest.raises(_abi.CylinderB200Error):
        ops.HbmWorkspace(4, 4, torch.complex64, cuda)


FOURIER_CASES = ["k05_n6_21", "k1_n1_11", "k08_n2_30", "k12_n3_02"]
FOURIER_NAMES = ("wavenumber", "radius", "alpha", "C", "u", "n", "kappa", "gamma", "intrinsic_curvature")


@pytest.mark.parametrize("tag", FOURIER_CASES)
def test_fourier_energy_matches_reference(cuda, golden_dir, tag):
    """Cylinder2D.calc_field_energy_amplitude_change / calc_field_energy and Cylinder1D.calc_surface_energy recorded
    from the reference (tensor form, scipy quad) vs the single-integral CUDA kernel: <= 1e-10 relative."""
    import torch
    from cylinder_b200 import ops
    g = np.load(os.path.join(golden_dir, "fourier.npz"))
    P = dict(zip(FOURIER_NAMES, g[tag + "_params"].tolist()))
    nz, nth = (int(x) for x in g[tag + "_nfc"])
    amps = g[tag + "_amps"]
    B = len(amps)
    par = ops.make_params(cuda, temperature=.1, **P)
    par = par.repeat(B, 1).contiguous()
    c = torch.tensor(g[tag + "_coeffs"], dtype=torch.complex128, device=cuda)[None].repeat(B, 1, 1).contiguous()
    E = ops.fourier_energy(par, nz, nth, torch.tensor(amps, device=cuda), c).cpu().numpy()
    np.testing.assert_allclose(E, g[tag + "_E_change"], rtol=1e-10)
    np.testing.assert_allclose(E, g[tag + "_E_stored"], rtol=1e-10)
    Es = ops.fourier_surface_energy(par, torch.tensor(amps, device=cuda)).cpu().numpy()
    np.testing.assert_allclose(Es, g[tag + "_E_surf"].real, rtol=1e-10)


def test_hbm_path_with_amplitude_moves_tracks_smem_path(cuda):
    """fp64: the HBM path (energy and profile sums accumulated per tile, decision kernel) and the shared-memory path (one CTA
    per replica) see the same Philox streams, so the amplitude trajectory, the lattice and the accumulated observables
    agree to rounding."""
    import torch
    from cylinder_b200 import ops, _abi
    Z, T = 32, 24
    P = dict(wavenumber=1.0, radius=1.0, gamma=1.0, kappa=0.2, intrinsic_curvature=0.1, alpha=-1.0, u=1.0, C=1.0, n=2, temperature=.3)
    par = ops.make_params(cuda, **P)
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    psi = torch.zeros(1, Z, T, dtype=torch.complex128, device=cuda)
    ops.lattice_init(psi, Zt, .3, seed=21)
    ws = ops.HbmWorkspace(Z, T, torch.complex128, cuda)
    ws.pack(psi[0])
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
    c1, c2 = ops.new_chain(1, cuda, amplitude=.1, sigma_site=.2, sigma_amp=.05), ops.new_chain(1, cuda, amplitude=.1, sigma_site=.2, sigma_amp=.05)
    o1, p1 = ops.new_obs(1, Z, cuda)
    o2, p2 = ops.new_obs(1, Z, cuda)
    nsw = 25
    ops.sweep_smem(psi, Zt, par, c1, 5, 0, nsw, 0, flags, obs=o1, prof=p1)
    ws.sweep(par, c2[0], 5, 0, nsw, 0, flags, obs=o2[0], prof=p2[0])
    assert c1[0, _abi.CH_AMP_ACCEPTED].item() >= 2
    np.testing.assert_allclose(c2.cpu().numpy()[:, :3], c1.cpu().numpy()[:, :3], rtol=1e-10)
    assert c2[0, _abi.CH_AMP_ACCEPTED].item() == c1[0, _abi.CH_AMP_ACCEPTED].item()
    assert c2[0, _abi.CH_SITE_ACCEPTED].item() == c1[0, _abi.CH_SITE_ACCEPTED].item()
    np.testing.assert_allclose(ws.unpack().cpu().numpy(), psi[0].cpu().numpy(), rtol=0, atol=1e-10)
    np.testing.assert_allclose(o2.cpu().numpy()[0, :9], o1.cpu().numpy()[0, :9], rtol=1e-9)
    np.testing.assert_allclose(p2.cpu().numpy(), p1.cpu().numpy(), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("Z,T,variant,temp,tile_rows", [
    (70, 260, lo.SIMPLE, .3, 0), (71, 131, lo.RESCALED_1ST, .3, 0), (40, 136, lo.RESCALED_0TH, .5, 0),
    (33, 130, lo.DECOUPLED, .3, 0), (64, 256, lo.SIMPLE, 0.0, 0),
    # several tiles away from every lattice edge (their energy and profile sums come from the pass over the finished tile;
    # the tiles on the edges close their bonds inside the colour passes): 2 colours, 3 colours (T odd; Z and T odd), and the
    # tile heights whose last warp rows stay empty
    (100, 520, lo.SIMPLE, .3, 24), (96, 516, lo.RESCALED_0TH, .3, 32), (124, 392, lo.RESCALED_1ST, .3, 30),
    (75, 389, lo.SIMPLE, .3, 24), (90, 391, lo.RESCALED_0TH, .3, 28), (101, 517, lo.SIMPLE, .3, 32)])
def test_hbm_sweep_with_amplitude_moves_matches_oracle(cuda, Z, T, variant, temp, tile_rows):
    """Multi-tile lattices (interior, border and partial tiles, 2 and 3 colours): the energies accumulated tile by tile (inside
    the colour passes on the lattice's edge tiles, in one pass over the finished tile elsewhere), the profile sums and the
    amplitude decisions of the HBM path equal the oracle's whole-lattice evaluation (fp64, same Philox streams)."""
    import torch
    from cylinder_b200 import ops, _abi
    P = dict(wavenumber=.9, radius=1.1, gamma=1.0, kappa=0.3, intrinsic_curvature=0.2, alpha=-.8, u=1.2, C=.9, n=3, temperature=temp)
    rng = np.random.RandomState(Z + 7 * T)
    psi0 = rng.normal(0, .3, (Z, T)) + 1j * rng.normal(0, .3, (Z, T))
    st = _oracle_state(psi0.copy(), P, variant, .05)
    am = cb.AmpState(amplitude=.1, sigma_amp=.004, target=.3, bound=.8)
    par = _params(ops, cuda, P, amp_bound=.8, amp_target_acceptance=.3)
    ws = ops.HbmWorkspace(Z, T, torch.complex128, cuda)
    ws.pack(torch.tensor(psi0, dtype=torch.complex128, device=cuda))
    chain = ops.new_chain(1, cuda, amplitude=.1, sigma_site=.05, sigma_amp=.004)
    obs, prof = ops.new_obs(1, Z, cuda)
    seed, rep, nsw = 4711, 6, 8
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE | (_abi.sweep_tile_rows(tile_rows) if tile_rows else 0)
    ws.sweep(par, chain[0], seed, 3, nsw, variant, flags, replica=rep, obs=obs[0], prof=prof[0])
    sum_abs_a = sum_ef = sum_p2 = 0.0
    prof_o = np.zeros((Z, 3))
    for s in range(3, 3 + nsw):
        cb.sweep(st, am.a, seed, rep, s)
        sum_abs_a += abs(am.a)
        sum_ef += lo.surface_field_energy(st, am.a, am.a)
        sum_p2 += float(np.mean(st.psi2))
        ps, pa = lo.measure_profiles(st.psi)
        prof_o += np.stack([ps.real, ps.imag, pa], 1)
        cb.amplitude_move(st, am, P, seed, rep, s)
    ch = chain[0].cpu().numpy()
    assert ch[_abi.CH_AMP_ACCEPTED] == am.accepted and ch[_abi.CH_AMP_PROPOSED] == am.proposed == nsw
    assert ch[_abi.CH_SITE_ACCEPTED] == st.accepted
    assert ch[_abi.CH_AMPLITUDE] == pytest.approx(am.a, rel=1e-12)
    assert ch[_abi.CH_SIGMA_AMP] == pytest.approx(am.sigma, rel=1e-12)
    assert ch[_abi.CH_SIGMA_SITE] == pytest.approx(st.sigma, rel=1e-12)
    np.testing.assert_allclose(ws.unpack().cpu().numpy(), st.psi, rtol=0, atol=1e-11)
    o = obs[0].cpu().numpy()
    assert o[_abi.OB_COUNT] == nsw
    assert o[_abi.OB_ABS_A] == pytest.approx(sum_abs_a, rel=1e-12)
    assert o[_abi.OB_E_FIELD] == pytest.approx(sum_ef, rel=1e-10)
    assert o[_abi.OB_PSI2] == pytest.approx(sum_p2, rel=1e-11)
    np.testing.assert_allclose(prof[0].cpu().numpy(), prof_o, rtol=1e-10, atol=1e-11)
    if temp > 0:
        assert 1 <= am.accepted, "case should exercise an accepted amplitude move"


def test_step_host_equals_device_resident_run(cuda):
    """The pipelined host round trip (ReplicaBatch.step_host) gives the same chains as the device-resident run."""
    import torch
    from cylinder_b200.replicas import ReplicaBatch
    P = dict(wavenumber=np.linspace(.5, 1.25, 12), radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
    a = ReplicaBatch(P, dims=(50, 50), dtype=torch.complex64, device=cuda, seed=3)
    b = ReplicaBatch(P, dims=(50, 50), dtype=torch.complex64, device=cuda, seed=3)
    psi_h, chain_h = b.psi.cpu().pin_memory(), b.chain.cpu().pin_memory()
    psi_out = torch.empty_like(psi_h).pin_memory()
    rec = b.record()
    rec_out = torch.empty(rec.shape, dtype=rec.dtype).pin_memory()
    a.run(7)
    b.step_host(psi_h, chain_h, psi_out, rec_out, 7, nchunks=5)
    assert torch.equal(psi_out, a.psi.cpu())
    assert torch.equal(rec_out, a.record().cpu())


def test_time_series_matches_oracle(cuda):
    """cyl_sweep_smem_series: one record per sweep (amplitude, energies, <|psi|^2>, <|psi|>, width, accept flags), fp64, against
    the oracle's sweep-by-sweep values; the records sum to the accumulated observables."""
    import torch
    from cylinder_b200 import ops, _abi
    Z, T, variant = 11, 10, lo.RESCALED_0TH
    P = dict(wavenumber=.9, radius=1.1, gamma=1.0, kappa=0.3, intrinsic_curvature=0.2, alpha=-.8, u=1.2, C=.9, n=3, temperature=.3)
    rng = np.random.RandomState(9)
    psi0 = rng.normal(0, .3, (Z, T)) + 1j * rng.normal(0, .3, (Z, T))
    st = _oracle_state(psi0.copy(), P, variant, .05)
    am = cb.AmpState(amplitude=.1, sigma_amp=.05, target=.3, bound=.8)
    par = _params(ops, cuda, P, amp_bound=.8, amp_target_acceptance=.3)
    psi = torch.tensor(psi0, dtype=torch.complex128, device=cuda)[None].contiguous()
    Zt = torch.tensor([Z], dtype=torch.int32, device=cuda)
    chain = ops.new_chain(1, cuda, amplitude=.1, sigma_site=.05, sigma_amp=.05)
    obs, prof = ops.new_obs(1, Z, cuda)
    nsw, seed, rep = 10, 31, 2
    ts = torch.zeros(1, nsw, 8, dtype=torch.float64, device=cuda)
    flags = _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_AMP_MOVES | _abi.SWEEP_MEASURE
    ops.sweep_smem(psi, Zt, par, chain, seed, 0, nsw, variant, flags, replica0=rep, obs=obs, prof=prof, series=ts)
    rec = ts[0].cpu().numpy()
    for s in range(nsw):
        sigma_used, acc0 = st.sigma, st.accepted
        cb.sweep(st, am.a, seed, rep, s)
        assert rec[s, 0] == pytest.approx(am.a, rel=1e-12, abs=1e-15)
        assert rec[s, 1] == pytest.approx(lo.surface_field_energy(st, am.a, am.a), rel=1e-10)
        assert rec[s, 2] == pytest.approx(lo.surface_energy(P["wavenumber"], P["radius"], P["gamma"], P["kappa"], P["intrinsic_curvature"], am.a), rel=1e-10)
        assert rec[s, 3] == pytest.approx(float(np.mean(st.psi2)), rel=1e-11)
        assert rec[s, 4] == pytest.approx(float(np.mean(np.abs(st.psi))), rel=1e-11)
        assert rec[s, 5] == pytest.approx(sigma_used, rel=1e-12)
        assert rec[s, 7] == pytest.approx((st.accepted - acc0) / (Z * T), rel=1e-12)
        before = am.accepted
        cb.amplitude_move(st, am, P, seed, rep, s)
        assert rec[s, 6] == am.accepted - before
    o = obs[0].cpu().numpy()
    assert o[_abi.OB_E_FIELD] == pytest.approx(rec[:, 1].sum(), rel=1e-12) and o[_abi.OB_PSI2] == pytest.approx(rec[:, 3].sum(), rel=1e-12)
    with pytest.raises(Exception):
        ops.sweep_smem(psi, Zt, par, chain, seed, 0, nsw, variant, _abi.SWEEP_ADAPT_SIGMA, replica0=rep, series=ts)     # needs MEASURE


@pytest.mark.parametrize("Z,T", [(65, 257), (97, 136)])
def test_hbm_thin_last_tile_matches_oracle(cuda, Z, T):
    """Odd lengths whose last tile is a single row / column (65 = 2 x 32 + 1, 257 = 2 x 128 + 1; 97 = 3 x 32 + 1): the tile before
    it reaches across the periodic seam with its halo."""
    import torch
    from cylinder_b200 import ops, _abi
    a = -0.25
    P = dict(wavenumber=.9, radius=1.1, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-.8, u=1.2, C=.9, n=3, temperature=.05)
    rng = np.random.RandomState(Z + T)
    psi0 = rng.normal(0, .3, (Z, T)) + 1j * rng.normal(0, .3, (Z, T))
    st = _oracle_state(psi0.copy(), P, lo.SIMPLE, .05)
    par = _params(ops, cuda, P)
    for th in (32, 24):
        st = _oracle_state(psi0.copy(), P, lo.SIMPLE, .05)
        ws = ops.HbmWorkspace(Z, T, torch.complex128, cuda)
        ws.pack(torch.tensor(psi0, dtype=torch.complex128, device=cuda))
        chain = ops.new_chain(1, cuda, amplitude=a, sigma_site=.05)
        ws.sweep(par, chain[0], 123, 0, 2, lo.SIMPLE, _abi.SWEEP_ADAPT_SIGMA | _abi.sweep_tile_rows(th), replica=1)
        for s in range(2):
            cb.sweep(st, a, 123, 1, s)
        np.testing.assert_allclose(ws.unpack().cpu().numpy(), st.psi, rtol=0, atol=1e-11)
        assert chain[0, _abi.CH_SITE_ACCEPTED].item() == st.accepted


@pytest.mark.parametrize("T", [50, 600, 1100, 2050])
def test_row_moments_long_rows_match_torch(cuda, T):
    """Row moments (cyl_energy.cuh: S1..S5, sum |psi|, sum psi) on rows long enough that 2, 4 or 8 warps share a row, for the
    packed kernel, the plane-split (HBM workspace) kernel and the whole-lattice-move kernel with a plane wave: all three against
    the same sums taken in torch fp64.  Odd column counts (2050 / 2 planes, 1100) and a ragged batch included."""
    import math
    import torch
    from cylinder_b200 import ops
    Zs = [9, 6]
    P = dict(radius=1.2, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-.7, u=1.1, C=.8, n=2, temperature=.1)
    ks = [1.0, .8]
    par = ops.make_params(cuda, wavenumber=ks, **P)
    g = torch.Generator(device="cpu").manual_seed(T)
    psi = torch.zeros(2, 9, T, dtype=torch.complex128)
    for b, z in enumerate(Zs):
        psi[b, :z] = torch.complex(torch.randn(z, T, generator=g, dtype=torch.float64), torch.randn(z, T, generator=g, dtype=torch.float64)) * .4
    psi = psi.to(cuda)
    Zt = torch.tensor(Zs, dtype=torch.int32, device=cuda)

    def expect(field, b):
        z = Zs[b]
        f = field[b, :z]
        lz = (2 * math.pi / ks[b]) / z
        lth = (2 * math.pi * P["radius"]) / T
        dz = (f - torch.roll(f, 1, 0)) / lz
        dth = (f - torch.roll(f, 1, 1)) / lth
        p2 = f.real ** 2 + f.imag ** 2
        cols = [p2, p2 ** 2, dz.real ** 2 + dz.imag ** 2, dth.real ** 2 + dth.imag ** 2, (f.conj() * dth).imag, p2.sqrt(), f.real, f.imag]
        return torch.stack([c.sum(1) for c in cols], 1)

    S = ops.row_moments(psi, Zt, par)
    for b, z in enumerate(Zs):
        torch.testing.assert_close(S[b, :z], expect(psi, b), rtol=1e-11, atol=1e-11)
        assert (S[b, z:] == 0).all()
    # plane wave added on the fly: same moments as those of the explicitly transformed field
    qz, qth, c = 2 * math.pi, 6 * math.pi, .07 - .02j
    moves = ops.make_moves(cuda, [ops.MOVE_WAVE, ops.MOVE_WAVE], B=2, c=[c, c], qz=[qz, qz], qth=[qth, qth])
    Sm = ops.move_row_moments(psi, Zt, par, moves)
    moved = psi.clone()
    for b, z in enumerate(Zs):
        i = torch.arange(z, device=cuda, dtype=torch.float64)[:, None]
        j = torch.arange(T, device=cuda, dtype=torch.float64)[None, :]
        moved[b, :z] += c * torch.exp(1j * (i * qz / z + j * qth / T))
        torch.testing.assert_close(Sm[b, :z], expect(moved, b), rtol=1e-10, atol=1e-10)
    if T % 2 == 0 and T >= 8:
        ws = ops.HbmWorkspace(Zs[0], T, torch.complex128, cuda)
        ws.pack(psi[0].contiguous())
        torch.testing.assert_close(ws.row_moments(par[0:1]), expect(psi, 0), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("tag", FOURIER_CASES)
def test_fourier_tensor_tables_match_reference(cuda, golden_dir, tag):
    """Cylinder2D.evaluate_A_integrals / evaluate_B_integrals (system_cylinder2D.py:103-147; the attributes the reference's own
    tests/test_matrixA.py and tests/test_1Dvs2D.py read): tmp_A_integrals and tmp_B_matrix recorded from the reference (scipy quad)
    vs the device tables, and the reference's index layout of tmp_A_matrix / tmp_D_matrix; contracted by einsum they reproduce the
    recorded energies (system_cylinder2D.py:176-198), i.e. the tensor route and the single-integral kernel agree."""
    import math
    import cylinder_b200
    g = np.load(os.path.join(golden_dir, "fourier.npz"))
    P = dict(zip(FOURIER_NAMES, g[tag + "_params"].tolist()))
    nz, nth = (int(x) for x in g[tag + "_nfc"])
    sys2 = cylinder_b200.Cylinder2D(num_field_coeffs=(nz, nth), device=cuda, **P)
    assert len(sys2.quartic_selection_rule) == sum(1 for a in range(2 * nth + 1) for b in range(2 * nth + 1) for c in range(2 * nth + 1)
                                                   for d in range(2 * nth + 1) if a + b == c + d)
    c = g[tag + "_coeffs"]
    for ia, a in enumerate(g[tag + "_amps"]):
        sys2.evaluate_A_integrals(float(a))
        sys2.evaluate_B_integrals(float(a))
        scale = max(1.0, float(np.abs(g[tag + "_Aint"][ia]).max()))
        np.testing.assert_allclose(sys2.tmp_A_integrals, g[tag + "_Aint"][ia], rtol=0, atol=1e-10 * scale)
        np.testing.assert_allclose(sys2.tmp_B_matrix, g[tag + "_B"][ia], rtol=0, atol=1e-10 * max(1.0, float(np.abs(g[tag + "_B"][ia]).max())))
        A, D, B = sys2.tmp_A_matrix, sys2.tmp_D_matrix, sys2.tmp_B_matrix
        lz = 2 * nz + 1
        for i in range(lz):
            for j in range(lz):
                assert A[i, j] == sys2.tmp_A_integrals[(i - j) + 4 * nz]           # A[i, j] = A_int[d = i - j]
        assert D.shape == (lz, lz, lz, lz) and D[0, lz - 1, 0, lz - 1] == sys2.tmp_A_integrals[4 * nz]
        Ae = sum(np.einsum("ij, i, j -> ", A, c[b], c[b].conjugate()) for b in range(2 * nth + 1))
        De = sum(np.einsum("ijkl, i, j, k, l -> ", D, c[i1], c[i2], c[i3].conjugate(), c[i4].conjugate())
                 for (i1, i2, i3, i4) in sys2.quartic_selection_rule)
        Be = np.einsum("ijab, ai, bj -> ", B, c, c.conjugate())
        E = P["alpha"] * Ae.real + P["C"] * Be.real + 0.5 * P["u"] * De.real
        assert E == pytest.approx(float(g[tag + "_E_change"][ia]), rel=1e-9)
        assert sys2.calc_field_energy_amplitude_change(float(a), c) == pytest.approx(E, rel=1e-9)
        sys2.save_temporary_matrices()
        assert sys2.A_matrix is sys2.tmp_A_matrix and sys2.calc_field_energy(c) == pytest.approx(E, rel=1e-9)
        sys2.evaluate_A_integral_0(float(a))
        assert sys2.tmp_A_integrals_0 == pytest.approx(complex(sys2.tmp_A_integrals[4 * nz]))
    sys2.evaluate_A_integrals(0.0)                        # tests/test_matrixA.py: the diff = 0 element at a = 0 is 2 pi r / k
    assert sys2.tmp_A_integrals[4 * nz] == pytest.approx(2 * math.pi * P["radius"] / P["wavenumber"], rel=1e-12)


def test_cylinder1d_matches_reference(cuda, golden_dir):
    """Cylinder1D (system_cylinder1D.py) = the 2-D model without theta modes: energies recorded from the reference's 1-D class,
    its attribute names, and 1D == 2D on the (n, 0) slice as in the reference's tests/test_1Dvs2D.py."""
    import cylinder_b200
    from cylinder_b200.surfaces_and_fields.system_cylinder1D import Cylinder1D
    g = np.load(os.path.join(golden_dir, "fourier.npz"))
    kw = dict(wavenumber=1., radius=1, alpha=-1, C=1, u=1, n=1, kappa=1, gamma=1)
    s1 = Cylinder1D(num_field_coeffs=1, device=cuda, **kw)
    s2 = cylinder_b200.Cylinder2D(num_field_coeffs=(1, 0), device=cuda, **kw)
    co = g["c1d_coeffs"]
    assert s1.len_arrays == 3 and s1.tmp_A_integrals.shape == (9,)
    for a, e in zip(g["c1d_amps"], g["c1d_E"]):
        e1 = s1.calc_field_energy_amplitude_change(float(a), co)
        assert e1 == pytest.approx(float(e), rel=1e-10)
        assert s2.calc_field_energy_amplitude_change(float(a), co[None, :]) == pytest.approx(e1, rel=1e-13)
        s1.evaluate_B_integrals(float(a))
        assert s1.tmp_B_integrals.shape == (3, 3)
        s1.save_temporary_matrices()
        assert s1.calc_field_energy(co) == pytest.approx(e1, rel=1e-13) and s1.B_integrals.shape == (3, 3)
        s2.evaluate_B_integrals(float(a))
        np.testing.assert_array_equal(s1.tmp_B_integrals, s2.tmp_B_matrix[:, :, 0, 0])
