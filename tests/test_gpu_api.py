"""GPU tests of the reference-facing Python API (the drop-in boundary): Cylinder, Lattice (both variants),
Cylinder2D, MetropolisEngine wiring, run.single_run -- against golden vectors recorded from the reference."""
import math
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BASE = dict(amplitude=0, wavenumber=1, radius=1, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1, C=1, n=6,
            temperature=.01, temperature_final=.01, dims=(50, 50))


def _case(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "lattice_%s.npz" % tag))
    P = dict(zip(g["params_names"].tolist(), g["params"].tolist()))
    kw = dict(amplitude=0, wavenumber=P["wavenumber"], radius=P["radius"], gamma=P["gamma"], kappa=P["kappa"],
              intrinsic_curvature=P["intrinsic_curvature"], alpha=P["alpha"], u=P["u"], C=P["C"], n=P["n"],
              temperature=P["temperature"], temperature_final=P["temperature"], dims=tuple(int(x) for x in g["dims"]))
    return g, kw


@pytest.mark.parametrize("tag,cls", [("simple_c1_a03", "Lattice"), ("rescaled0_odd_am06", "LatticeRescaled0thOrder"),
                                     ("simple_T0", "Lattice"), ("rescaled1_odd_am06", "LatticeRescaled1stOrder"),
                                     ("decoupled_c1_a03", "LatticeDecoupled")])
def test_lattice_sequential_path_reproduces_reference_stream(cuda, golden_dir, tag, cls):
    """random.seed(s); Lattice(...); step_lattice(a) x N consumes Python's generator exactly like the reference: same
    initial field, same accept/reject sequence, same adaptive width, same final field."""
    import cylinder_b200
    g, kw = _case(golden_dir, tag)
    random.seed(int(g["seed"]))
    lat = getattr(cylinder_b200, cls)(**kw)
    assert (lat.z_len, lat.th_len) == tuple(g["shape"])
    assert lat.z_pixel_len == float(g["z_pixel_len"]) and lat.th_pixel_len == float(g["th_pixel_len"])
    np.testing.assert_allclose(lat.lattice, g["psi0"], rtol=0, atol=1e-17)
    for a2, ar, e in g["field_energy_initial"]:
        assert lat.surface_field_energy(a2, ar, lat.lattice, lat.psi_squared, lat.dz, lat.dth) == pytest.approx(e, rel=1e-10)
    amp, n = float(g["amp"]), 400
    acc = []
    for i in range(n):
        before = lat.acceptance_counter
        lat.step_lattice(amp)
        acc.append(lat.acceptance_counter - before)
        assert lat.sampling_width == pytest.approx(float(g["sigma"][i]), rel=1e-12)
    assert acc == g["accept"][:n].tolist()
    assert lat.step_counter == n
    # caches are derived from psi
    assert np.allclose(lat.dz, (lat.lattice - np.roll(lat.lattice, 1, 0)) / lat.z_pixel_len)
    assert np.allclose(lat.psi_squared, np.abs(lat.lattice) ** 2)


def test_known_answer_through_lattice_api(cuda):
    """tests/test_On_lattice.py:75-88 of the reference, with the current signature."""
    import cylinder_b200
    random.seed(1)
    lat = cylinder_b200.Lattice(radius=1, wavenumber=1, amplitude=0, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1,
                                C=1, n=1, temperature=0.01, temperature_final=0.01, dims=(50, 25))
    ones = np.ones((lat.z_len, lat.th_len), dtype=complex)
    zeros = np.zeros_like(ones)
    e = lat.surface_field_energy(0, 0, ones, np.ones(ones.shape), zeros, zeros)
    assert e == pytest.approx(-.5 * 4 * math.pi ** 2, rel=1e-12)
    assert lat.z_pixel_len * lat.th_pixel_len * lat.z_len * lat.th_len == pytest.approx(4 * math.pi ** 2)   # pixel x count


def test_cylinder_api(cuda, golden_dir):
    import cylinder_b200
    t = np.load(os.path.join(golden_dir, "geometry.npz"))["table"]
    for k, r, a, z, rr, sgt, gt, sgz, gz, ath in t[::37]:
        c = cylinder_b200.Cylinder(wavenumber=k, radius=r, kappa=0)
        assert c.radius_rescaled(a) == pytest.approx(rr, rel=1e-15)
        assert c.sqrt_g_theta(a, z) == pytest.approx(sgt, rel=2e-15, abs=1e-16)
        assert c.g_theta(a, z) == pytest.approx(gt, rel=4e-15, abs=1e-16)
        assert c.sqrt_g_z(a, z) == pytest.approx(sgz, rel=2e-15)
        assert c.g_z(a, z) == pytest.approx(gz, rel=4e-15)
        assert c.A_theta(a, z) == pytest.approx(ath, rel=4e-15, abs=1e-16)
    # reference tests/test_system.py:22-68
    c = cylinder_b200.Cylinder(wavenumber=1.0, radius=1.0, kappa=1.0, gamma=1.0)
    assert c.sqrt_g_theta(0, .3) == 1 and c.sqrt_g_z(0, .3) == 1 and c.sqrt_g_z(.3, .2) > 1 and c.radius_rescaled(.4) <= 1
    g = np.load(os.path.join(golden_dir, "surface_energy.npz"))
    for k, r, gam, kap, H0, a, A0, bend, E in g["direct"][::5]:
        c = cylinder_b200.Cylinder(wavenumber=k, radius=r, kappa=kap, gamma=gam, intrinsic_curvature=H0)
        assert c.calc_surface_energy(a) == pytest.approx(E, rel=1e-10)
        assert c.evaluate_A_integral_0(a) == pytest.approx(A0, rel=1e-12)
    with pytest.raises(AssertionError):
        cylinder_b200.Cylinder(wavenumber=-1, radius=1, kappa=0)


def test_cylinder2d_api(cuda, golden_dir):
    import cylinder_b200
    g = np.load(os.path.join(golden_dir, "fourier.npz"))
    names = ("wavenumber", "radius", "alpha", "C", "u", "n", "kappa", "gamma", "intrinsic_curvature")
    tag = "k05_n6_21"
    P = dict(zip(names, g[tag + "_params"].tolist()))
    se = cylinder_b200.Cylinder2D(num_field_coeffs=tuple(int(x) for x in g[tag + "_nfc"]), **P)
    c = g[tag + "_coeffs"]
    for ia, a in enumerate(g[tag + "_amps"]):
        assert se.calc_field_energy_amplitude_change(float(a), c) == pytest.approx(float(g[tag + "_E_change"][ia]), rel=1e-10)
        se.save_temporary_matrices()
        assert se.calc_field_energy(c) == pytest.approx(float(g[tag + "_E_stored"][ia]), rel=1e-10)
        assert se.calc_surface_energy(float(a)) == pytest.approx(g[tag + "_E_surf"][ia].real, rel=1e-10)
    Eb = se.calc_field_energy_batch(g[tag + "_amps"], np.broadcast_to(c, (len(g[tag + "_amps"]),) + c.shape)).cpu().numpy()
    np.testing.assert_allclose(Eb, g[tag + "_E_change"], rtol=1e-10)


def test_lattice_run_and_engine_wiring(cuda, tmp_path):
    import cylinder_b200
    random.seed(7)
    lat = cylinder_b200.LatticeRescaled0thOrder(**dict(BASE, dims=(20, 16), temperature=.05, fieldsteps_per_ampstep=10))
    # host-driven engine move evaluates its energies on the device
    e0 = dict(lat.me.energy)
    assert e0["field"] == pytest.approx(lat.surface_field_energy(0, 0), rel=1e-12)
    assert e0["surface"] == pytest.approx(lat.surface.calc_surface_energy(0), rel=1e-12)
    lat.me.step_real_group()
    lat.run(30, 50)
    assert len(lat.field_profile_history) == 30 and lat.field_profile_history[0].shape == (lat.z_len,)
    assert abs(lat.amplitude) < .8 and lat.amplitude == lat.me.real_params[0]
    assert 0.3 < lat.acceptance_counter / lat.step_counter < 0.7         # Robbins-Monro drives towards 50 %
    lat.me.save_time_series()
    assert len(lat.me.df) == 30 and {"amplitude", "field_energy", "surface_energy"} <= set(lat.me.df.columns)
    prof, prof_abs = lat.get_profiles_df()
    assert prof.shape == (30, lat.z_len)
    lat.plot_save(str(tmp_path), "t")
    assert os.path.exists(tmp_path / "t_snapshot.csv")
    assert lat.me.energy["field"] == pytest.approx(lat.surface_field_energy(lat.amplitude, lat.amplitude), rel=1e-9)


@pytest.mark.parametrize("cls", ["Lattice", "LatticeRescaled1stOrder", "LatticeDecoupled"])
def test_lattice_run_fixed_amplitude(cuda, cls):
    """run_fixed_amplitude (:519-530): a whole-lattice shift proposal and the site sweeps per step on a frozen shape; the
    engine's field energy follows the lattice, the amplitude never moves, the width adapts."""
    import cylinder_b200
    random.seed(11)
    lat = getattr(cylinder_b200, cls)(**dict(BASE, amplitude=.25, dims=(16, 12), temperature=.05, fieldsteps_per_ampstep=4))
    w0 = lat.sampling_width
    lat.run_fixed_amplitude(12, 50)
    assert lat.amplitude == .25 and len(lat.field_abs_profile_history) == 12
    assert lat.me.energy["field"] == pytest.approx(lat.surface_field_energy(.25, .25), rel=1e-9)
    assert lat.sampling_width != w0 and lat.step_counter >= 12 * (16 * 12 + 1)
    lat.me.save_time_series()
    assert len(lat.me.df) == 12


def test_single_run_lattice_and_fourier(cuda, tmp_path, monkeypatch):
    from cylinder_b200 import run
    monkeypatch.chdir(tmp_path)
    random.seed(11)
    np.random.seed(11)
    out = tmp_path / "out"
    out.mkdir()
    names, obs, means, cov = run.single_run(temp=.05, temp_final=.05, n_steps=20, field_type="lattice", method="sequential",
                                            num_field_coeffs=(0, 0), fieldsteps_per_ampstep=10, measure_every=1, alpha=-1, C=1, n=6, u=1,
                                            gamma=1, kappa=0, radius=1, intrinsic_curvature=0, n_substeps=None, dims=(20, 16),
                                            wavenumber=1, amplitude=0, outdir=str(out), title="lat")
    assert names == ["amplitude"] and obs == ["abs_amplitude", "amplitude_squared"]
    assert {"abs_amplitude", "amplitude_squared", "field_energy"} <= set(means)
    for f in ("lat.csv", "lat_profile.csv", "lat_profile_abs.csv", "lat_snapshot.csv", "lat_avglattice.csv"):
        assert (out / f).exists(), f
    assert (tmp_path / "last_sigma.pickle").exists()
    names, obs, means, cov = run.single_run(temp=.1, temp_final=.1, n_steps=15, field_type="fourier", method="sequential",
                                            num_field_coeffs=(1, 1), fieldsteps_per_ampstep=3, measure_every=2, alpha=-1, C=1, n=6, u=1,
                                            gamma=1, kappa=0, radius=1, intrinsic_curvature=0, n_substeps=None, dims=(20, 16),
                                            wavenumber=.5, amplitude=0, outdir=str(out), title="four")
    assert len(names) == 1 + 9 and names[1] == "coeff_-1_-1" and cov.shape == (9, 9)
    assert (out / "four.csv").exists() and (tmp_path / "last_cov.pickle").exists()


@pytest.mark.parametrize("mode", ["cta", "warp"])
def test_batched_fourier_chains_match_oracle(cuda, mode):
    """cyl_fourier_am_steps_f64 (amplitude + coefficient moves, Robbins-Monro widths, recursive mean and covariance with
    re-factorisation after measurement 50) vs the numpy restatement on the same Philox draws, in both layouts: one CTA per chain
    (few chains: the energy quadrature shared by eight warps) and one warp per chain (many chains)."""
    import torch
    from cylinder_b200.fourier_chains import FourierChains
    from oracle import fourier_am_oracle as fa
    B, nfc = 3, (1, 1)
    P = dict(wavenumber=[.5, .8, 1.0], radius=1.0, gamma=1.0, kappa=[0.0, .3, 0.0], intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0,
             n=[6.0, 2.0, 1.0], temperature=.1, amp_bound=.99)
    rng = np.random.RandomState(3)
    c0 = rng.normal(0, .15, (B, 9)) + 1j * rng.normal(0, .15, (B, 9))
    fc = FourierChains(P, num_field_coeffs=nfc, amplitude=.05, field_coeffs=c0, seed=99, chain0=10)
    fc.run(40, measure_every=2, fieldsteps_per_ampstep=2, layout=mode)
    fc.run(25, measure_every=2, fieldsteps_per_ampstep=2, layout=mode)          # resumes; crosses measurement 50 (covariance adaptation on)
    for b in range(B):
        Pb = {k: (v[b] if isinstance(v, list) else v) for k, v in P.items()}
        ch = fa.Chain(Pb, 1, 1, .05, c0[b], bound=.99)
        ch.run(40, 2, 2, seed=99, chain=10 + b)
        ch.run(25, 2, 2, seed=99, chain=10 + b, step0=40)
        s = fc.state[b].cpu().numpy()
        assert (s[5], s[6]) == (ch.acc_r, ch.acc_c) and ch.acc_c > 5
        assert fc.amp[b].item() == pytest.approx(ch.a, rel=1e-9, abs=1e-12)
        np.testing.assert_allclose(fc.coeffs[b].cpu().numpy(), ch.c, rtol=1e-8, atol=1e-11)
        assert s[0] == pytest.approx(ch.sig_r, rel=1e-10) and s[1] == pytest.approx(ch.sig_c, rel=1e-10)
        assert s[7] == pytest.approx(ch.e_field, rel=1e-9) and s[8] == pytest.approx(ch.e_surf, rel=1e-9)
        np.testing.assert_allclose(fc.mean[b].cpu().numpy(), ch.mean, rtol=1e-9)
        np.testing.assert_allclose(fc.covariance[b].cpu().numpy(), ch.cov, rtol=1e-7, atol=1e-12)
    summ = fc.summary()
    assert torch.all(summ["complex_acceptance"] > 0) and torch.all(fc.amp.abs() < .99)


@pytest.mark.parametrize("method", ["simultaneous", "fixed-amplitude", "no-field"])
def test_batched_fourier_chains_other_drivers_match_oracle(cuda, method):
    """The other drivers of run.py:295-317 in the batched kernel vs the numpy restatement, same Philox draws; the run crosses
    measurement 50, so the joint proposal is re-factorised from the adapted covariance."""
    import torch
    from cylinder_b200.fourier_chains import FourierChains
    from oracle import fourier_am_oracle as fa
    B, nfc = 2, (1, 1)
    P = dict(wavenumber=[.6, 1.0], radius=1.0, gamma=1.0, kappa=[0.2, 0.0], intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0,
             n=[2.0, 1.0], temperature=.1, amp_bound=.99)
    rng = np.random.RandomState(8)
    c0 = rng.normal(0, .15, (B, 9)) + 1j * rng.normal(0, .15, (B, 9))
    fc = FourierChains(P, num_field_coeffs=nfc, amplitude=.1, field_coeffs=c0, seed=5, chain0=3)
    fc.run(35, measure_every=3, fieldsteps_per_ampstep=2, method=method)
    fc.run(30, measure_every=3, fieldsteps_per_ampstep=2, method=method)
    for b in range(B):
        Pb = {k: (v[b] if isinstance(v, list) else v) for k, v in P.items()}
        ch = fa.Chain(Pb, 1, 1, .1, c0[b], bound=.99)
        ch.run(35, 3, 2, seed=5, chain=3 + b, method=method)
        ch.run(30, 3, 2, seed=5, chain=3 + b, step0=35, method=method)
        s = fc.state[b].cpu().numpy()
        assert (s[5], s[6]) == (ch.acc_r, ch.acc_c) and ch.acc_r + ch.acc_c > 5
        assert fc.amp[b].item() == pytest.approx(ch.a, rel=1e-9, abs=1e-12)
        np.testing.assert_allclose(fc.coeffs[b].cpu().numpy(), ch.c, rtol=1e-8, atol=1e-11)
        assert s[0] == pytest.approx(ch.sig_r, rel=1e-10) and s[1] == pytest.approx(ch.sig_c, rel=1e-10)
        assert s[7] == pytest.approx(ch.e_field, rel=1e-9, abs=1e-14) and s[8] == pytest.approx(ch.e_surf, rel=1e-9)
        np.testing.assert_allclose(fc.mean[b].cpu().numpy(), ch.mean, rtol=1e-9)
        np.testing.assert_allclose(fc.covariance[b].cpu().numpy(), ch.cov, rtol=1e-7, atol=1e-12)
    if method == "fixed-amplitude":
        assert torch.all(fc.amp == .1)
    if method == "no-field":
        np.testing.assert_array_equal(fc.coeffs.cpu().numpy(), c0)


def test_farm_run_writes_the_reference_files(cuda, tmp_path):
    """infile.txt + varfile.txt -> one batched launch: one <var1>_<x1>_<var2>_<x2>_mean.csv (+ profile CSVs) per line, named
    with the strings of the varfile, joinable by postprocess(); results do not depend on how the lines are batched."""
    import pandas as pd
    from cylinder_b200 import farm
    (tmp_path / "infile.txt").write_text("--n_steps 12\n--field_type lattice\n--method sequential\n--alpha -1\n--C 1\n--u 1\n--n 2\n"
                                         "--kappa 0\n--gamma 1\n--temp .2\n--intrinsic_curvature 0\n--amplitude 0\n--radius 1\n"
                                         "--wavenumber 1\n--measure_every 1\n--fieldsteps_per_ampstep 10\n--dims 12 10\n")
    farm.preprocess("t", "alpha", [-1.5, 0, .5], "wavenumber", [.8, 1.3, .25], outdir=str(tmp_path))
    names, lines = farm.read_varfile(str(tmp_path / "varfile.txt"))
    assert len(lines) == 6
    out = farm.run(str(tmp_path / "infile.txt"), str(tmp_path / "varfile.txt"), outdir=str(tmp_path / "out"), seed=3, write_series=True)
    ts = pd.read_csv(tmp_path / "out" / (farm.title_of(names, lines[2]) + ".csv"), index_col=0)
    assert len(ts) == 12 and {"amplitude", "field_energy", "surface_energy", "energy", "psi_squared"} <= set(ts.columns)
    assert out["abs_amplitude"].shape == (6,) and np.all(out["n_measurements"] == 6) if "n_measurements" in out else True
    for strings in lines:
        t = farm.title_of(names, strings)
        df = pd.read_csv(tmp_path / "out" / (t + "_mean.csv"), index_col=0)
        assert {"abs_amplitude", "amplitude_squared", "field_energy", "surface_energy", "energy"} <= set(df.columns)
        Z = int(round(12 / float(strings[1])))
        assert len(pd.read_csv(tmp_path / "out" / (t + "_profile_abs.csv"), index_col=0)) == Z
    cols = farm.postprocess(str(tmp_path / "out"))
    tab = pd.read_csv(tmp_path / "out" / "abs_amplitude.csv", index_col=0)
    assert "abs_amplitude" in cols and tab.shape == (2, 3)
    # a sub-range of the lines run alone gives the same numbers (replica streams depend on the global line index only)
    (tmp_path / "var2.txt").write_text("alpha wavenumber\n" + "\n".join(" ".join(s) for s in lines[:1]) + "\n")
    out1 = farm.run(str(tmp_path / "infile.txt"), str(tmp_path / "var2.txt"), outdir=str(tmp_path / "out1"), seed=3, write_profiles=False)
    assert out1["abs_amplitude"][0] == out["abs_amplitude"][0] and out1["field_energy"][0] == out["field_energy"][0]


@pytest.mark.parametrize("tag", ["warm", "cold"])
def test_step_lattice_all_reproduces_the_reference_stream(cuda, golden_dir, tag):
    """random.seed(s); Lattice(...); step_lattice_all(a) x N: same draws, same energies, same decisions, same adapted width and
    same final field as the sequence recorded from the reference (On_lattice_simple.py:247-261)."""
    import cylinder_b200
    g = np.load(os.path.join(golden_dir, "collective.npz"))
    P = g[tag + "_params"].tolist()
    random.seed(int(g[tag + "_seed"]))
    lat = cylinder_b200.Lattice(amplitude=0, wavenumber=P[0], radius=P[1], gamma=P[2], kappa=P[3], intrinsic_curvature=P[4], alpha=P[5],
                                u=P[6], C=P[7], n=P[8], temperature=P[9], temperature_final=P[9], dims=tuple(int(x) for x in g[tag + "_dims"]))
    np.testing.assert_allclose(lat.lattice, g[tag + "_psi0"], rtol=0, atol=1e-17)
    a = float(g[tag + "_amp"])
    lat.sampling_width = .02
    lat.energy = lat.surface_field_energy(a, a, lat.lattice, lat.psi_squared, lat.dz, lat.dth)
    assert lat.energy == pytest.approx(float(g[tag + "_e0"]), rel=1e-10)
    for i in range(len(g[tag + "_accept"])):
        assert lat.energy == pytest.approx(float(g[tag + "_e_old"][i]), rel=1e-10)
        acc = lat.step_lattice_all(a)
        assert acc == bool(g[tag + "_accept"][i])
        assert lat.sampling_width == pytest.approx(float(g[tag + "_sigma"][i]), rel=1e-12)
    np.testing.assert_allclose(lat.lattice, g[tag + "_psi_final"], rtol=0, atol=1e-14)


def test_collective_moves_match_oracle_and_batch(cuda):
    """cyl_move_row_moments / cyl_move_apply for the three proposal kinds, a ragged batch, fp64: energy of the proposal equals the
    oracle's energy of the explicitly transformed field (1e-10), the committed field equals the numpy transform, the accept mask
    leaves the other replicas untouched; the Lattice methods draw from Python's generator in the reference's order."""
    import torch
    import cylinder_b200
    from cylinder_b200 import ops
    from oracle import lattice_oracle as lo
    Zs, T = [9, 12, 7], 10
    ks = [12 / z for z in Zs]                                                  # dims[0] = 12 -> z_len = round(12 / k)
    P = dict(radius=1.1, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-.8, u=1.2, C=.9, n=3, temperature=.2)
    par = ops.make_params(cuda, wavenumber=ks, **P)
    rng = np.random.RandomState(0)
    psi_h = np.zeros((3, 12, T), complex)
    for b, z in enumerate(Zs):
        psi_h[b, :z] = rng.normal(0, .4, (z, T)) + 1j * rng.normal(0, .4, (z, T))
    Zt = torch.tensor(Zs, dtype=torch.int32, device=cuda)
    a = torch.tensor([.3, -.2, .5], device=cuda)
    kinds = [ops.MOVE_SHIFT, ops.MOVE_WAVE, ops.MOVE_TWIST]
    moves = ops.make_moves(cuda, kinds, B=3, c=[.05 - .02j, .1 + .03j, 0], qz=[0, 2 * math.pi, 0], qth=[0, -4 * math.pi, 0],
                           r0=[0, 0, 5], nr=[0, 0, 3], n_rot=[0, 0, -1], phase=[0, 0, 1.1])
    for variant in (lo.SIMPLE, lo.RESCALED_1ST):
        psi = torch.tensor(psi_h, dtype=torch.complex128, device=cuda)
        S = ops.move_row_moments(psi, Zt, par, moves)
        E = ops.field_energy(S, Zt, T, par, a, a, variant).cpu().numpy()
        np.testing.assert_array_equal(psi.cpu().numpy(), psi_h)                # evaluation does not touch the field
        expect = []
        for b, z in enumerate(Zs):
            st = lo.LatticeState(psi_h[b, :z], ks[b], P["radius"], P["alpha"], P["u"], P["C"], P["n"], P["temperature"], variant=variant)
            if b == 0:
                new = st.psi + (.05 - .02j)
            elif b == 1:
                new = st.psi + lo.wave_array(.1 + .03j, (2 * math.pi, -4 * math.pi), z, T)
            else:
                new = st.psi.copy()
                for row in (5, 6, 0):
                    new[row] = st.psi[row] * np.exp(1j * (2 * math.pi / T * np.arange(T) * -1 + 1.1))
            expect.append(new)
            assert E[b] == pytest.approx(lo._energy_of(st, float(a[b]), new), rel=1e-10)
        ops.move_apply(psi, Zt, moves, accept=torch.tensor([1, 0, 1], dtype=torch.uint8, device=cuda))
        out = psi.cpu().numpy()
        np.testing.assert_allclose(out[0, :9], expect[0], atol=1e-15)
        np.testing.assert_array_equal(out[1], psi_h[1])
        np.testing.assert_allclose(out[2, :7], expect[2], atol=1e-15)
        assert np.all(out[2, 7:] == 0)
    # Lattice methods: row twist at T = 1e9 is always accepted and winds exactly the drawn block
    random.seed(4)
    lat = cylinder_b200.Lattice(amplitude=0, wavenumber=1, radius=1, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1, C=1, n=2,
                                temperature=1e9, temperature_final=1e9, dims=(10, 8))
    lat.lattice = np.abs(lat.lattice) + .5
    st_r = random.getstate()
    n_rot, phase, num_rows, row_index = random.choice([-1, 1]), random.uniform(0, 2 * math.pi), random.randint(1, 3), random.randint(0, 9)
    random.setstate(st_r)
    assert lat.step_row_rotate(.2, maxrows=3)
    w = lo.winding_numbers(lat.lattice)
    assert w.tolist() == [n_rot if (i - row_index) % 10 < num_rows else 0 for i in range(10)]
    lat.energy = lat.surface_field_energy(.2, .2)
    lat.step_lattice_wavevector(.2, (2 * math.pi, 0.0))
    lat.step_lattice_random_wavevector(.2)
    assert lat.bump_step_counter == 2
    assert lat.surface_field_energy(.2, .2) == pytest.approx(lat.energy, rel=1e-10)      # bookkeeping follows the accepted moves


def test_lattice_run_on_the_tiled_hbm_path(cuda):
    """A lattice too large for shared memory runs through the TMA-tiled workspace; the packed copy is refreshed lazily.  Checks
    the bookkeeping that crosses the two layouts: energies, profiles, running average, host edits between sweeps."""
    import cylinder_b200
    random.seed(3)
    lat = cylinder_b200.Lattice(**dict(BASE, dims=(150, 100), temperature=.05, fieldsteps_per_ampstep=50))
    assert not lat._fits_shared_memory()
    lat.run(4, 50, sweeps_per_step=2)
    assert lat._hbm is not None and len(lat.field_profile_history) == 4
    e = lat.me.energy["field"]
    psi = lat.lattice                                          # forces the unpack
    assert e == pytest.approx(lat.surface_field_energy(lat.amplitude, lat.amplitude, psi, None, None, None), rel=1e-9)
    np.testing.assert_allclose(lat.field_abs_profile_history[-1].shape, (150,))
    assert np.isfinite(lat.avg_lattice).all() and np.abs(lat.avg_lattice).max() > 0
    # a host edit is picked up by the next sweep (repacked), and the sweep's result is visible afterwards
    new = psi.copy()
    new[0, 0] = 5 + 5j
    lat.lattice = new
    lat.sweep(1)
    out = lat.lattice
    assert out.shape == (150, 100) and not np.array_equal(out, new) and abs(out[3, 3] - new[3, 3]) < 1.0
    # direct comparison of the two evaluation routes on the same field
    lat.sweep(1, measure=True)
    e_hbm = lat.surface_field_energy(.1, .1)                   # from the tiled layout
    e_packed = lat.surface_field_energy(.1, .1, lat.lattice.copy(), None, None, None)
    assert e_hbm == pytest.approx(e_packed, rel=1e-12)


def test_run_experiment_scans_a_grid(cuda, tmp_path):
    """run.run_experiment (reference run.py:133-179): notes.csv + one table per observable; the batched route runs the grid as
    one launch, the serial route is the reference's loop over single_run; both fill the same `<name>_mean` tables."""
    import pandas as pd
    from cylinder_b200 import run
    kw = dict(n_steps=8, dims=(12, 10), temp=.05, fieldsteps_per_ampstep=10, measure_every=1, field_type="lattice", n=2, notes="t")
    res, d = run.run_experiment(("wavenumber", "alpha"), [.8, 1.0, 1.2], [-1.0, 0.0], exp_dir=str(tmp_path / "b"), **kw)
    assert os.path.exists(os.path.join(d, "notes.csv"))
    tab = pd.read_csv(os.path.join(d, "abs_amplitude_mean.csv"), index_col=0)
    assert tab.shape == (3, 2) and np.isfinite(tab.to_numpy()).all()
    assert np.asarray(res["field_energy_mean"]).shape == (3, 2)
    random.seed(2)
    cwd = os.getcwd()
    os.chdir(tmp_path)                                        # single_run drops last_sigma.pickle in the working directory
    try:
        res2, d2 = run.run_experiment(("alpha", "C"), [-1.0], [.5, 1.0], batched=False, exp_dir=str(tmp_path / "s"), **dict(kw, n_steps=3))
    finally:
        os.chdir(cwd)
    assert np.asarray(res2["abs_amplitude_mean"]).shape == (1, 2) and os.path.exists(os.path.join(d2, "abs_amplitude_mean.csv"))
