"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only:
/root/reference is absent on the GPU box).  Output: small .npz files committed beside this script.

    python tests/golden/make_golden.py

What is recorded:
  geometry.npz          metric factors of system_cylinder.py:25-41 on a grid of (k, r, a, z)
  surface_energy.npz    a sub-sample of the reference's own tables surfenergy_H0.csv / curvenergy_H0.csv
                        plus Cylinder.calc_surface_energy at assorted (k, r, gamma, kappa, H0, a)
  lattice_<variant>_<case>.npz   (variants: simple, rescaled0, rescaled1, decoupled)
                        seeded Lattice construction (random.seed), then a recorded random tape of
                        proposals driven through the reference's own step_lattice(); per proposal
                        (iz, ith, z_re, z_im, u|NaN, dE, accept, sigma_after), final psi and caches,
                        and surface_field_energy(a', a_ref) at several amplitudes.
  known_answers.npz     psi==1 field energy -2*pi^2 (tests/test_On_lattice.py:75-88) on both lattices
  fourier.npz           Cylinder2D / Cylinder1D tensors and energies (system_cylinder2D.py:103-198)
  collective.npz        sequences of Lattice.step_lattice_all moves (On_lattice_simple.py:247-261)
  farm/                 varfile.txt written by parallel_preprocessing.py; tables written by parallel_postprocessing.py
"""
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness as rh  # noqa: E402


def gen_geometry():
    cyl = rh.load("system_cylinder")
    rows = []
    rng = random.Random(11)
    for k in (0.05, 0.5, 0.7, 1.0, 1.395):
        for r in (0.5, 1.0, 2.0):
            c = cyl.Cylinder(wavenumber=k, radius=r, kappa=0)
            for a in (0.0, 0.12, -0.3, 0.5, -0.79, 0.9, -0.99):
                for _ in range(6):
                    z = rng.uniform(-0.2, 2 * math.pi / k)
                    rows.append((k, r, a, z, c.radius_rescaled(a), c.sqrt_g_theta(a, z), c.g_theta(a, z),
                                 c.sqrt_g_z(a, z), c.g_z(a, z), c.A_theta(a, z)))
    arr = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), table=arr,
                        columns=np.array(["k", "r", "a", "z", "radius_rescaled", "sqrt_g_theta", "g_theta",
                                          "sqrt_g_z", "g_z", "A_theta"]))
    print("geometry", arr.shape)


def gen_surface_energy():
    import pandas as pd
    cyl = rh.load("system_cylinder")
    surf = pd.read_csv(os.path.join(rh.REFERENCE_ROOT, "surfenergy_H0.csv"), index_col=0)
    curv = pd.read_csv(os.path.join(rh.REFERENCE_ROOT, "curvenergy_H0.csv"), index_col=0)
    ai = np.arange(0, surf.shape[0], 9)              # 23 amplitudes incl. -0.99 and 0.99
    ki = np.arange(0, surf.shape[1], 7)              # 20 wavenumbers incl. 0.005 and 1.395... (139 -> add)
    ki = np.unique(np.append(ki, surf.shape[1] - 1))
    ai = np.unique(np.append(ai, surf.shape[0] - 1))
    a_vals = surf.index.values[ai].astype(float)
    k_vals = surf.columns.values[ki].astype(float)
    # direct calls
    rows = []
    for (k, r, g, kap, H0) in [(1, 1, 1, 0, 0), (.5, 1, 1, 1, 0), (1.2, 1, 1, .3, .8), (.05, 1, 1, 1, 2),
                               (1.395, 1, 1, 1, 0), (.7, 2, .5, .25, 1.5), (1, .5, 2, 0, 0)]:
        c = cyl.Cylinder(wavenumber=k, radius=r, kappa=kap, gamma=g, intrinsic_curvature=H0)
        for a in (0, .3, -.7, .79, -.99, .95, 1e-3):
            rows.append((k, r, g, kap, H0, a, c.evaluate_A_integral_0(a), c.calc_bending_energy(a),
                         c.calc_surface_energy(a)))
    np.savez_compressed(os.path.join(HERE, "surface_energy.npz"),
                        a_vals=a_vals, k_vals=k_vals,
                        surfenergy_H0=surf.values[np.ix_(ai, ki)], curvenergy_H0=curv.values[np.ix_(ai, ki)],
                        direct=np.array(rows),
                        direct_columns=np.array(["k", "r", "gamma", "kappa", "H0", "a", "A_integral_0",
                                                 "bending_energy", "surface_energy"]))
    print("surface_energy table", len(ai), "x", len(ki), "direct", len(rows))


def record_lattice_case(modname, tag, seed, kwargs, amp, nprop, amps_eval):
    mod = rh.load(modname)
    random.seed(seed)
    lat = rh.quiet(mod.Lattice, **kwargs)
    psi0 = lat.lattice.copy()
    sigma0 = lat.sampling_width

    def field_energy(a2, ar):
        if modname == "On_lattice_decoupled":          # surface_field_energy(amplitude) only, On_lattice_decoupled.py:179
            return rh.quiet(lat.surface_field_energy, a2)
        return rh.quiet(lat.surface_field_energy, a2, ar, lat.lattice, lat.psi_squared, lat.dz, lat.dth)

    e0 = [(a2, ar, field_energy(a2, ar)) for (a2, ar) in amps_eval]
    dec = lat.me.decisions
    iz = np.zeros(nprop, np.int32)
    ith = np.zeros(nprop, np.int32)
    zre = np.zeros(nprop)
    zim = np.zeros(nprop)
    u = np.full(nprop, np.nan)
    dE = np.zeros(nprop)
    acc = np.zeros(nprop, np.uint8)
    sig = np.zeros(nprop)
    for i in range(nprop):
        with rh.Tape() as tape:
            lat.step_lattice(amp)
        assert len(tape.randrange) == 2 and len(tape.normals) == 2 and len(tape.uniforms) <= 1
        iz[i], ith[i] = tape.randrange
        zre[i], zim[i] = tape.normals
        old, new, accept, uu = dec[-1]
        assert old == 0
        if uu is not None:
            assert tape.uniforms == [uu]
            u[i] = uu
        dE[i], acc[i], sig[i] = new, accept, lat.sampling_width
    assert len(dec) == nprop
    e1 = [(a2, ar, field_energy(a2, ar)) for (a2, ar) in amps_eval]
    surf = [(a2, lat.surface.calc_surface_energy(a2)) for (a2, _) in amps_eval]
    prof = np.array([sum(col) for col in lat.lattice]), np.array([sum(abs(col)) for col in lat.lattice])
    np.savez_compressed(
        os.path.join(HERE, "lattice_%s.npz" % tag),
        variant=np.array(modname), seed=seed, amp=amp, sigma0=sigma0,
        params=np.array([kwargs[x] for x in ("wavenumber", "radius", "gamma", "kappa", "intrinsic_curvature",
                                               "alpha", "u", "C", "n", "temperature")], dtype=float),
        params_names=np.array(["wavenumber", "radius", "gamma", "kappa", "intrinsic_curvature", "alpha", "u",
                               "C", "n", "temperature"]),
        dims=np.array(kwargs["dims"]), shape=np.array(psi0.shape),
        z_pixel_len=lat.z_pixel_len, th_pixel_len=lat.th_pixel_len,
        psi0=psi0, iz=iz, ith=ith, zre=zre, zim=zim, u=u, dE=dE, accept=acc, sigma=sig,
        psi_final=lat.lattice, psi2_final=lat.psi_squared, dz_final=lat.dz, dth_final=lat.dth,
        step_counter=lat.step_counter, acceptance_counter=lat.acceptance_counter,
        field_energy_initial=np.array(e0), field_energy_final=np.array(e1), surface_energy=np.array(surf),
        profile_sum=prof[0], profile_abs_sum=prof[1])
    print(tag, psi0.shape, "accepted", int(acc.sum()), "of", nprop, "uniforms drawn", int(np.isfinite(u).sum()))


def gen_lattice():
    base = dict(amplitude=0, wavenumber=1, radius=1, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1,
                C=1, n=6, temperature=.01, temperature_final=.01, dims=(50, 50))
    amps_eval = [(0.0, 0.0), (0.3, 0.3), (-0.6, -0.6), (0.35, 0.3), (0.25, 0.3), (-0.79, -0.6)]
    for modname, short in (("On_lattice_simple", "simple"), ("On_lattice_rescaled_0thorder", "rescaled0")):
        # C1: 50x50, a=0.3
        record_lattice_case(modname, short + "_c1_a03", 12345, base, 0.3, 3000, amps_eval)
        # a = 0 (flat): local dE equals the global energy difference here
        record_lattice_case(modname, short + "_c1_a0", 1234, base, 0.0, 1500, amps_eval)
        # odd-sized lattice 71 x 25 (k=.7, dims=(50,25)), negative amplitude, warmer, different couplings
        odd = dict(base, wavenumber=.7, dims=(50, 25), temperature=.1, alpha=-.5, C=.75, n=2, u=1.5, radius=1.25)
        record_lattice_case(modname, short + "_odd_am06", 777, odd, -0.6, 2000, amps_eval)
        # T = 0 branch of the accept rule
        cold = dict(base, temperature=0, dims=(20, 10), wavenumber=1)
        record_lattice_case(modname, short + "_T0", 5, cold, 0.79, 600, amps_eval)


def gen_lattice_more_variants():
    """rescaled_1storder and decoupled: the same four situations as the two live variants, plus for the
    1st-order background a case whose log argument goes non-positive (alpha*area < -C q_min^2: On_lattice_rescaled_1storder.py:97-102)
    and one where the whole O(u) correction is skipped (:103-105)."""
    base = dict(amplitude=0, wavenumber=1, radius=1, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1,
                C=1, n=6, temperature=.01, temperature_final=.01, dims=(50, 50))
    amps_eval = [(0.0, 0.0), (0.3, 0.3), (-0.6, -0.6), (0.35, 0.3), (0.25, 0.3), (-0.79, -0.6)]
    for modname, short in (("On_lattice_rescaled_1storder", "rescaled1"), ("On_lattice_decoupled", "decoupled")):
        record_lattice_case(modname, short + "_c1_a03", 12345, base, 0.3, 3000, amps_eval)
        record_lattice_case(modname, short + "_c1_a0", 1234, base, 0.0, 1500, amps_eval)
        odd = dict(base, wavenumber=.7, dims=(50, 25), temperature=.1, alpha=-.5, C=.75, n=2, u=1.5, radius=1.25)
        record_lattice_case(modname, short + "_odd_am06", 777, odd, -0.6, 2000, amps_eval)
        cold = dict(base, temperature=0, dims=(20, 10), wavenumber=1)
        record_lattice_case(modname, short + "_T0", 5, cold, 0.79, 600, amps_eval)
    # coarse lattice, strongly negative alpha: large cells make alpha*area + C q^2 <= 0 for the lower (and the upper) cut-off
    coarse = dict(base, dims=(8, 6), alpha=-3, C=.5, temperature=.2)
    record_lattice_case("On_lattice_rescaled_1storder", "rescaled1_lowcut", 21, coarse, 0.2, 300, amps_eval)
    coarser = dict(base, dims=(6, 4), alpha=-8, C=.25, temperature=.2)
    record_lattice_case("On_lattice_rescaled_1storder", "rescaled1_nocorr", 22, coarser, -0.3, 300, amps_eval)
    # alpha > 0: both logarithms regular (:96), no ordered-reference term (:110-113)
    warm = dict(base, dims=(12, 10), alpha=.5, temperature=.3)
    record_lattice_case("On_lattice_rescaled_1storder", "rescaled1_posalpha", 23, warm, 0.4, 300, amps_eval)


def gen_known_answers():
    mod = rh.load("On_lattice_simple")
    out = {}
    for name, dims in (("fine", (500, 250)), ("coarse", (50, 25))):
        random.seed(1)
        lat = rh.quiet(mod.Lattice, radius=1, wavenumber=1, amplitude=0, gamma=1, kappa=0, intrinsic_curvature=0,
                       alpha=-1, u=1, C=1, n=1, temperature=0.01, temperature_final=0.01, dims=dims)
        Z, T = lat.z_len, lat.th_len
        ones = np.ones((Z, T), dtype=complex)
        zeros = np.zeros((Z, T), dtype=complex)
        out[name] = lat.surface_field_energy(0, 0, ones, np.ones((Z, T)), zeros, zeros)
        out[name + "_a04"] = lat.surface_field_energy(.4, .4, ones, np.ones((Z, T)), zeros, zeros)
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **out)
    print("known answers", out)


def gen_fourier():
    c2 = rh.load("system_cylinder2D")
    c1 = rh.load("system_cylinder1D")
    out = {}
    rng = np.random.RandomState(7)
    cases = [("k05_n6_21", dict(wavenumber=.5, radius=1, alpha=-1, C=1, u=1, n=6, kappa=0, gamma=1), (2, 1)),
             ("k1_n1_11", dict(wavenumber=1., radius=1, alpha=-1, C=1, u=1, n=1, kappa=1, gamma=1), (1, 1)),
             ("k08_n2_30", dict(wavenumber=.8, radius=.5, alpha=.5, C=2, u=.5, n=2, kappa=.5, gamma=2,
                                intrinsic_curvature=.3), (3, 0)),
             ("k12_n3_02", dict(wavenumber=1.2, radius=1, alpha=-2, C=.5, u=2, n=3, kappa=0, gamma=1), (0, 2))]
    for tag, kw, nfc in cases:
        sys2 = c2.Cylinder2D(num_field_coeffs=nfc, **kw)
        nz, nth = nfc
        coeffs = (rng.normal(0, .3, (2 * nth + 1, 2 * nz + 1)) + 1j * rng.normal(0, .3, (2 * nth + 1, 2 * nz + 1)))
        amps = np.array([0.0, 0.3, -0.6, 0.9, 0.05])
        e_change = []
        e_stored = []
        Aint = []
        Bmat = []
        e_surf = []
        for a in amps:
            e_change.append(sys2.calc_field_energy_amplitude_change(a, coeffs))
            sys2.save_temporary_matrices()
            e_stored.append(sys2.calc_field_energy(coeffs))
            Aint.append(sys2.tmp_A_integrals.copy())
            Bmat.append(sys2.tmp_B_matrix.copy())
            e_surf.append(complex(sys2.calc_surface_energy(a)))
        out[tag + "_params"] = np.array([kw.get(x, 0) for x in ("wavenumber", "radius", "alpha", "C", "u", "n",
                                                                "kappa", "gamma", "intrinsic_curvature")], dtype=float)
        out[tag + "_nfc"] = np.array(nfc)
        out[tag + "_coeffs"] = coeffs
        out[tag + "_amps"] = amps
        out[tag + "_E_change"] = np.array(e_change)
        out[tag + "_E_stored"] = np.array(e_stored)
        out[tag + "_Aint"] = np.array(Aint)
        out[tag + "_B"] = np.array(Bmat)
        out[tag + "_E_surf"] = np.array(e_surf)
        print("fourier", tag, "E", e_change[:2], "surf", e_surf[1])
    # 1D model on the (nz, 0) slice, as in tests/test_1Dvs2D.py
    kw = dict(wavenumber=1., radius=1, alpha=-1, C=1, u=1, n=1, kappa=1, gamma=1)
    s1 = c1.Cylinder1D(num_field_coeffs=1, **kw)
    co = rng.normal(0, .3, 3) + 1j * rng.normal(0, .3, 3)
    e1 = [s1.calc_field_energy_amplitude_change(a, co) for a in (0.0, .22, -.23, .4, -.5, .6, -.8, -.35)]
    out["c1d_coeffs"] = co
    out["c1d_amps"] = np.array([0.0, .22, -.23, .4, -.5, .6, -.8, -.35])
    out["c1d_E"] = np.array(e1)
    np.savez_compressed(os.path.join(HERE, "fourier.npz"), **out)


def gen_collective():
    """Lattice.step_lattice_all (On_lattice_simple.py:247-261) driven on the unmodified reference: per move the two unit
    normals, the uniform (NaN if none was drawn), the real parts of old / new energy, the decision and the adapted width.
    (The reference's energies pick up an imaginary rounding residue of order 1e-19 from psi*conj(psi); only real parts are kept.)"""
    import warnings
    mod = rh.load("On_lattice_simple")
    out = {}
    for tag, seed, kw, amp, n in (("warm", 31, dict(temperature=.5, dims=(20, 10), alpha=-1, n=2), 0.25, 60),
                                  ("cold", 32, dict(temperature=.01, dims=(16, 12), alpha=-.5, n=6), -0.4, 60)):
        base = dict(amplitude=0, wavenumber=1, radius=1, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1, C=1, n=6,
                    temperature=.01, temperature_final=.01, dims=(50, 50))
        base.update(kw)
        random.seed(seed)
        lat = rh.quiet(mod.Lattice, **base)
        lat.sampling_width = .02
        lat.energy = lat.surface_field_energy(amp, amp, lat.lattice, lat.psi_squared, lat.dz, lat.dth)
        rec = dict(psi0=lat.lattice.copy(), e0=float(np.real(lat.energy)), zre=[], zim=[], u=[], e_old=[], e_new=[], accept=[], sigma=[])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(n):
                with rh.Tape() as tape:
                    rh.quiet(lat.step_lattice_all, amp)
                old, new, accept, uu = lat.me.decisions[-1]
                rec["zre"].append(tape.normals[0]); rec["zim"].append(tape.normals[1])
                rec["u"].append(np.nan if uu is None else uu)
                rec["e_old"].append(float(np.real(old))); rec["e_new"].append(float(np.real(new)))
                rec["accept"].append(bool(accept)); rec["sigma"].append(lat.sampling_width)
        rec["psi_final"] = lat.lattice.copy()
        P = [base[x] for x in ("wavenumber", "radius", "gamma", "kappa", "intrinsic_curvature", "alpha", "u", "C", "n", "temperature")]
        for k, v in rec.items():
            out[tag + "_" + k] = np.array(v)
        out[tag + "_params"] = np.array(P, dtype=float)
        out[tag + "_amp"] = amp
        out[tag + "_seed"] = seed
        out[tag + "_dims"] = np.array(base["dims"])
        print("collective", tag, "accepted", int(np.sum(rec["accept"])), "of", n, "uniforms", int(np.isfinite(rec["u"]).sum()))
    np.savez_compressed(os.path.join(HERE, "collective.npz"), **out)


def gen_farm():
    """File formats of the job farm, produced by the reference's own scripts (run as subprocesses in a scratch directory):
    varfile.txt of parallel_preprocessing.py and the per-column tables of parallel_postprocessing.py for a small set of
    synthetic *_mean.csv files (committed as inputs)."""
    import shutil
    import subprocess
    import tempfile
    import pandas as pd
    dst = os.path.join(HERE, "farm")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([sys.executable, os.path.join(rh.REFERENCE_ROOT, "parallel_preprocessing.py"), "-n", "golden scan",
                        "--var1name", "alpha", "--var1range", "-2", "1.1", ".75", "--var2name", "C", "--var2range", ".3", "1.0", ".3"],
                       cwd=tmp, check=True)
        shutil.copy(os.path.join(tmp, "varfile.txt"), os.path.join(dst, "varfile_ref.txt"))
        shutil.copy(os.path.join(tmp, "notes.txt"), os.path.join(dst, "notes_ref.txt"))
    with tempfile.TemporaryDirectory() as tmp:
        rng = np.random.RandomState(3)
        for v1 in ("-2.0", "-1.25", "0.25"):
            for v2 in ("0.3", "0.6"):
                d = {"abs_amplitude": [rng.rand()], "amplitude_squared": [rng.rand()], "field_energy": [rng.randn()]}
                name = "alpha_%s_C_%s_mean.csv" % (v1, v2)
                pd.DataFrame.from_dict(d).to_csv(os.path.join(tmp, name))
                shutil.copy(os.path.join(tmp, name), os.path.join(dst, name))
        subprocess.run([sys.executable, os.path.join(rh.REFERENCE_ROOT, "parallel_postprocessing.py")], cwd=tmp, check=True)
        for col in ("abs_amplitude", "amplitude_squared", "field_energy"):
            shutil.copy(os.path.join(tmp, col + ".csv"), os.path.join(dst, col + "_ref.csv"))
    print("farm fixtures:", sorted(os.listdir(dst)))


if __name__ == "__main__":
    what = sys.argv[1:] or ["geometry", "surface", "lattice", "lattice2", "known", "fourier", "farm", "collective"]
    if "geometry" in what:
        gen_geometry()
    if "surface" in what:
        gen_surface_energy()
    if "lattice" in what:
        gen_lattice()
    if "lattice2" in what:
        gen_lattice_more_variants()
    if "known" in what:
        gen_known_answers()
    if "fourier" in what:
        gen_fourier()
    if "farm" in what:
        gen_farm()
    if "collective" in what:
        gen_collective()
