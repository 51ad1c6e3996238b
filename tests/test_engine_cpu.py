"""Host-side engine mirror (cylinder_b200/metropolisengine.py): accept rule known answers recovered from the
reference's stale test bytecode, Robbins-Monro constants, adaptive behaviour on an analytic
target, recursive mean / covariance.  CPU only (the engine is control logic; energies here are plain lambdas)."""
import math
import random

import numpy as np
import pytest

from cylinder_b200.metropolisengine import MetropolisEngine, _ppf, robbins_monro_ratio


def _engine(temp, energy=lambda real, cplx: 0.0, a0=0.0, **kw):
    return MetropolisEngine(energy_functions={"real": {"surface": energy}, "all": {"surface": energy}},
                            initial_real_params=[a0], temp=temp, **kw)


def test_metropolis_decision_known_answers():
    never = random.uniform

    def boom(*a):
        raise AssertionError("no uniform draw expected")
    random.uniform = boom
    try:
        assert _engine(1e-3).metropolis_decision(12.45, -1451) is True
        assert _engine(0).metropolis_decision(-145.4, -145.3) is False
        assert _engine(0).metropolis_decision(-145.4, -145.5) is True
        assert _engine(0).metropolis_decision(0, 0) is True
        assert _engine(1.5).metropolis_decision(14.3, 14.3) is True          # equal energies are accepted
    finally:
        random.uniform = never
    e = _engine(0.5)
    e.set_temperature(2.0)
    assert e.temp == 2.0


def test_robbins_monro_constants():
    import scipy.stats
    for p in (1e-4, .01, .15, .25, .5, .9, .999):
        assert _ppf(p) == pytest.approx(scipy.stats.norm.ppf(p), rel=1e-12, abs=1e-14)
    assert robbins_monro_ratio(.5, 1) == pytest.approx(4.0, rel=1e-15)      # On_lattice_simple.py:95-100
    # m = 1: the first term vanishes
    assert robbins_monro_ratio(.3, 1) == pytest.approx(1 / (.3 * .7), rel=1e-15)


def test_real_group_samples_gaussian_and_adapts():
    """E(a) = a^2 / (2 s^2) at T = 1 => a ~ N(0, s^2); acceptance adapts to the target."""
    random.seed(3)
    s = 0.2
    e = _engine(1.0, energy=lambda real, cplx: real[0] ** 2 / (2 * s * s), sampling_width=.01)
    xs, acc = [], 0
    for i in range(60000):
        acc += e.step_real_group()
        if i > 10000:
            xs.append(e.real_params[0])
    xs = np.array(xs)
    assert abs(xs.mean()) < 0.02
    assert xs.std() == pytest.approx(s, rel=0.08)
    assert e.real_group_accepted / e.real_group_step_counter == pytest.approx(.3, abs=.05)
    assert e.energy["surface"] == pytest.approx(e.real_params[0] ** 2 / (2 * s * s))


def test_reject_condition_counts_as_rejection():
    random.seed(1)
    e = _engine(1.0, sampling_width=.5)
    e.set_reject_condition(lambda real, cplx: abs(real[0]) >= .1)
    w0 = e.real_group_sampling_width
    for _ in range(200):
        e.step_real_group()
        assert abs(e.real_params[0]) < .1
    assert e.real_group_sampling_width < w0                                   # mostly rejected => width shrinks


def test_measure_running_mean_and_covariance():
    random.seed(5)
    np.random.seed(5)
    f = lambda real, cplx: 0.0     # noqa: E731
    e = MetropolisEngine(energy_functions={"real": {"s": f}, "complex": {"s": f}, "all": {"s": f}}, initial_real_params=[.1],
                         initial_complex_params=np.array([.3 + .1j, -.2j]), temp=1.0, complex_sample_method="magnitude-phase",
                         sampling_width=.05)
    states = []
    for i in range(300):
        e.step_real_group()
        e.step_complex_group()
        e.measure()
        states.append(np.concatenate([[abs(e.real_params[0])], np.abs(e.complex_params)]))
    states = np.array(states)
    k = e.measure_step_counter
    # running mean started from the initial state as sample 1
    init = np.array([.1, abs(.3 + .1j), .2])
    assert np.allclose(e.mean, (init + states.sum(0)) / k, rtol=1e-12)
    assert e.covariance_matrix.shape == (3, 3) and np.allclose(e.covariance_matrix, e.covariance_matrix.T)
    assert np.all(np.linalg.eigvalsh(e.covariance_matrix) > 0)
    assert e.covariance_matrix_complex.shape == (2, 2)
    e.save_time_series()
    assert len(e.df) == 300 and "energy" in e.df.columns
    means = e.save_equilibrium_stats()
    assert "s_energy" in means


def test_scan_driver_loop_and_plot_save(tmp_path):
    """run.loop / run.plot_save (reference run.py:21-58, :103-130): result naming and table layout, with a stand-in single_run."""
    import pandas as pd
    from cylinder_b200 import run
    calls = []

    def fake(v1, v2):
        calls.append((v1, v2))
        cov = np.array([[1.0 + v1, .5], [.5, 2.0 + v2]])
        return ["amplitude", "coeff_0_0", "coeff_0_1"], ["abs_amplitude"], {"abs_amplitude": v1 * 10 + v2}, cov
    res = run.loop(fake, [1, 2, 3], [0, 5])
    assert calls == [(1, 0), (1, 5), (2, 0), (2, 5), (3, 0), (3, 5)]
    assert res["abs_amplitude_mean"] == [[10, 15], [20, 25], [30, 35]]
    assert res["coeff_0_0_variance"][1] == [3.0, 3.0] and res["coeff_0_1_variance"][2] == [2.0, 7.0]
    assert res["coeff_0_0_coeff_0_1_covariance"] == [[.5, .5]] * 3 and len(res["time_per_experiment"]) == 3
    run.plot_save([1, 2, 3], [0, 5], res["abs_amplitude_mean"], "abs_amplitude_mean", str(tmp_path))
    df = pd.read_csv(tmp_path / "abs_amplitude_mean.csv", index_col=0)
    assert df.shape == (3, 2) and df.loc[2, "5"] == 25
    # results shorter than the ranges belong to the END of the ranges (run.py:117-118)
    run.plot_save([1, 2, 3], [0, 5], res["abs_amplitude_mean"][1:], "tail", str(tmp_path))
    assert list(pd.read_csv(tmp_path / "tail.csv", index_col=0).index) == [2, 3]
    with pytest.raises(KeyError):
        run.run_experiment(("alpha", "n"), [0], [1], exp_dir=str(tmp_path / "x"))
