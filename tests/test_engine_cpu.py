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
