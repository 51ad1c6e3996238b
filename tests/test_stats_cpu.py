"""Equilibrium cut-off on block means (cylinder_b200/stats.py): plain tensor arithmetic, runs on the CPU."""
import numpy as np
import torch

from cylinder_b200 import stats


def test_mser_finds_the_end_of_a_transient_and_leaves_stationary_series_alone():
    rng = np.random.RandomState(0)
    n = 40
    t = np.arange(n)
    stationary = 1.0 + .05 * rng.randn(n)
    transient = 1.0 + 3.0 * np.exp(-t / 3.0) + .05 * rng.randn(n)
    x = torch.tensor(np.stack([stationary, transient], axis=1))
    cut = stats.mser_cut(x)
    assert cut.shape == (2,) and cut[0] <= 5 and 6 <= cut[1] <= 20
    mean, err, cut2 = stats.equilibrated(x)
    assert torch.equal(cut, cut2)
    assert abs(mean[0] - 1.0) < 4 * err[0] and abs(mean[1] - 1.0) < 4 * err[1] + .02
    assert abs(x[:, 1].mean() - 1.0) > .15                     # the untruncated mean is biased by the transient
    # cut of one series applied to another ("include profiles in cutoff and averaging", run.py:342)
    m2, e2, _ = stats.equilibrated(x[:, :1].expand(-1, 2), by=x)
    assert m2.shape == (2,)


def test_truncated_mean_matches_numpy():
    rng = np.random.RandomState(1)
    x = rng.randn(12, 3, 2)
    cut = torch.tensor([[0, 3], [5, 1], [6, 6]])
    mean, err = stats.truncated_mean(torch.tensor(x), cut)
    for i in range(3):
        for j in range(2):
            seg = x[int(cut[i, j]):, i, j]
            assert abs(mean[i, j].item() - seg.mean()) < 1e-14
            assert abs(err[i, j].item() - seg.std(ddof=1) / np.sqrt(len(seg))) < 1e-14
