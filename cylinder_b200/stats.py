"""Equilibrium statistics of block-averaged time series.

The reference delegates this to the un-vendored engine (`me.save_equilibrium_stats`, `me.equilibrated_means`, run.py:341-346:
"include profiles in cutoff and averaging"); its rule is not in the tree (parity unpinned), so a standard one is used:
the Marginal Standard Error Rule (White 1997) picks the truncation point d that minimises the estimated squared standard
error of the mean of what is left, var(x[d:]) / (n - d), restricted to the first half of the series; error bars come from
the spread of the retained block means (blocks are whole launches of many sweeps, i.e. much longer than the
autocorrelation time of the observables they average).

Everything is plain tensor arithmetic over [nblocks, ...] arrays (torch, any device), vectorised over replicas.
"""
import torch


def mser_cut(x, max_fraction=.5):
    """x: [n, ...] block means.  Returns the truncation index d in [0, max_fraction*n] per trailing element (int64 tensor)."""
    x = torch.as_tensor(x, dtype=torch.float64)
    n = x.shape[0]
    dmax = max(0, min(int(n * max_fraction), n - 2))
    # suffix sums give mean and variance of x[d:] for every d in one pass
    s1 = torch.flip(torch.cumsum(torch.flip(x, [0]), 0), [0])
    s2 = torch.flip(torch.cumsum(torch.flip(x * x, [0]), 0), [0])
    cnt = torch.arange(n, 0, -1, dtype=torch.float64, device=x.device).reshape([n] + [1] * (x.dim() - 1))
    var = (s2 / cnt - (s1 / cnt) ** 2).clamp(min=0.0)
    score = var / cnt
    return torch.argmin(score[: dmax + 1], dim=0)


def truncated_mean(x, cut):
    """Mean and standard error of x[cut:] along dim 0, `cut` per trailing element (as returned by mser_cut)."""
    x = torch.as_tensor(x, dtype=torch.float64)
    n = x.shape[0]
    idx = torch.arange(n, device=x.device).reshape([n] + [1] * (x.dim() - 1))
    keep = (idx >= cut.unsqueeze(0)).to(torch.float64)
    cnt = keep.sum(0)
    mean = (x * keep).sum(0) / cnt
    var = (((x - mean.unsqueeze(0)) ** 2) * keep).sum(0) / (cnt - 1).clamp(min=1.0)
    return mean, torch.sqrt(var / cnt)


def equilibrated(x, by=None, max_fraction=.5):
    """Truncate every series of x [n, ...] at the MSER point of `by` (default: its own) and return (mean, stderr, cut)."""
    cut = mser_cut(x if by is None else by, max_fraction)
    mean, err = truncated_mean(x, cut)
    return mean, err, cut
