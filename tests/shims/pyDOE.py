"""Stand-in for pyDOE.lhs (test infrastructure, see README.md): Latin hypercube sample in [0, 1]^n."""
import numpy as np


def lhs(n, samples=None, criterion=None, iterations=None):
    samples = n if samples is None else int(samples)
    cut = (np.arange(samples)[:, None] + np.random.rand(samples, n)) / samples
    for j in range(n):
        cut[:, j] = cut[np.random.permutation(samples), j]
    return cut
