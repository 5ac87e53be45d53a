"""Synthetic benchmark workloads: the BASELINE.json configs with the canonical inputs of SURVEY.md §8d.
(oracle/gp_oracle.py holds an independent copy of the same recipe for the CPU side; tests/test_workloads.py
checks that the two agree bit for bit.)"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass(frozen=True)
class Workload:
    name: str
    N: int
    D: int
    M_total: int  # candidates over the whole job at the config's GPU count
    gpus: int  # GPU count the config is quoted on
    corr: str
    acq: str
    q: int
    nugget: float
    describe: str

    @property
    def M_per_gpu(self) -> int:
        return self.M_total // self.gpus


WORKLOADS = {
    "C2": Workload("C2", 1024, 8, 1_000_000, 1, "squared_exponential", "EI", 1, 1e-6,
                   "GP fit N=1024 D=8 RBF fp64 + EI over 1e6 candidates, 1xB200"),
    "C3": Workload("C3", 4096, 16, 10_000_000, 8, "matern52", "MGFI", 32, 1e-6,
                   "ParallelBO MGFI q=32, N=4096 D=16 Matern-5/2 ARD, 1e7 candidates over 8xB200"),
    "C4": Workload("C4", 8192, 32, 20_000_000, 8, "squared_exponential", "UCB", 32, 1e-2,
                   "NoisyBO UCB, N=8192 D=32 RBF + diagonal noise, 2e7 candidates over 8xB200"),
    "C5": Workload("C5", 2048, 64, 4_000_000, 4, "squared_exponential", "MGFI", 32, 1e-6,
                   "AnnealingBO MGFI t sweep, N=2048 D=64, 4e6 candidates over 4xB200"),
}


def canonical_problem(N: int, D: int):
    """X ~ U(0,1)^(N x D) seed 42; y = standardise(sum_j sin(2 pi x_j)/sqrt(D) + 0.1 N(0,1) seed 43) (mirrors
    bayes_optim/base.py:437); ARD theta_j = 10^linspace(-0.3, 0.7, D) * 8/D."""
    X = np.random.default_rng(42).uniform(0, 1, (N, D))
    y_raw = np.sin(2 * np.pi * X).sum(axis=1) / np.sqrt(D) + 0.1 * np.random.default_rng(43).standard_normal(N)
    y = (y_raw - y_raw.mean()) / y_raw.std()
    theta = 10.0 ** np.linspace(-0.3, 0.7, D) * (8.0 / D)
    return X, y, theta


def canonical_candidates(M: int, D: int, shard: int = 0, out: np.ndarray = None) -> np.ndarray:
    rng = np.random.default_rng(1000 + shard)
    if out is None:
        return rng.uniform(0, 1, (M, D))
    rng.random(out=out)
    return out


def acquisition_params(w: Workload) -> np.ndarray:
    """q criterion parameters per config: C3 ParallelBO t_i = exp(ln 2 + 0.5 xi) (bayes_opt.py:85);
    C4 alpha_i = 1/(1+exp(4*0.5-2+0.6 xi)) (:89); C5 AnnealingBO exponential t schedule 2 -> 0.1 (:127-130);
    xi from numpy's global RNG seeded with 42."""
    rs = np.random.RandomState(42)
    if w.name == "C3":
        return np.array([np.exp(np.log(2.0) + 0.5 * rs.randn()) for _ in range(w.q)])
    if w.name == "C4":
        return np.array([1 / (1 + np.exp((0.5 * 4 - 2) + 0.6 * rs.randn())) for _ in range(w.q)])
    if w.name == "C5":
        a = (0.1 / 2.0) ** (1.0 / w.q)
        return 2.0 * a ** np.arange(w.q)
    return np.zeros(w.q)


ACQ_IDS = {"EI": _lib.ACQ_EI, "PI": _lib.ACQ_PI, "UCB": _lib.ACQ_UCB, "MGFI": _lib.ACQ_MGFI}
