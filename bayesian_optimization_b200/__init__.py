"""bayesian_optimization_b200 -- B200-native GP surrogate + acquisition engine.

Drop-in for the one hot path of wangronin/Bayesian-Optimization:
``bayes_optim.surrogate.GaussianProcess.{fit,predict}`` and ``bayes_optim.acquisition.{EI,PI,UCB,MGFI}``
(see DESIGN.md / INTEGRATION.md).  Everything numeric runs in libb200bo.so (hand-written sm_100a CUDA behind
the C ABI of include/b200bo.h); importing this package does not need a GPU, computing does.
"""
from . import _lib
from ._lib import B200BOError, Engine
from .acquisition import EI, MGFI, PI, UCB, AcquisitionFunction, EpsilonPI, ImprovementBased
from .candidates import argmax_candidates, sample_candidates
from .gp import GaussianProcess, resolve_corr
from .trend import BasisExpansionTrend, constant_trend, linear_trend, quadratic_trend

__all__ = [
    "GaussianProcess", "Engine", "B200BOError", "EI", "PI", "EpsilonPI", "UCB", "MGFI",
    "AcquisitionFunction", "ImprovementBased", "argmax_candidates", "sample_candidates", "constant_trend", "linear_trend", "quadratic_trend", "BasisExpansionTrend", "resolve_corr",
]
__version__ = "0.1.0"
