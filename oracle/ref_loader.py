"""TEST INFRASTRUCTURE ONLY.  Loader for the *real* reference implementation.

Imports ``gpr.py / kernel.py / trend.py`` (and ``acquisition_fun.py``) of wangronin/Bayesian-Optimization
by file path from ``/root/reference`` under a throw-away stub package, bypassing
``bayes_optim/__init__.py`` (which needs pyDOE / sobol_seq / py_expression_eval -- not installed here,
SURVEY.md App. C).  Nothing is copied into this repository; the reference stays where it lies.

``/root/reference`` exists only in the build container.  On the GPU box ``available()`` is False and
every caller must fall back to the committed golden vectors.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("B200BO_REFERENCE_ROOT", "/root/reference")
_PKG = "_b200bo_refpkg"
_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "bayes_optim/surrogate/gaussian_process/gpr.py"))


def _load(modname: str, path: str):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return a namespace with the reference's GaussianProcess, kernels, trends and acquisition classes."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise ImportError(f"reference not found under {REF_ROOT}")
    gp_dir = os.path.join(REF_ROOT, "bayes_optim/surrogate/gaussian_process")
    # stub package tree:  _PKG, _PKG.surrogate, _PKG.surrogate.gaussian_process, _PKG.acquisition
    filters = list(warnings.filters)
    try:
        for name in (_PKG, f"{_PKG}.surrogate", f"{_PKG}.surrogate.gaussian_process", f"{_PKG}.acquisition"):
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        base = f"{_PKG}.surrogate.gaussian_process"
        for leaf in ("boundary_handling", "cma_es", "kernel", "trend", "gpr"):
            mod = _load(f"{base}.{leaf}", os.path.join(gp_dir, leaf + ".py"))
            setattr(sys.modules[base], leaf, mod)
        gpr = sys.modules[f"{base}.gpr"]
        trend = sys.modules[f"{base}.trend"]
        kernel = sys.modules[f"{base}.kernel"]
        sur = sys.modules[f"{_PKG}.surrogate"]
        sur.GaussianProcess = gpr.GaussianProcess
        sur.trend = trend

        class RandomForest:  # placeholder: acquisition_fun.py only uses the name in type hints
            pass

        sur.RandomForest = RandomForest
        # gpr.py:18 has just made every warning an error, which would turn the SyntaxWarning raised by
        # a docstring escape in acquisition_fun.py:109 into a SyntaxError; the real package import runs
        # with "ignore" prepended (surrogate/gaussian_process/__init__.py:22), do the same here.
        warnings.filters[:] = filters
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            acq = _load(
                f"{_PKG}.acquisition.acquisition_fun",
                os.path.join(REF_ROOT, "bayes_optim/acquisition/acquisition_fun.py"),
            )
    finally:
        # gpr.py:18 sets warnings.filterwarnings("error") process-wide at import; the package
        # __init__ (surrogate/gaussian_process/__init__.py:22) then prepends "ignore".  Reproduce the
        # package's net effect without leaking either into the test process.
        warnings.filters[:] = filters
    ns = types.SimpleNamespace(
        gpr=gpr,
        kernel=kernel,
        trend=trend,
        acquisition_fun=acq,
        GaussianProcess=gpr.GaussianProcess,
        constant_trend=trend.constant_trend,
        linear_trend=trend.linear_trend,
        quadratic_trend=trend.quadratic_trend,
        matern=kernel.matern,
        squared_exponential=kernel.squared_exponential,
        EI=acq.EI,
        EpsilonPI=acq.EpsilonPI,
        UCB=acq.UCB,
        MGFI=acq.MGFI,
    )
    _cache["ns"] = ns
    return ns


def fixed_theta_fit(gp, X, y, theta, sigma2=None):
    """The reference's own fit() with the optimiser loop cut out (SURVEY.md §8c oracle recipe (1)):
    gpr.py:375-376 (_check_data), :1183-1188 (final likelihood evaluation filling env),
    :402-415 (attribute copy + compute_beta_gamma)."""
    import numpy as np

    gp._check_data(X, y)
    env = {}
    if gp.estimation_mode == "noiseless":
        par = np.asarray(theta, dtype=float)
    elif gp.estimation_mode == "noisy":
        par = np.r_[np.asarray(theta, dtype=float), float(sigma2)]
    else:  # noise_estim: last entry is alpha
        par = np.r_[np.asarray(theta, dtype=float), float(sigma2)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        llf = gp.log_likelihood_concentrated(par, env)
    if not np.isfinite(llf):
        return llf
    gp.theta_ = np.asarray(theta, dtype=float)
    gp.noise_var = env["noise_var"]
    gp.sigma2 = np.atleast_1d(env["sigma2"])
    gp.rho, gp.Yt, gp.C = env["rho"], env["Yt"], env["C"]
    if gp.estimate_trend:
        gp.Ft, gp.G, gp.Q = env["Ft"], env["G"], env["Q"]
    gp.compute_beta_gamma()
    gp.is_fitted = True
    gp.log_likelihood_ = llf
    return llf
