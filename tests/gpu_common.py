"""Helpers shared by the -m gpu parity tests: build a device model for a golden case."""
import functools

import numpy as np

import bayesian_optimization_b200 as b2
from oracle import gp_oracle as go


def matern(theta, X, nu=1.5):  # name-recognised stand-in for the reference kernel callable
    raise AssertionError


CORR_ARG = {
    go.CORR_RBF: "squared_exponential",
    go.CORR_MATERN32: "matern",
    go.CORR_MATERN52: functools.partial(matern, nu=2.5),
    go.CORR_MATERN12: functools.partial(matern, nu=0.5),
    go.CORR_ABSEXP: "absolute_exponential",
    go.CORR_CUBIC: "cubic",
    go.CORR_GENEXP: "generalized_exponential",
}


def device_gp(c, D, beta=None):
    """GaussianProcess configured like tests/golden/make_golden.py:make_gp"""
    mode, ok = int(c["mode"]), bool(c["ok"])
    mean = b2.constant_trend(D) if ok else b2.constant_trend(D, beta=float(np.ravel(c["beta_in"])[0]) if beta is None else beta)
    nt = D + 1 if int(c["corr"]) == go.CORR_GENEXP else D  # generalized_exponential: theta carries the exponent
    kw = dict(mean=mean, corr=CORR_ARG[int(c["corr"])], thetaL=[1e-5] * nt, thetaU=[1e2] * nt)
    if mode == go.MODE_NOISELESS:
        kw.update(nugget=None)
    elif mode == go.MODE_NOISY:
        kw.update(nugget=float(c["nugget"]))
    else:
        kw.update(nugget=float(c["nugget"]), noise_estim=True)
    return b2.GaussianProcess(**kw)


def fit_case(c, X=None, y=None):
    X = c["X"] if X is None else X
    y = c["y"] if y is None else y
    gp = device_gp(c, X.shape[1])
    last = None if int(c["mode"]) == go.MODE_NOISELESS else float(c["par_last"])
    llf = gp.fit_fixed(X, y, c["theta"], last)
    return gp, llf


def oracle_case(c, X=None, y=None):
    X = c["X"] if X is None else X
    y = c["y"] if y is None else y
    mode = int(c["mode"])
    kw = {}
    if not bool(c["ok"]):
        kw["beta_fixed"] = [float(np.ravel(c["beta_in"])[0])]
    if mode == go.MODE_NOISY:
        kw.update(sigma2=float(c["par_last"]), noise_var=float(c["nugget"]))
    elif mode == go.MODE_NOISE_ESTIM:
        kw.update(alpha=float(c["par_last"]))
    return go.fit_fixed(X, y, int(c["corr"]), c["theta"], mode, **kw)
