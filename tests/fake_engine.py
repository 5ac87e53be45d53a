"""TEST INFRASTRUCTURE ONLY.  A stand-in for ``bayesian_optimization_b200._lib.Engine`` whose arithmetic is the numpy
oracle (oracle/gp_oracle.py), so that the HOST logic of the Python mirror -- parameter handling per estimation mode, the
L-BFGS-B restart loop and its RNG consumption, noise escalation, restricted likelihood, trends, multi-target plumbing,
pickling -- runs in the CPU test tier.  It is never importable from the product (it lives under tests/), and the product
has no switch that could select it: the tests monkeypatch ``gp.Engine``."""
from __future__ import annotations

import numpy as np

from bayesian_optimization_b200 import _lib
from oracle import gp_oracle as go


class FakeEngine:
    def __init__(self, device: int = 0):
        self.device = device
        self.N = self.D = 0
        self.gp = None
        self.restricted = False
        self.n_factor = 0

    # ---- configuration (no-ops) -----------------------------------------------------------------------
    def set_precision(self, prec):
        self.prec = prec

    def set_fast_kernel(self, g):
        pass

    def set_fast_products(self, p):
        pass

    def close(self):
        pass

    # ---- fit ---------------------------------------------------------------------------------------------
    def set_train(self, X, y):
        self.X, self.y = np.array(X, dtype=np.float64), np.array(y, dtype=np.float64).ravel()
        self.N, self.D = self.X.shape
        self.gp = None

    def factor(self, corr, theta, mode, par_last=0.0, noise_var=0.0, trend=_lib.TREND_CONSTANT, beta=None):
        self.n_factor += 1
        kw = dict(trend=trend, beta_fixed=beta)
        if mode == go.MODE_NOISY:
            kw.update(sigma2=par_last, noise_var=noise_var)
        elif mode == go.MODE_NOISE_ESTIM:
            kw.update(alpha=par_last)
        gp = go.fit_fixed(self.X, self.y, corr, theta, mode, **kw)
        self._last = (corr, np.array(theta, dtype=np.float64), mode, par_last, noise_var, trend, beta)
        self.restricted = False
        if not np.isfinite(gp.llf):
            self.gp = None
            return -np.inf, np.nan, np.nan, _lib.FIT_REJECTED
        self.gp, self._alpha = gp, (par_last if mode == go.MODE_NOISE_ESTIM else None)
        return gp.llf, gp.sigma2, gp.noise_var, _lib.FIT_OK

    def append(self, X_new, y_all):
        corr, theta, mode, par_last, noise_var, trend, beta = self._last
        self.set_train(np.vstack([self.X, np.asarray(X_new, dtype=np.float64)]), y_all)
        return self.factor(corr, theta, mode, par_last, noise_var, trend, beta)

    def llf_grad(self, n_par):
        return np.asarray(go.llf_grad(self.gp, self._alpha), dtype=np.float64).ravel()[:n_par]

    def factor_restricted(self, corr, theta, sigma2, noise_var=0.0, trend=_lib.TREND_CONSTANT, beta=None):
        self.n_factor += 1
        gp = go.fit_fixed_restricted(self.X, self.y, corr, theta, sigma2, noise_var, trend=trend, beta_fixed=beta)
        self.restricted = True
        if not np.isfinite(gp.llf):
            self.gp = None
            return -np.inf, _lib.FIT_REJECTED
        self.gp, self._rpar = gp, (corr, np.array(theta, dtype=np.float64), sigma2, noise_var, trend, beta)
        return gp.llf, _lib.FIT_OK

    def llf_grad_restricted(self, n_par):
        corr, theta, s2, nv, trend, beta = self._rpar
        _, g = go.fit_fixed_restricted(self.X, self.y, corr, theta, s2, nv, trend=trend, beta_fixed=beta, eval_grad=True, n_par=n_par)
        return np.asarray(g, dtype=np.float64)

    def state(self, what, p=1):
        g = self.gp
        return {_lib.STATE_L: lambda: g.L, _lib.STATE_GAMMA: lambda: g.gamma.ravel(), _lib.STATE_YT: lambda: g.Yt.ravel(),
                _lib.STATE_RHO: lambda: g.rho.ravel(), _lib.STATE_BETA: lambda: g.beta.ravel(),
                _lib.STATE_FT: lambda: (g.Ft if p > 1 else g.Ft.ravel()), _lib.STATE_G: lambda: (g.G if p > 1 else g.G.ravel()),
                _lib.STATE_LINV: lambda: np.linalg.inv(g.L)}[what]()

    # ---- predict / acquisition ---------------------------------------------------------------------------
    def predict(self, Xc, eval_mse=True):
        Xc = np.asarray(Xc, dtype=np.float64)
        if eval_mse:
            yh, ms = go.predict_chunked(self.gp, Xc, 2048)
            return yh.ravel(), ms.ravel()
        return go.predict_chunked(self.gp, Xc, 2048, eval_MSE=False).ravel(), None

    def gradient(self, Xc):
        Xc = np.asarray(Xc, dtype=np.float64)
        yh, ms = self.predict(Xc, True)
        g = [go.posterior_gradient(self.gp, x) for x in Xc]
        return yh, ms, np.array([a.ravel() for a, _ in g]), np.array([b.ravel() for _, b in g])

    def acq(self, Xc, acq_id, minimize, plugin, params, return_values=False, device_vals=None):
        yh, ms = self.predict(Xc, True)
        vals = np.array([go.acquisition(acq_id, yh, ms, self.gp.sigma2, plugin, par, minimize) for par in np.atleast_1d(params)])
        return vals.max(axis=1), vals.argmax(axis=1).astype(np.int64), (vals if return_values else None)

    def acq_grad(self, Xc, acq_id, minimize, plugin, param):
        yh, ms, ydx, mdx = self.gradient(Xc)
        out = [go.acquisition_dx(acq_id, yh[i], ms[i], ydx[i], mdx[i], self.gp.sigma2, plugin, param, minimize) for i in range(len(yh))]
        return np.array([v for v, _ in out]), np.array([d for _, d in out])
