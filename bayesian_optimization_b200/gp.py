"""``GaussianProcess`` -- drop-in for ``bayes_optim.surrogate.GaussianProcess`` backed by libb200bo.so.

Same constructor keywords, ``fit / predict / update`` signatures, attributes and error behaviour as the
reference class (bayes_optim/surrogate/gaussian_process/gpr.py:78-1248); all arithmetic runs on the B200
through the C ABI of include/b200bo.h.  There is no CPU path.

Attribute map (reference -> here): ``X, y, theta_, sigma2, noise_var, gamma, C, Yt, Ft, G, Q, rho,
mean.beta, is_fitted, log_likelihood_, par, estimation_mode`` -- the large ones (``C, gamma, Yt, Ft, rho,
Q``) are fetched lazily from device memory.
"""
from __future__ import annotations

import functools
import warnings
from typing import Optional

import numpy as np
from sklearn.utils import check_array, check_random_state, check_X_y

from . import _lib
from ._lib import Engine
from .trend import BasisExpansionTrend, constant_trend

_CORR_IDS = {
    "squared_exponential": _lib.CORR_RBF,
    "matern": _lib.CORR_MATERN32,  # the string API is nu = 1.5 only (kernel.py:159, gpr.py:206)
    "matern12": _lib.CORR_MATERN12,
    "matern32": _lib.CORR_MATERN32,
    "matern52": _lib.CORR_MATERN52,
    "absolute_exponential": _lib.CORR_ABSEXP,
    "cubic": _lib.CORR_CUBIC,
    "generalized_exponential": _lib.CORR_GENEXP,  # theta carries the exponent as its last entry (kernel.py:367-373)
}
_TREND_IDS = {"constant_trend": _lib.TREND_CONSTANT, "linear_trend": _lib.TREND_LINEAR, "quadratic_trend": _lib.TREND_QUADRATIC}
_MODES = {"noiseless": _lib.MODE_NOISELESS, "noisy": _lib.MODE_NOISY, "noise_estim": _lib.MODE_NOISE_ESTIM}


def resolve_corr(corr) -> int:
    """Map the reference's ``corr`` argument (gpr.py:201-207, :1199-1208) to a device kernel id.  Callables
    cannot run on the GPU; the one callable idiom the reference needs -- ``functools.partial(matern, nu=..)``,
    the only route to Matern-5/2 -- is recognised by name."""
    if isinstance(corr, str):
        if corr in _CORR_IDS:
            return _CORR_IDS[corr]
        raise ValueError("corr should be one of %s or callable, %s was given." % (list(_CORR_IDS), corr))
    if isinstance(corr, functools.partial) and getattr(corr.func, "__name__", "") == "matern":
        nu = corr.keywords.get("nu", 1.5)
        if not float(nu) > 0:
            raise ValueError(f"matern nu={nu} must be positive")
        # 0.5 / 1.5 / 2.5 have closed forms (kernel.py:189-200); any other nu goes through K_nu (kernel.py:201-207)
        return {0.5: _lib.CORR_MATERN12, 1.5: _lib.CORR_MATERN32, 2.5: _lib.CORR_MATERN52}.get(float(nu), _lib.CORR_MATERN_NU)
    name = getattr(corr, "__name__", None)
    if name in _CORR_IDS:
        return _CORR_IDS[name]
    raise ValueError(f"corr={corr!r} is a callable without a device kernel")


class GaussianProcess:
    """The Gaussian Process model class (B200).  Keyword-compatible with gpr.py:211-228."""

    _optimizer_types = ["BFGS", "CMA"]
    _likelihood_functions = ["concentrated", "restricted"]

    def __init__(self, mean=None, corr="squared_exponential", theta0=None, thetaL=None, thetaU=None, sigma2=None,
                 nugget=1e-6, noise_estim=False, optimizer="BFGS", likelihood="concentrated", random_start=1,
                 wait_iter=5, eval_budget=None, random_state=None, verbose=False, device: int = 0,
                 precision: str = "fp64"):
        # gpr.py:229-277; ``device`` and ``precision`` are the only additions.  precision="fast": predict /
        # acquisition arg-max run the tcgen05 tensor-core pass (arg-max still exact via the fp64 re-score)
        self.mean = mean
        self.corr = corr
        self.sigma2 = sigma2
        self.verbose = verbose
        self.corr_type = corr
        self.is_fitted = False
        self.theta0 = np.array(theta0).flatten() if theta0 is not None else None
        self.thetaL = np.array(thetaL).flatten()
        self.thetaU = np.array(thetaU).flatten()
        if not (np.isfinite(self.thetaL.astype(float)).all() and np.isfinite(self.thetaU.astype(float)).all()):
            raise ValueError("all bounds are required finite.")
        self.optimizer = optimizer
        self.random_start = random_start
        self.random_state = random_state
        self.wait_iter = wait_iter
        self.eval_budget = eval_budget
        self.nugget = nugget
        self.noise_var = np.atleast_1d(nugget) if nugget else 0
        self.noise_estim = noise_estim
        self.noisy = self.noise_var or self.noise_estim
        if not self.noisy:
            self.estimation_mode = "noiseless"
        elif self.noise_estim:
            self.estimation_mode = "noise_estim"
        else:
            self.estimation_mode = "noisy"
        assert likelihood in self._likelihood_functions
        self.likelihood = likelihood
        if self.mean is None:
            self.mean = constant_trend(len(self.thetaU), beta=0)  # simple kriging, gpr.py:269-270
        if _is_basis_trend(self.mean):
            self.mean_type = "basis_expansion"
            self.estimate_trend = True if self.mean.beta is None else False
        else:
            raise ValueError("only BasisExpansionTrend means have a device implementation")
        self.device = int(device)
        if precision not in ("fp64", "fast"):
            raise ValueError("precision should be 'fp64' or 'fast', %s was given." % precision)
        self.precision = precision
        self._engine: Optional[Engine] = None
        self._cache = {}
        self._pool = []      # extra engine handles of the concurrent hyper-parameter restarts: [engine, training-set generation]
        self._train_gen = 0
        self._sub = None  # multi-target y (N, k > 1): one single-target model per column (same R, own Yt / rho / gamma)

    # ---- engine plumbing ---------------------------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if getattr(self, "_sub", None):
            return self._sub[0].engine
        if self._engine is None:
            self._engine = Engine(self.device)  # raises without libb200bo.so / a B200
            self._engine.set_precision(_lib.PREC_FAST if getattr(self, "precision", "fp64") == "fast" else _lib.PREC_FP64)
            if getattr(self, "X", None) is not None:
                self._engine.set_train(self.X, self.y[:, 0])
                if self.is_fitted:
                    self._refactor()
        return self._engine

    def _engine_pool(self, n: int):
        """``n`` engine handles that hold this model's training set, the primary one first.  The extra handles (own
        CUDA stream, own factorisation buffers) let the restarts of the hyper-parameter search run concurrently; they
        are kept for the next fit() of a BO loop and follow the training set lazily."""
        if self._sub or n <= 1:
            return [self.engine]
        pool = [e for e in getattr(self, "_pool", [])][: n - 1]
        while len(pool) < n - 1:
            pool.append([Engine(self.device), -1])
        for slot in pool:
            if slot[1] != self._train_gen:
                slot[0].set_train(self.X, self.y[:, 0])
                slot[1] = self._train_gen
        self._pool = pool
        return [self.engine] + [slot[0] for slot in pool]

    def close_pool(self):
        """release the extra engine handles of the hyper-parameter search (3 N^2 float64 matrices each)"""
        for slot in getattr(self, "_pool", []):
            slot[0].close()
        self._pool = []

    def __getstate__(self):
        # dill/pickle (BaseBO.save, base.py:499-519; joblib workers, bayes_opt.py:108-111): drop the device
        # handles and the lazily fetched arrays; the deterministic factorisation is redone on first use.
        st = self.__dict__.copy()
        st["_engine"] = None
        st["_cache"] = {}
        st["_pool"] = []
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)

    def get_params(self, deep=True):  # sklearn BaseEstimator surface used by clone()
        keys = ["mean", "corr", "theta0", "thetaL", "thetaU", "sigma2", "nugget", "noise_estim", "optimizer",
                "likelihood", "random_start", "wait_iter", "eval_budget", "random_state", "verbose"]
        return {k: getattr(self, k if k != "corr" else "corr_type") for k in keys}

    # ---- checks (gpr.py:279-310, :1199-1248) -------------------------------------------------------
    def _check_params(self):
        self._corr_id = resolve_corr(self.corr_type)
        # general-nu Matern: nu is an argument of the callable upstream; the device takes it as one more entry behind theta
        self._corr_extra = float(self.corr_type.keywords["nu"]) if self._corr_id == _lib.CORR_MATERN_NU else None
        if self.thetaL is not None and self.thetaU is not None:
            if self.thetaL.size != self.thetaU.size:
                raise ValueError("thetaL and thetaU must have the same length.")
            if self.theta0 is not None and self.theta0.size != self.thetaL.size:
                raise ValueError("theta0, thetaL, and thetaU must have the same length.")
            if np.any(self.thetaL <= 0) or np.any(self.thetaU < self.thetaL):
                raise ValueError("The bounds must satisfy O < thetaL <= thetaU.")
        self.verbose = bool(self.verbose)
        if self.optimizer not in self._optimizer_types:
            raise ValueError("optimizer should be one of %s" % self._optimizer_types)
        self.random_start = int(self.random_start)

    def _check_data(self, X, y):
        X, y = check_X_y(X, y, multi_output=True, y_numeric=True)
        if len(y.shape) == 1:
            y = y.reshape(-1, 1)
        if y.shape[1] != 1:
            return self._check_data_multi(X, y)
        self._sub = None
        tname = type(self.mean).__name__
        if tname not in _TREND_IDS:
            raise NotImplementedError("trend %s has no device implementation (constant, linear, quadratic do)" % tname)
        self._trend_id = _TREND_IDS[tname]
        self._p = int(self.mean.n_dim)
        if self._p > 64:
            raise NotImplementedError("at most 64 trend basis functions are supported on device")
        self.X, self.y = np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
        self._check_params()
        self._cache = {}
        self._train_gen = getattr(self, "_train_gen", 0) + 1
        if self.estimate_trend and self._p > self.X.shape[0]:  # gpr.py:298-309: beta cannot be estimated from fewer rows
            raise Exception("Ordinary least squares problem is undetermined n_samples=%d must be greater than the "
                            "regression model size p=%d." % (self.X.shape[0], self._p))
        self.engine.set_train(self.X, self.y[:, 0])
        if self.estimate_trend:
            self.F = self.mean.F(self.X)

    # ---- multi-target y (N, k): gpr.py keeps Yt, rho, gamma, beta, sigma2 per column over ONE factorisation ---------
    # (:799 Yt = L^-1 y, :806-808 rho, :934-979 sigma2 (k,), likelihood vector summed at :1040, predict :490, :502-505).
    # Here every column gets its own single-target model / device handle; each repeats the deterministic factorisation
    # (k is the number of objectives: 2-3), so the per-column states are exactly those of the reference.
    def _check_data_multi(self, X, y):
        import copy

        self.X, self.y = np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
        self._check_params()
        self._cache = {}
        if self.estimate_trend:
            self.F = self.mean.F(self.X)
        if self._sub is None or len(self._sub) != y.shape[1]:
            self._sub = []
            for _ in range(y.shape[1]):
                g = GaussianProcess(mean=copy.deepcopy(self.mean), corr=self.corr, theta0=self.theta0, thetaL=self.thetaL,
                                    thetaU=self.thetaU, nugget=self.nugget, noise_estim=self.noise_estim,
                                    optimizer=self.optimizer, likelihood=self.likelihood, device=self.device,
                                    precision=self.precision)
                self._sub.append(g)
        for t, g in enumerate(self._sub):
            g._check_data(self.X, self.y[:, t])

    def _sync_sub(self):
        for g in self._sub:
            g.estimation_mode, g.noise_var = self.estimation_mode, self.noise_var

    def _llf_multi(self, par, env, eval_grad):
        self._sync_sub()
        n_par = np.size(par)
        envs = [{} for _ in self._sub]
        outs = [g.log_likelihood_concentrated(par, e, eval_grad) for g, e in zip(self._sub, envs)]
        llfs = [o[0] if eval_grad else o for o in outs]
        self._cache = {}
        if not np.all(np.isfinite(llfs)):  # any(log_likelihood > 0) or a failed factorisation: gpr.py:981-982
            return (-np.inf, np.zeros((n_par, 1))) if eval_grad else -np.inf
        if env is not None:
            env["sigma2"] = np.array([np.atleast_1d(e["sigma2"])[0] for e in envs])
            nv = np.array([np.atleast_1d(e["noise_var"])[0] for e in envs])
            env["noise_var"] = nv if self.estimation_mode == "noise_estim" else envs[0]["noise_var"]
        llf = float(np.sum(llfs))  # gpr.py:1040
        if eval_grad:
            # sum of the per-target gradients (upstream's k > 1 formula mixes gamma gamma^T of ALL targets with every
            # sigma2_t, gpr.py:997-1010; the optimum it steers to is the same stationary point of the summed likelihood)
            return llf, np.sum([np.asarray(o[1], dtype=np.float64).reshape(-1, 1) for o in outs], axis=0)
        return llf

    # ---- likelihood at given hyper-parameters (gpr.py:920-991) -----------------------------------------
    def _split_par(self, par):
        par = np.asarray(par, dtype=np.float64).ravel()
        if self.estimation_mode == "noiseless":
            return par, 0.0
        return par[:-1], float(par[-1])

    def _theta_dev(self, theta):
        theta = np.asarray(theta, dtype=np.float64).ravel()
        return theta if getattr(self, "_corr_extra", None) is None else np.r_[theta, self._corr_extra]

    def _beta_fixed(self):
        return None if self.estimate_trend else np.asarray(self.mean.beta, dtype=np.float64).ravel()

    def _factor(self, par, engine=None):
        theta, last = self._split_par(par)
        nv = float(np.atleast_1d(self.noise_var)[0]) if self.estimation_mode == "noisy" else 0.0
        llf, s2, nvo, status = (engine or self.engine).factor(self._corr_id, self._theta_dev(theta), _MODES[self.estimation_mode],
                                                              last, nv, self._trend_id, self._beta_fixed())
        self._cache = {}
        return llf, s2, nvo, status

    def _likelihood_on(self, engine, par, restricted=False, env=None, eval_grad=False):
        """the likelihood (and gradient) at ``par`` evaluated on a given engine handle -- what one restart of the
        hyper-parameter search calls; the public methods below run it on the model's primary handle"""
        n_par = np.size(par)
        fail = (-np.inf, np.zeros((n_par, 1))) if eval_grad else -np.inf
        if restricted:
            if self._sub:
                raise NotImplementedError("the restricted likelihood is implemented for one target")
            theta, s2, nv = self._split_par_restricted(par)
            llf, status = engine.factor_restricted(self._corr_id, self._theta_dev(theta), s2, nv, self._trend_id, self._beta_fixed())
            self._cache = {}
            if status != _lib.FIT_OK:
                return fail
            if env is not None:
                env["sigma2"] = s2
                env["noise_var"] = nv
            return (llf, engine.llf_grad_restricted(n_par).reshape(-1, 1)) if eval_grad else llf
        if self._sub:
            return self._llf_multi(par, env, eval_grad)
        llf, s2, nvo, status = self._factor(par, engine)
        if status != _lib.FIT_OK:
            return fail
        if env is not None:
            env["sigma2"] = np.atleast_1d(s2)
            env["noise_var"] = nvo
        return (llf, engine.llf_grad(n_par)) if eval_grad else llf

    def log_likelihood_concentrated(self, par, env=None, eval_grad=False):
        """Concentrated log-likelihood at ``par`` on the device; -inf when the factorisation fails or the
        value is positive (gpr.py:981-982).  ``env`` receives sigma2 / noise_var like the reference's."""
        return self._likelihood_on(None if self._sub else self.engine, par, False, env, eval_grad)

    def _split_par_restricted(self, par):
        """gpr.py:826-835: (theta, sigma2, noise_var) per estimation mode"""
        par = np.asarray(par, dtype=np.float64).ravel()
        if self.estimation_mode == "noiseless":
            return par[:-1], float(par[-1]), 0.0
        if self.estimation_mode == "noisy":
            return par[:-1], float(par[-1]), float(np.atleast_1d(self.noise_var)[0])
        return par[:-2], float(par[-2]), float(par[-1])

    def log_likelihood_restricted(self, par, env=None, eval_grad=False):
        """Restricted (REML) log-likelihood at ``par`` on the device (gpr.py:813-918); -inf when the factorisation
        fails (:842-847) or exp(llf) > 1 (:872-875).  ``env`` receives sigma2 / noise_var as upstream (:904-906)."""
        return self._likelihood_on(self.engine, par, True, env, eval_grad)

    def _refactor(self):
        if getattr(self, "_restricted_par", None) is not None:
            _, status = self.engine.factor_restricted(self._corr_id, self._theta_dev(self.theta_), self._restricted_par[0],
                                                      self._restricted_par[1], self._trend_id, self._beta_fixed())
            if status != _lib.FIT_OK:  # pragma: no cover - the same inputs factored before
                raise RuntimeError("re-factorisation of a fitted model failed")
            return
        par = self._par_vector()
        _, _, _, status = self._factor(par)
        if status != _lib.FIT_OK:  # pragma: no cover - the same inputs factored before
            raise RuntimeError("re-factorisation of a fitted model failed")

    def _par_vector(self):
        if self.estimation_mode == "noiseless":
            return np.asarray(self.theta_, dtype=np.float64)
        return np.r_[np.asarray(self.theta_, dtype=np.float64), self._par_last]

    # ---- fit -------------------------------------------------------------------------------------------
    def fit_fixed(self, X, y, theta, par_last=None):
        """Fit at GIVEN hyper-parameters: the reference's fit() with the optimiser loop removed, i.e.
        ``_check_data`` (gpr.py:375-376) + the final likelihood evaluation (:1183-1188) + the attribute copy
        and ``compute_beta_gamma`` (:402-415).  ``par_last`` = sigma2 ("noisy") or alpha ("noise_estim").
        Returns the log-likelihood; the model is fitted iff it is finite."""
        self.random_state = check_random_state(self.random_state)
        self._check_data(X, y)
        theta = np.asarray(theta, dtype=np.float64).ravel()
        par = theta if self.estimation_mode == "noiseless" else np.r_[theta, float(par_last)]
        env = {}
        llf = self.log_likelihood_concentrated(par, env)
        self.log_likelihood_ = llf
        if not np.isfinite(llf):
            self.is_fitted = False
            return llf
        self._adopt(theta, None if self.estimation_mode == "noiseless" else float(par_last), env)
        return llf

    def fit_fixed_restricted(self, X, y, theta, sigma2, noise_var=None):
        """``fit_fixed`` for likelihood="restricted": parameters (theta, sigma2[, noise_var]) as gpr.py:826-835 reads
        them (noise_var: 0 in "noiseless" mode, the nugget in "noisy" mode, given in "noise_estim" mode)."""
        assert self.likelihood == "restricted"
        self.random_state = check_random_state(self.random_state)
        self._check_data(X, y)
        theta = np.asarray(theta, dtype=np.float64).ravel()
        par = np.r_[theta, float(sigma2)] if self.estimation_mode != "noise_estim" else np.r_[theta, float(sigma2), float(noise_var)]
        env = {}
        llf = self.log_likelihood_restricted(par, env)
        self.log_likelihood_ = llf
        if not np.isfinite(llf):
            self.is_fitted = False
            return llf
        self._adopt(theta, float(sigma2), env)
        return llf

    def _adopt(self, theta, par_last, env):
        self.theta_ = np.asarray(theta, dtype=np.float64)
        self._par_last = par_last
        self.noise_var = env["noise_var"]
        self.sigma2 = np.atleast_1d(env["sigma2"])
        self._restricted_par = ((float(self.sigma2[0]), float(np.atleast_1d(self.noise_var)[0]))
                                if self.likelihood == "restricted" else None)
        assert len(self.sigma2) == self.y.shape[1]
        if self._sub:
            for t, g in enumerate(self._sub):
                e = {"sigma2": np.atleast_1d(self.sigma2[t]), "noise_var": np.atleast_1d(np.atleast_1d(self.noise_var)[t if np.size(self.noise_var) > 1 else 0])}
                g._adopt(theta, par_last, e)
            if self.estimate_trend:
                self.mean.beta = np.hstack([np.asarray(g.mean.beta).reshape(-1, 1) for g in self._sub])  # (p, k)
            self.is_fitted = True
            return
        if self.estimate_trend:
            self.mean.beta = self.engine.state(_lib.STATE_BETA, self._p)  # gpr.py:787
        self.is_fitted = True

    def fit(self, X, y):
        """gpr.py:355-417.  Hyper-parameters by maximum likelihood on the device likelihood + gradient with the
        reference's host L-BFGS-B restart loop; see ``hyperopt.optimize_hyperparameter``."""
        from .hyperopt import optimize_hyperparameter

        self.random_state = check_random_state(self.random_state)
        self._check_data(X, y)
        while True:
            self.par, self.log_likelihood_, env = optimize_hyperparameter(self)
            if np.isinf(self.log_likelihood_):
                print("Invalid likelihood value. Increasing nugget...")  # gpr.py:390
                if self.estimation_mode == "noiseless":
                    self.estimation_mode = "noisy"
                    self.noise_var = 1e-5
                else:
                    self.noise_var *= 10
            else:
                break
        last = None
        if "sigma2" in self.par:
            last = float(self.par["sigma2"][0])
        if "alpha" in self.par:
            last = float(self.par["alpha"][0])
        self._adopt(self.par["theta"], last, env)
        return self

    def update(self, X, y, reoptimize: bool = True):
        """``update(X, y)`` is ``fit(X, y)`` upstream, with a "TODO: implement the rank-one update" next to it
        (gpr.py:419-422); that stays the default.  ``reoptimize=False`` keeps the hyper-parameters and, when ``X`` is
        the current training set followed by new rows, APPENDS those rows to the factorisation on the device
        (``b200bo_append``: O(m N^2) instead of O(N^3)); ``y`` holds every target (BaseBO re-standardises them whenever a
        point arrives, base.py:437).  Anything else -- reordered rows, a trend with p > 1, several targets, a failed
        update -- falls back to a fixed-parameter refit, then to ``fit``."""
        if reoptimize or not self.is_fitted or self._sub or getattr(self, "_trend_id", 0) != _lib.TREND_CONSTANT:
            return self.fit(X, y)
        Xn, yn = check_X_y(X, y, multi_output=True, y_numeric=True)
        yn = yn.reshape(len(yn), -1)
        if yn.shape[1] != 1:
            return self.fit(X, y)
        N0, D = self.X.shape
        Xn = np.ascontiguousarray(Xn, dtype=np.float64)
        appendable = (Xn.shape[1] == D and Xn.shape[0] > N0 and np.array_equal(Xn[:N0], self.X) and self._engine is not None
                      and self._engine.N == N0)
        if appendable:
            llf, s2, nv, status = self._engine.append(Xn[N0:], yn[:, 0])
            if status == _lib.FIT_OK:
                self.X, self.y = Xn, np.ascontiguousarray(yn, dtype=np.float64)
                self._cache = {}
                self._train_gen += 1
                self.log_likelihood_ = llf
                if self.estimate_trend:
                    self.F = self.mean.F(self.X)
                env = {"sigma2": np.atleast_1d(s2), "noise_var": nv}
                if self.likelihood == "restricted":
                    env = {"sigma2": self._restricted_par[0], "noise_var": self._restricted_par[1]}
                self._adopt(self.theta_, self._par_last, env)
                return self
        if self.likelihood == "restricted":
            llf = self.fit_fixed_restricted(Xn, yn, self.theta_, self._restricted_par[0], self._restricted_par[1])
        else:
            llf = self.fit_fixed(Xn, yn, self.theta_, self._par_last)
        return self if np.isfinite(llf) else self.fit(X, y)

    # ---- lazily fetched state ------------------------------------------------------------------------
    def _state(self, key, what, shape=None):
        if self._sub and key not in self._cache:
            if key == "C":  # one factorisation for all targets
                self._cache[key] = self._sub[0]._state(key, what, shape)
            else:           # gamma, Yt, rho: one column per target
                self._cache[key] = np.hstack([g._state(key, what, (-1, 1)) for g in self._sub])
        if key not in self._cache:
            a = self.engine.state(what)
            self._cache[key] = a if shape is None else a.reshape(shape)
        return self._cache[key]

    @property
    def C(self):
        return self._state("C", _lib.STATE_L)

    @property
    def gamma(self):
        return self._state("gamma", _lib.STATE_GAMMA, (-1, 1))

    @property
    def Yt(self):
        return self._state("Yt", _lib.STATE_YT, (-1, 1))

    @property
    def rho(self):
        return self._state("rho", _lib.STATE_RHO, (-1, 1))

    @property
    def Ft(self):
        if not self.estimate_trend:
            return None
        if self._sub:
            return self._sub[0].Ft
        if "Ft" not in self._cache:
            self._cache["Ft"] = self.engine.state(_lib.STATE_FT, self._p).reshape(-1, self._p)
        return self._cache["Ft"]

    @property
    def G(self):
        if not self.estimate_trend:
            return None
        if self._sub:
            return self._sub[0].G
        if "G" not in self._cache:
            self._cache["G"] = self.engine.state(_lib.STATE_G, self._p).reshape(self._p, self._p)
        return self._cache["G"]

    @property
    def Q(self):
        return np.linalg.solve(self.G.T, self.Ft.T).T if self.estimate_trend else None  # Ft = Q G

    # ---- predict (gpr.py:424-535) ----------------------------------------------------------------------
    def predict(self, X, eval_MSE=False, batch_size=None):
        assert hasattr(self, "X")
        X = check_array(X)
        n_features = self.X.shape[1]
        self._check_params()
        if X.shape[1] != n_features:
            raise ValueError(
                "The number of features in X (X.shape[1] = %d) should match the number of features used "
                "for fit() which is %d." % (X.shape[1], n_features)
            )
        if batch_size is not None and (type(batch_size) is not int or batch_size <= 0):
            raise Exception("batch_size must be a positive integer")
        # batch_size only bounds host memory in the reference (and its branch is dead on Python 3,
        # gpr.py:520); the engine streams candidates in SM-count-sized chunks regardless.
        if self._sub:  # (M, k) columns; the MSE factor is shared, sigma2 is per target (gpr.py:502-505)
            outs = [g.predict(X, eval_MSE=eval_MSE) for g in self._sub]
            if eval_MSE:
                return np.hstack([o[0] for o in outs]), np.hstack([o[1] for o in outs])
            return np.hstack(outs)
        yhat, mse = self.engine.predict(X, eval_mse=bool(eval_MSE))
        if eval_MSE:
            return yhat.reshape(-1, 1), mse.reshape(-1, 1)
        return yhat.reshape(-1, 1)


    # ---- posterior gradient (gpr.py:537-576) --------------------------------------------------------------
    def gradient(self, x):
        """d yhat / dx and d MSE / dx at ONE point: ((D, 1), (D, 1)) exactly as the reference returns them."""
        if self._sub:
            raise NotImplementedError("gradient is implemented for one target")
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        n_eval, nf = x.shape
        if nf != self.X.shape[1]:
            raise Exception("x does not have the right size!")
        if n_eval != 1:
            raise Exception("x must be a vector!")
        _, _, ydx, mdx = self.engine.gradient(x)
        return ydx.reshape(-1, 1), mdx.reshape(-1, 1)

    def gradient_batch(self, X):
        """The same for every row of X in one device pass: (yhat (M,1), mse (M,1), y_dx (M,D), mse_dx (M,D))."""
        X = check_array(X)
        if X.shape[1] != self.X.shape[1]:
            raise ValueError("The number of features in X should match the number of features used for fit().")
        yh, ms, ydx, mdx = self.engine.gradient(X)
        return yh.reshape(-1, 1), ms.reshape(-1, 1), ydx, mdx


def _is_basis_trend(mean) -> bool:
    return isinstance(mean, BasisExpansionTrend) or any(
        c.__name__ == "BasisExpansionTrend" for c in type(mean).__mro__
    )
