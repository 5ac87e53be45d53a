"""TEST INFRASTRUCTURE ONLY -- float64 numpy/scipy restatement of the reference hot path.

Reference: wangronin/Bayesian-Optimization (``bayes-optim`` 0.3.0).  All ``file:line`` citations are
relative to ``/root/reference/bayes_optim/`` and name the statement each function follows.

PARITY PINNING: the reference's own tests contain no numeric vectors for this path (SURVEY.md §4,
§8c), so this restatement is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks it against them (and, when ``/root/reference`` is present,
against the live reference).

This module is the *checker* for the CUDA path and the timed CPU baseline of ``bench.py``.  It is
never imported by ``bayesian_optimization_b200`` (the product), which has no CPU fallback.

Notation: N training points, D features, M candidates, p trend-basis size, k targets (k = 1).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
from scipy.linalg import cho_solve, cholesky, qr, solve_triangular
from scipy.special import ndtr

# correlation ids shared with include/b200bo.h (B200BO_CORR_*)
CORR_RBF = 0  # "squared_exponential"   surrogate/gaussian_process/kernel.py:289-329
CORR_MATERN12 = 1  # matern(nu=0.5)       kernel.py:189-190
CORR_MATERN32 = 2  # matern(nu=1.5) = the string "matern"   kernel.py:192-195, gpr.py:206
CORR_MATERN52 = 3  # matern(nu=2.5)       kernel.py:197-200
CORR_ABSEXP = 4  # "absolute_exponential" kernel.py:247-286
CORR_CUBIC = 5  # "cubic"                  kernel.py:419-466
CORR_GENEXP = 6  # "generalized_exponential" kernel.py:332-374 (theta = [theta_1..n | theta, p])
CORR_MATERN_NU = 7  # matern(nu) for any other nu: kernel.py:201-207 via scipy.special.kv (theta = [theta_1..n | theta, nu])

CORR_NAMES = {
    "squared_exponential": CORR_RBF,
    "matern": CORR_MATERN32,
    "matern12": CORR_MATERN12,
    "matern32": CORR_MATERN32,
    "matern52": CORR_MATERN52,
    "absolute_exponential": CORR_ABSEXP,
    "cubic": CORR_CUBIC,
}

MODE_NOISELESS, MODE_NOISY, MODE_NOISE_ESTIM = 0, 1, 2
TREND_CONSTANT, TREND_LINEAR, TREND_QUADRATIC = 0, 1, 2


# --------------------------------------------------------------------------------------------------
# distances and correlation functions
# --------------------------------------------------------------------------------------------------
def cross_abs_diff(X: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """|x_i - y_j| component-wise, flattened to (len(X)*len(Y), D).  gpr.py:42-47 (predict branch of
    ``l1_cross_distances``): broadcast subtract, in-place abs, reshape."""
    diff = X[:, np.newaxis, :] - Y[np.newaxis, :, :]
    np.abs(diff, out=diff)
    return diff.reshape(-1, X.shape[1])


def pair_abs_diff(X: np.ndarray):
    """Strict upper-triangle component-wise distances and their (i, j) index pairs, row-major over i<j.
    gpr.py:48-61 (fit branch).  Vectorised with triu_indices -- same ordering as the reference's
    ``for k in range(n-1)`` loop."""
    iu, ju = np.triu_indices(X.shape[0], 1)
    return np.abs(X[iu] - X[ju]), np.c_[iu, ju]


def corr_values(corr: int, theta: np.ndarray, d: np.ndarray) -> np.ndarray:
    """Stationary correlation r(theta, d) for component-wise distances d (P, D) -> (P,).

    theta multiplies the *squared* distance (no 1/2, not a length-scale): kernel.py:326-329 (RBF);
    Matern uses h = sqrt(sum theta_j d_j^2): kernel.py:184-187, then nu=.5 :189-190, nu=1.5 :192-195,
    nu=2.5 :197-200.  theta of size 1 is isotropic.  absolute_exponential kernel.py:280-286,
    cubic kernel.py:455-466."""
    theta = np.asarray(theta, dtype=np.float64).ravel()
    d = np.asarray(d, dtype=np.float64)
    nf = d.shape[1]
    if corr == CORR_GENEXP:  # kernel.py:362-374: theta = [theta | theta_1..n, p]
        if theta.size not in (2, nf + 1):
            raise ValueError("Length of theta must be 2 or %s" % (nf + 1))
        th = np.repeat(theta[0], nf) if theta.size == 2 and nf > 1 else theta[:-1]
        return np.exp(-np.sum(th.reshape(1, nf) * np.abs(d) ** theta[-1], axis=1))
    if corr == CORR_MATERN_NU:  # nu rides as the last entry of theta (an argument of the callable upstream)
        from scipy.special import gamma as _gamma, kv as _kv

        if theta.size not in (2, nf + 1):
            raise ValueError("Length of theta must be 2 or %s" % (nf + 1))
        nu, th = float(theta[-1]), theta[:-1]
        h = np.sqrt(th[0] * np.sum(d**2, axis=1)) if th.size == 1 else np.sqrt(np.sum(th.reshape(1, nf) * d**2, axis=1))
        K = h.copy()
        K[K == 0.0] += np.finfo(float).eps  # kernel.py:203
        tmp = math.sqrt(2 * nu) * K
        return (2 ** (1.0 - nu)) / _gamma(nu) * tmp**nu * _kv(nu, tmp)
    if theta.size not in (1, nf):
        raise ValueError("Length of theta must be 1 or %s" % nf)
    if corr == CORR_RBF:
        if theta.size == 1:
            return np.exp(-theta[0] * np.sum(d**2, axis=1))
        return np.exp(-np.sum(theta.reshape(1, nf) * d**2, axis=1))
    if corr in (CORR_MATERN12, CORR_MATERN32, CORR_MATERN52):
        if theta.size == 1:
            h = np.sqrt(theta[0] * np.sum(d**2, axis=1))
        else:
            h = np.sqrt(np.sum(theta.reshape(1, nf) * d**2, axis=1))
        if corr == CORR_MATERN12:
            return np.exp(-h)
        if corr == CORR_MATERN32:
            k = h * math.sqrt(3)
            return (1.0 + k) * np.exp(-k)
        k = h * math.sqrt(5)
        return (1.0 + k + k**2 / 3.0) * np.exp(-k)
    if corr == CORR_ABSEXP:
        d = np.abs(d)
        if theta.size == 1:
            return np.exp(-theta[0] * np.sum(d, axis=1))
        return np.exp(-np.sum(theta.reshape(1, nf) * d, axis=1))
    if corr == CORR_CUBIC:
        td = np.abs(d) * (theta if theta.size == 1 else theta.reshape(1, nf))
        td = np.minimum(td, 1.0)
        return np.prod(1.0 - td**2.0 * (3.0 - 2.0 * td), axis=1)
    raise ValueError(f"unknown correlation id {corr}")


def trend_basis(trend: int, X: np.ndarray) -> np.ndarray:
    """F(X) (M, p).  constant: ones  trend.py:76-79; linear: [1, x]  trend.py:104-107;
    quadratic: [1, x, {x_k * x_j, j >= k}]  trend.py:129-135."""
    M, D = X.shape
    if trend == TREND_CONSTANT:
        return np.ones((M, 1))
    if trend == TREND_LINEAR:
        return np.c_[np.ones(M), X]
    if trend == TREND_QUADRATIC:
        cols = [np.ones((M, 1)), X]
        for k in range(D):
            cols.append(X[:, k, np.newaxis] * X[:, k:])
        return np.concatenate(cols, axis=1)
    raise ValueError(f"unknown trend id {trend}")


# --------------------------------------------------------------------------------------------------
# fitted state
# --------------------------------------------------------------------------------------------------
@dataclass
class OracleGP:
    """Fixed-hyper-parameter GP state, the quantities ``GaussianProcess.fit`` stores (gpr.py:402-415)."""

    X: np.ndarray
    y: np.ndarray  # (N, 1)
    corr: int
    theta: np.ndarray
    mode: int
    trend: int = TREND_CONSTANT
    beta_fixed: Optional[np.ndarray] = None  # None => ordinary/universal kriging (beta estimated)
    # filled by fit_fixed
    sigma2: float = float("nan")
    noise_var: float = 0.0
    llf: float = -float("inf")
    L: np.ndarray = field(default=None, repr=False)  # "C" in the reference
    Yt: np.ndarray = field(default=None, repr=False)
    Ft: np.ndarray = field(default=None, repr=False)
    Q: np.ndarray = field(default=None, repr=False)
    G: np.ndarray = field(default=None, repr=False)
    rho: np.ndarray = field(default=None, repr=False)
    beta: np.ndarray = field(default=None, repr=False)
    gamma: np.ndarray = field(default=None, repr=False)
    R0: np.ndarray = field(default=None, repr=False)

    @property
    def estimate_trend(self) -> bool:
        return self.beta_fixed is None


def correlation_matrix(corr: int, theta: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Dense symmetric R0 with unit diagonal.  gpr.py:772-782 (scatter of corr(theta, D) into eye(N))."""
    N = X.shape[0]
    D, ij = pair_abs_diff(X)
    r = corr_values(corr, theta, D)
    R = np.eye(N)
    R[ij[:, 0], ij[:, 1]] = r
    R[ij[:, 1], ij[:, 0]] = r
    return R


def _aux(R: np.ndarray, gp: OracleGP):
    """gpr.py:790-811 (_compute_aux_var): L = chol(R); Yt = L^-1 y;
    OK: Ft = L^-1 F, thin QR, rho = Yt - Q Q^T Yt;   SK: rho = Yt - L^-1 (F beta)."""
    L = cholesky(R, lower=True)
    Yt = solve_triangular(L, gp.y, lower=True)
    F = trend_basis(gp.trend, gp.X)
    if gp.estimate_trend:
        Ft = solve_triangular(L, F, lower=True)
        Q, G = qr(Ft, mode="economic")
        rho = Yt - Q.dot(Q.T).dot(Yt)
    else:
        Ft, Q, G = None, None, None
        rho = Yt - solve_triangular(L, F.dot(gp.beta_fixed.reshape(-1, 1)), lower=True)
    return L, Ft, Yt, Q, G, rho


def fit_fixed(
    X,
    y,
    corr: int,
    theta,
    mode: int,
    sigma2: Optional[float] = None,
    noise_var: float = 0.0,
    alpha: Optional[float] = None,
    trend: int = TREND_CONSTANT,
    beta_fixed=None,
) -> OracleGP:
    """Concentrated log-likelihood at fixed hyper-parameters + the state fit() keeps.

    gpr.py:920-991 (log_likelihood_concentrated, three estimation modes), :981-982 (any llf > 0 is
    rejected as -inf), :784-788 (compute_beta_gamma).  ``sigma2`` is an input only in ``noisy`` mode;
    ``alpha`` only in ``noise_estim`` mode."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1)
    theta = np.asarray(theta, dtype=np.float64).ravel()
    bf = None if beta_fixed is None else np.asarray(beta_fixed, dtype=np.float64).ravel()
    gp = OracleGP(X=X, y=y, corr=corr, theta=theta, mode=mode, trend=trend, beta_fixed=bf)
    N = X.shape[0]
    R0 = correlation_matrix(corr, theta, X)
    gp.R0 = R0
    try:
        if mode == MODE_NOISELESS:  # gpr.py:932-947
            L, Ft, Yt, Q, G, rho = _aux(R0, gp)
            k = np.linalg.matrix_rank(Q.dot(Q.T)) if Q is not None else 0
            s2 = (rho**2.0).sum(axis=0) / (N - k)
            nv = 0.0
            llf = -0.5 * (N * np.log(2.0 * np.pi * s2) + 2.0 * np.log(np.diag(L)).sum() + N)
        elif mode == MODE_NOISE_ESTIM:  # gpr.py:949-961
            R = alpha * R0 + (1 - alpha) * np.eye(N)
            L, Ft, Yt, Q, G, rho = _aux(R, gp)
            s2t = (rho**2.0).sum(axis=0) / N
            s2, nv = alpha * s2t, (1 - alpha) * s2t
            llf = -0.5 * (N * np.log(2.0 * np.pi * s2t) + 2.0 * np.log(np.diag(L)).sum() + N)
        elif mode == MODE_NOISY:  # gpr.py:963-979
            nv = float(noise_var)
            s2t = sigma2 + nv
            C = sigma2 * R0 + nv * np.eye(N)
            R = C / s2t
            s2 = np.repeat(float(sigma2), y.shape[1])
            L, Ft, Yt, Q, G, rho = _aux(R, gp)
            llf = -0.5 * (
                N * np.log(2.0 * np.pi * s2t) + 2.0 * np.log(np.diag(L)).sum() + np.diag(rho.T.dot(rho)) / s2t
            )
        else:
            raise ValueError(mode)
    except (np.linalg.LinAlgError, ValueError):
        llf = None
    if llf is None or np.any(np.asarray(llf) > 0) or not np.all(np.isfinite(llf)):
        gp.llf = -np.inf  # gpr.py:981-982
        return gp
    gp.llf = float(np.sum(llf))
    gp.sigma2 = float(np.atleast_1d(s2)[0])
    gp.noise_var = float(np.atleast_1d(nv)[0])
    gp.L, gp.Ft, gp.Yt, gp.Q, gp.G, gp.rho = L, Ft, Yt, Q, G, rho
    # gpr.py:784-788
    if gp.estimate_trend:
        gp.beta = solve_triangular(G, Q.T.dot(Yt))
    else:
        gp.beta = bf.reshape(-1, 1)
    gp.gamma = solve_triangular(L.T, rho).reshape(-1, y.shape[1])
    return gp


def fit_fixed_restricted(X, y, corr: int, theta, sigma2: float, noise_var: float = 0.0, trend: int = TREND_CONSTANT,
                         beta_fixed=None, eval_grad: bool = False, n_par: Optional[int] = None):
    """Restricted (REML) log-likelihood at fixed hyper-parameters + the state fit() keeps.  gpr.py:813-918.
    All three estimation modes build R = (sigma2 R0 + noise_var I) / (sigma2 + noise_var) (:826-839; noise_var = 0,
    the nugget, or the last parameter), so the mode only decides n_par.  Note the SIGN of the log-determinant term
    in the simple-kriging branch (:866), kept as is.  With eval_grad: (gp, gradient (n_par,)) per :876-902, whose
    tensor slices are indexed by the parameter number (the isotropic-theta quirk of SURVEY App. A g3)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1)
    theta = np.asarray(theta, dtype=np.float64).ravel()
    bf = None if beta_fixed is None else np.asarray(beta_fixed, dtype=np.float64).ravel()
    gp = OracleGP(X=X, y=y, corr=corr, theta=theta, mode=MODE_NOISY, trend=trend, beta_fixed=bf)
    N = X.shape[0]
    R0 = correlation_matrix(corr, theta, X)
    gp.R0 = R0
    tv = sigma2 + noise_var
    R = (sigma2 * R0 + noise_var * np.eye(N)) / tv
    try:
        L, Ft, Yt, Q, G, rho = _aux(R, gp)
    except np.linalg.LinAlgError:
        return (gp, np.zeros(n_par or theta.size + 1)) if eval_grad else gp
    F = trend_basis(trend, X)
    if gp.estimate_trend:
        p = Ft.shape[1]
        llf = -0.5 * ((N - p) * np.log(2 * np.pi * tv) - np.log(np.linalg.det(F.T.dot(F))) + 2 * np.log(np.diag(L)).sum()
                      + np.log(np.diag(G).prod() ** 2) + rho.T.dot(rho) / tv).sum()
    else:
        llf = -0.5 * (N * np.log(2 * np.pi * tv) - 2 * np.log(np.diag(L)).sum() + rho.T.dot(rho) / tv).sum()
    with np.errstate(over="ignore"):
        gp.llf = float(llf) if not np.exp(llf) > 1 else -np.inf  # :872-875
    gp.sigma2, gp.noise_var = float(sigma2), float(noise_var)
    gp.L, gp.Ft, gp.Yt, gp.Q, gp.G, gp.rho = L, Ft, Yt, Q, G, rho
    gp.beta = solve_triangular(G, Q.T.dot(Yt)) if gp.estimate_trend else bf.reshape(-1, 1)
    gp.gamma = solve_triangular(L.T, rho).reshape(-1, y.shape[1])
    if not eval_grad:
        return gp
    n_par = n_par or theta.size + 1
    gamma_ = gp.gamma / tv
    Cinv = cho_solve((L, True), np.eye(N)) / tv
    if gp.estimate_trend:
        t_ = solve_triangular(L.T, Q)
        term = t_.dot(t_.T)
    D = X.shape[1]
    slices = [tv * corr_dtheta(gp, j) for j in range(D)] + [R0, np.eye(N)]
    grad = np.zeros(n_par)
    for i in range(n_par):
        Cg = slices[i]
        g = np.sum(Cinv * Cg) - float(gamma_.T.dot(Cg).dot(gamma_).sum())
        if gp.estimate_trend:
            g -= np.sum(term * Cg)
        grad[i] = -0.5 * g
    return gp, grad


def corr_dtheta(gp: OracleGP, j: int) -> np.ndarray:
    """dR0/dtheta_j as the reference's ``corr_grad_theta`` defines it (one (N,N) slice of its (N,N,D)
    tensor).  gpr.py:745 (squared differences), :748 (RBF: -diff*R), :753-757 (Matern-3/2 only:
    -3 exp(-sqrt3 h) diff / 2), :761 (absolute_exponential: -|diff| R)."""
    X = gp.X
    diff2 = (X[:, np.newaxis, j] - X[np.newaxis, :, j]) ** 2.0
    if gp.corr == CORR_RBF:
        return -diff2 * gp.R0
    if gp.corr == CORR_MATERN32:
        full = (X[:, np.newaxis, :] - X[np.newaxis, :, :]) ** 2.0
        h = np.sqrt(np.sum(gp.theta * full, axis=-1))
        return -3 * np.exp(-math.sqrt(3) * h) * diff2 / 2.0
    if gp.corr == CORR_ABSEXP:
        return -np.sqrt(diff2) * gp.R0
    raise NotImplementedError("reference leaves this kernel's theta-gradient unimplemented (gpr.py:758-768)")


def llf_grad(gp: OracleGP, alpha: Optional[float] = None) -> np.ndarray:
    """Analytic gradient of the concentrated log-likelihood exactly as the reference computes it
    (NOT the true derivative: quirks g2/g3 of SURVEY.md App. A are reproduced).  gpr.py:994-1038.
    Uses only the strict upper triangle for the theta components (triu_indices(N, 1))."""
    N = gp.X.shape[0]
    L, rho = gp.L, gp.rho
    gamma = solve_triangular(L.T, rho).reshape(-1, 1)
    Rinv = cho_solve((L, True), np.eye(N))
    iu = np.triu_indices(N, 1)
    Rinv_u = Rinv[iu]
    gg_u = gamma.dot(gamma.T)[iu]
    nt = gp.theta.size
    if gp.mode == MODE_NOISELESS:  # gpr.py:1002-1010
        g = np.zeros(nt)
        for i in range(nt):
            Ru = corr_dtheta(gp, i)[iu]
            g[i] = np.sum(gg_u * Ru) / gp.sigma2 - np.sum(Rinv_u * Ru)
        return g
    if gp.mode == MODE_NOISE_ESTIM:  # gpr.py:1011-1025
        s2t = gp.sigma2 + gp.noise_var
        g = np.zeros(nt + 1)
        for i in range(nt):
            Ru = alpha * corr_dtheta(gp, i)[iu]
            g[i] = np.sum(gg_u * Ru) / s2t - np.sum(Rinv_u * Ru)
        R_dv = gp.R0 - np.eye(N)
        g[nt] = -0.5 * (np.sum(Rinv * R_dv) - float(gamma.T.dot(R_dv.dot(gamma))[0, 0]) / s2t)
        return g
    # noisy: gpr.py:1026-1038
    s2t = gp.sigma2 + gp.noise_var
    gamma_ = gamma / s2t
    Cinv = Rinv / s2t
    g = np.zeros(nt + 1)
    for i in range(nt + 1):
        Cg = s2t * corr_dtheta(gp, i) if i < nt else gp.R0
        g[i] = -0.5 * (np.sum(Cinv * Cg) - float(gamma_.T.dot(Cg).dot(gamma_)[0, 0]))
    return g


# --------------------------------------------------------------------------------------------------
# predict
# --------------------------------------------------------------------------------------------------
def predict(gp: OracleGP, Xc: np.ndarray, eval_MSE: bool = True):
    """BLUP mean and MSE for one batch (caller chunks).  gpr.py:486-510:
    r = corr(theta, |Xc - X|) -- UNSCALED even in noisy mode (:486-488); yhat = F(Xc) beta + r gamma (:490);
    rt = L^-1 r^T (:494); OK: u = G^-T (Ft^T rt - F(Xc)^T) (:496-498); MSE = sigma2 (1 - sum rt^2 + sum u^2)
    (:502-505), clipped at 0 (:510)."""
    Xc = np.ascontiguousarray(Xc, dtype=np.float64)
    M = Xc.shape[0]
    N = gp.X.shape[0]
    dx = cross_abs_diff(Xc, gp.X)
    r = corr_values(gp.corr, gp.theta, dx).reshape(M, N)
    f = trend_basis(gp.trend, Xc)
    yhat = (f.dot(gp.beta) + r.dot(gp.gamma)).reshape(M, 1)
    if not eval_MSE:
        return yhat
    rt = solve_triangular(gp.L, r.T, lower=True)
    if gp.estimate_trend:
        u = solve_triangular(gp.G.T, np.dot(gp.Ft.T, rt) - f.T, lower=True)
    else:
        u = np.zeros((1, M))
    mse = (1.0 - (rt**2.0).sum(axis=0) + (u**2.0).sum(axis=0)).reshape(M, 1) * gp.sigma2
    mse[mse < 0.0] = 0.0
    return yhat, mse


def predict_chunked(gp: OracleGP, Xc: np.ndarray, chunk: int, eval_MSE: bool = True):
    """External chunking (the reference's own batch_size branch is dead on Python 3: gpr.py:520)."""
    ys, ms = [], []
    for a in range(0, Xc.shape[0], chunk):
        out = predict(gp, Xc[a : a + chunk], eval_MSE)
        if eval_MSE:
            ys.append(out[0])
            ms.append(out[1])
        else:
            ys.append(out)
    if eval_MSE:
        return np.concatenate(ys), np.concatenate(ms)
    return np.concatenate(ys)


def posterior_gradient(gp: OracleGP, x: np.ndarray):
    """d yhat / dx and d MSE / dx at one point (D,1),(D,1).  gpr.py:537-576 with corr_dx :600-661
    (RBF :634, Matern-3/2 :645-647, absolute_exponential :652)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    N, D = gp.X.shape
    f = trend_basis(gp.trend, x).reshape(-1, 1)
    if gp.trend == TREND_CONSTANT:
        f_dx = np.zeros((1, D))
    elif gp.trend == TREND_LINEAR:
        f_dx = np.r_[np.zeros((1, D)), np.eye(D)]
    else:
        raise NotImplementedError
    d = cross_abs_diff(x, gp.X)
    r = corr_values(gp.corr, gp.theta, d).reshape(1, N)
    diff = (x - gp.X).T  # (D, N)
    th = gp.theta.reshape(-1, 1)
    if gp.corr == CORR_RBF:
        r_dx = -2 * r * (th * diff)
    elif gp.corr == CORR_MATERN32:
        h = np.sqrt(np.sum(th * diff**2.0, axis=0))
        if np.any(h == 0):
            # gpr.py:628-630,660-661: the 0/0 RuntimeWarning is raised as an error inside corr_dx and the
            # WHOLE Jacobian is replaced by zeros (quirk: happens when x coincides with a training point)
            r_dx = np.zeros((D, N))
        else:
            r_dx = diff * th / h
            r_dx = r_dx * (-3.0 * h * np.exp(-math.sqrt(3) * h))
    elif gp.corr == CORR_ABSEXP:
        r_dx = -1.0 * r * th * np.sign(diff)
    else:
        raise NotImplementedError
    r_dx = r_dx.T  # (N, D)
    y_dx = gp.beta.T.dot(f_dx) + gp.gamma.T.dot(r_dx)
    rt = solve_triangular(gp.L, r.T, lower=True)
    rt_dx = solve_triangular(gp.L, r_dx, lower=True)
    mse_dx = -1.0 * rt.T.dot(rt_dx)
    if gp.estimate_trend:
        u = gp.Ft.T.dot(rt) - f
        u_dx = gp.Ft.T.dot(rt_dx) - f_dx
        Ft2inv = np.linalg.inv(gp.Ft.T.dot(gp.Ft))
        mse_dx = mse_dx + u.T.dot(Ft2inv).dot(u_dx)
    mse_dx = 2.0 * gp.sigma2 * mse_dx
    return y_dx.T, mse_dx.T


# --------------------------------------------------------------------------------------------------
# acquisition functions, batched with the reference's PER-ROW semantics
# --------------------------------------------------------------------------------------------------
ACQ_EI, ACQ_PI, ACQ_UCB, ACQ_MGFI = 0, 1, 2, 3
_SQRT_2PI = math.sqrt(2.0 * math.pi)  # scipy.stats.norm.pdf divides by this constant


def _mean_sd(yhat, mse, minimize: bool):
    """acquisition/acquisition_fun.py:52-64: negate yhat when maximising; sd = sqrt(MSE)."""
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    sd = np.sqrt(np.asarray(mse, dtype=np.float64).ravel())
    return (yhat if minimize else -yhat), sd


def plugin_value(y_train, minimize: bool, plugin=None):
    """acquisition_fun.py:96-104: default plug-in min(y) (or -max(y)); user value negated when maximising."""
    if plugin is None:
        return float(np.min(y_train)) if minimize else -float(np.max(y_train))
    return float(plugin) if minimize else -float(plugin)


def ei(yhat, mse, sigma2: float, plugin: float, minimize: bool = True) -> np.ndarray:
    """Expected improvement.  acquisition_fun.py:162-164 (0 when sd/sqrt(sigma2) < 1e-6),
    :170-174 (value = (f*-yhat) Phi(z) + sd phi(z))."""
    y, sd = _mean_sd(yhat, mse, minimize)
    out = np.zeros_like(y)
    ok = ~(sd / np.sqrt(sigma2) < 1e-6)
    d = plugin - y[ok]
    z = d / sd[ok]
    out[ok] = d * ndtr(z) + sd[ok] * (np.exp(-(z**2) / 2.0) / _SQRT_2PI)
    return out


def pi_eps(yhat, mse, plugin: float, epsilon: float = 0.0, minimize: bool = True) -> np.ndarray:
    """epsilon-PI; epsilon = 0 is "PI" (unconstructible upstream).
    acquisition_fun.py:212-216: coef = 1-eps if yhat > 0 else 1+eps; Phi((f* - coef yhat)/sd)."""
    y, sd = _mean_sd(yhat, mse, minimize)
    coef = np.where(y > 0, 1.0 - epsilon, 1.0 + epsilon)
    with np.errstate(divide="ignore", invalid="ignore"):
        return ndtr((plugin - coef * y) / sd)


def ucb(yhat, mse, alpha: float, minimize: bool = True) -> np.ndarray:
    """acquisition_fun.py:133: yhat + alpha * sd (yhat already negated for maximisation)."""
    y, sd = _mean_sd(yhat, mse, minimize)
    return y + alpha * sd


def mgfi(yhat, mse, plugin: float, t: float, minimize: bool = True) -> np.ndarray:
    """Moment-generating function of the improvement.  acquisition_fun.py:262 (t capped at 22.36),
    :274 (0 when isclose(sd, 0)), :280-283, :284-290 (overflow / inf -> 0)."""
    t = min(float(t), 22.36)
    y, sd = _mean_sd(yhat, mse, minimize)
    out = np.zeros_like(y)
    ok = ~np.isclose(sd, 0)
    yo, so = y[ok], sd[ok]
    y_p = yo - t * so**2.0
    beta_p = (plugin - y_p) / so
    term = t * (plugin - yo - 1)
    e = term + t**2.0 * so**2.0 / 2.0
    with np.errstate(over="ignore", invalid="ignore"):
        v = ndtr(beta_p) * np.exp(e)
    # the reference turns the exp-overflow RuntimeWarning into an exception -> 0 (:277-287)
    v[(e > np.log(np.finfo(np.float64).max)) | ~np.isfinite(v)] = 0.0
    out[ok] = v
    return out


def acquisition(acq: int, yhat, mse, sigma2, plugin, par, minimize=True) -> np.ndarray:
    if acq == ACQ_EI:
        return ei(yhat, mse, sigma2, plugin, minimize)
    if acq == ACQ_PI:
        return pi_eps(yhat, mse, plugin, par, minimize)
    if acq == ACQ_UCB:
        return ucb(yhat, mse, par, minimize)
    if acq == ACQ_MGFI:
        return mgfi(yhat, mse, plugin, par, minimize)
    raise ValueError(acq)


def acquisition_dx(acq: int, yhat, mse, y_dx, mse_dx, sigma2, plugin, par, minimize=True):
    """Value and gradient of one acquisition function at ONE point (the reference's return_dx=True path).
    yhat, mse: scalars from predict; y_dx, mse_dx: (D,) from posterior_gradient.  Returns (value, dx (D,)).
    acquisition_fun.py:66-80 (_gradient: y_dx negated when maximising), UCB :139-146, EI :162-168 (early-out
    zeros), :181-188, EpsilonPI :220-229, MGFI :274-275 (early-out), :292-309 (failures -> zeros)."""
    y = np.float64(yhat) if minimize else -np.float64(yhat)  # numpy scalars: x / 0 is inf / nan as in the reference
    sd = np.sqrt(np.float64(mse))
    sigma2, plugin, par = np.float64(sigma2), np.float64(plugin), np.float64(par)
    ydx = np.asarray(y_dx, dtype=np.float64).ravel() * (1.0 if minimize else -1.0)
    mdx = np.asarray(mse_dx, dtype=np.float64).ravel()
    D = ydx.size
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if acq == ACQ_UCB:
            return y + par * sd, ydx + par * (mdx / (2.0 * sd))
        if acq == ACQ_EI:
            if sd / np.sqrt(sigma2) < 1e-6:
                return 0.0, np.zeros(D)
            d = plugin - y
            z = d / sd
            cdf, pdf = ndtr(z), np.exp(-(z**2) / 2.0) / _SQRT_2PI
            return d * cdf + sd * pdf, -ydx * cdf + (mdx / (2.0 * sd)) * pdf
        if acq == ACQ_PI:
            coef = 1.0 - par if y > 0 else 1.0 + par
            z = (plugin - coef * y) / sd
            pdf = np.exp(-(z**2) / 2.0) / _SQRT_2PI
            return float(ndtr(z)), -(coef * ydx + z * (mdx / (2.0 * sd))) * pdf / sd
        if acq == ACQ_MGFI:
            t = np.float64(min(float(par), 22.36))
            if np.isclose(sd, 0):
                return 0.0, np.zeros(D)
            beta_p = (plugin - (y - t * sd**2.0)) / sd
            e = t * (plugin - y - 1) + t**2.0 * sd**2.0 / 2.0
            big = e > math.log(np.finfo(np.float64).max)
            val = 0.0 if big else float(ndtr(beta_p) * np.exp(e))
            if not np.isfinite(val):
                val = 0.0
            if big:  # exp overflow inside the gradient block raises -> zeros (:305-306)
                return val, np.zeros(D)
            sd_dx = mdx / (2.0 * sd)
            term = np.exp(t * (plugin + t * sd**2.0 / 2 - y - 1))
            m_prime_dx = ydx - 2.0 * t * sd * sd_dx
            beta_p_dx = -(m_prime_dx + beta_p * sd_dx) / sd
            pdf = np.exp(-(beta_p**2) / 2.0) / _SQRT_2PI
            return val, term * (pdf * beta_p_dx + float(ndtr(beta_p)) * ((t**2) * sd * sd_dx - t * ydx))
    raise ValueError(acq)


def argmax_first(v: np.ndarray) -> int:
    """numpy argmax semantics: first (lowest-index) maximum."""
    return int(np.argmax(v))


# --------------------------------------------------------------------------------------------------
# acquisition-parameter recipes (host scalars)
# --------------------------------------------------------------------------------------------------
def mgfi_t_samples(t: float, q: int, seed: int = 42) -> np.ndarray:
    """ParallelBO's log-normal sampler t_i = exp(log t + 0.5 xi), xi ~ N(0,1) from the GLOBAL numpy RNG.
    bayes_opt.py:82-85."""
    rs = np.random.RandomState(seed)
    return np.array([np.exp(np.log(t) + 0.5 * rs.randn()) for _ in range(q)])


def ucb_alpha_samples(alpha: float, q: int, seed: int = 42) -> np.ndarray:
    """ParallelBO's logit-normal sampler alpha_i = 1 / (1 + exp(4 alpha - 2 + 0.6 xi)).  bayes_opt.py:86-89."""
    rs = np.random.RandomState(seed)
    return np.array([1 / (1 + np.exp((alpha * 4 - 2) + 0.6 * rs.randn())) for _ in range(q)])


def annealing_t_schedule(t0: float, tf: float, steps: int) -> np.ndarray:
    """AnnealingBO exponential schedule t <- t * (tf/t0)^(1/max_iter).  bayes_opt.py:127-130."""
    a = (tf / t0) ** (1.0 / steps)
    return t0 * a ** np.arange(steps)


# --------------------------------------------------------------------------------------------------
# canonical synthetic inputs (SURVEY.md §8d / BASELINE.md §3)
# --------------------------------------------------------------------------------------------------
def canonical_problem(N: int, D: int):
    X = np.random.default_rng(42).uniform(0, 1, (N, D))
    y_raw = np.sin(2 * np.pi * X).sum(axis=1) / np.sqrt(D) + 0.1 * np.random.default_rng(43).standard_normal(N)
    y = (y_raw - y_raw.mean()) / y_raw.std()
    theta = 10.0 ** np.linspace(-0.3, 0.7, D) * (8.0 / D)
    return X, y, theta


def canonical_candidates(M: int, D: int, shard: int = 0):
    return np.random.default_rng(1000 + shard).uniform(0, 1, (M, D))
