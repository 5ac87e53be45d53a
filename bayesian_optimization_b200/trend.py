"""Host-side trend (prior mean) objects with the interface of the reference's
``bayes_optim/surrogate/gaussian_process/trend.py``: ``constant_trend`` (:69-91) is what ``fmin`` and every upstream GP
test use (bayes_optim/__init__.py:148, unittest/test_BO.py:34), ``linear_trend`` (:94-116) and ``quadratic_trend``
(:119-142) are the other two regression bases.  ``beta=None`` means the coefficients are estimated (ordinary / universal
kriging), a value means simple kriging (gpr.py:269-275).

Design: ONE polynomial-basis class parametrised by its degree; the three public names are thin subclasses, so that
``type(mean).__name__`` -- which is how ``GaussianProcess`` picks the device trend id -- reads as upstream.  The device
evaluates the bases itself (constant: the literal 1 inside the kernels; degree 1 / 2: csrc/trend_kernels.cuh); these
objects carry the coefficients and answer the host-side questions (``F``, ``Jacobian``, ``__call__``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["BasisExpansionTrend", "constant_trend", "linear_trend", "quadratic_trend"]


def _n_basis(n_feature: int, degree: int) -> int:
    n = int(n_feature)
    return (1, n + 1, (n + 1) * (n + 2) // 2)[degree]


class BasisExpansionTrend:
    """``m(X) = F(X) beta`` with ``beta`` kept as a (p, 1) column (trend.py:10-64), or (p, k) after a k-target fit."""

    def __init__(self, n_feature: int, n_dim: int, beta=None):
        self.n_feature, self.n_dim = int(n_feature), int(n_dim)
        self.beta = beta

    # -- coefficients --------------------------------------------------------------------------------------
    @property
    def beta(self):
        return self._beta

    @beta.setter
    def beta(self, value):
        if value is None:
            self._beta = None
            return
        if np.ndim(value) == 0:  # a scalar stands for every coefficient (trend.py:24-25)
            value = np.full(self.n_dim, value, dtype=np.float64)
        arr = np.atleast_2d(np.asarray(value, dtype=np.float64))
        if arr.shape[0] == self.n_dim and arr.shape[1] > 1 and (self.n_dim > 1 or np.ndim(value) == 2):
            # (p, k): one column per target of a multi-target fit with beta estimated.  Upstream flattens this and
            # raises (trend.py:25-28), which is why its ordinary kriging cannot take y (N, k > 1).
            self._beta = arr
            return
        col = arr.reshape(-1, 1)
        if col.shape[0] != self.n_dim:
            raise Exception("Shapes of beta and F do not match.")
        self._beta = col

    # -- evaluation ----------------------------------------------------------------------------------------
    def check_input(self, X):
        """2-D view of X with the features along axis 1; a (D, M) input is turned round silently (trend.py:51-58)."""
        X = np.atleast_2d(X)
        if X.shape[1] != self.n_feature:
            X = X.T
            if X.shape[1] != self.n_feature:
                raise Exception("X does not have the right size!")
        return X

    def F(self, X):  # pragma: no cover - provided by the polynomial subclass
        raise NotImplementedError

    def __call__(self, X):
        if self._beta is None:
            raise Exception("beta is not set!")
        return self.F(X) @ self._beta


class _PolynomialTrend(BasisExpansionTrend):
    """Full polynomial basis of total degree ``_degree`` in the order the reference lists it:
    1 | x_1 .. x_n | x_k x_j for k = 1..n, j = k..n."""

    _degree = 0

    def __init__(self, n_feature: int, beta=None):
        super().__init__(n_feature, _n_basis(n_feature, self._degree), beta)

    def F(self, X):
        X = self.check_input(X)
        blocks = [np.ones((X.shape[0], 1))]
        if self._degree >= 1:
            blocks.append(X)
        if self._degree >= 2:
            blocks.extend(X[:, k:] * X[:, k:k + 1] for k in range(self.n_feature))
        return np.hstack(blocks)

    def Jacobian(self, x):
        """d f / d x at ONE point, (p, n) (numerator layout as upstream); not defined upstream for degree 2"""
        x = self.check_input(x)
        if self._degree == 2:
            raise NotImplementedError  # trend.py:138-139
        if self._degree == 1:
            assert x.shape[0] == 1
            return np.vstack([np.zeros((1, self.n_feature)), np.eye(self.n_feature)])
        return np.zeros((1, self.n_feature))

    def Hessian(self, x):
        self.check_input(x)
        if self._degree == 2:
            raise NotImplementedError  # trend.py:141-142
        return np.zeros((self.n_feature, self.n_feature, self.n_dim))


class constant_trend(_PolynomialTrend):
    """zero-order polynomial, p = 1, f(x) = 1 (trend.py:69-91)"""

    _degree = 0


class linear_trend(_PolynomialTrend):
    """first-order polynomial, p = n + 1, f(x) = [1, x_1, ..., x_n] (trend.py:94-116)"""

    _degree = 1


class quadratic_trend(_PolynomialTrend):
    """second-order polynomial, p = (n + 1)(n + 2) / 2, f(x) = [1, {x_i}, {x_k x_j, j >= k}] (trend.py:119-142)"""

    _degree = 2
