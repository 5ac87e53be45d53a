"""GPU: the Cholesky trailing update on the tcgen05 tensor cores (csrc/oz_kernels.cuh, b200bo_set_chol_tc).

The rank-64 update C -= P P^T of the blocked factorisation (scipy.linalg.cholesky, gpr.py:795) runs as exact int8 digit
products; what is checked: (1) the kernel against an extended-precision host product -- the error must be at the level
of ONE float64 rounding of the result, i.e. below what the fp64 DMMA kernel itself leaves; (2) a fit whose trailing
updates all take the tensor-core route gives the likelihood, factor and posterior of the DMMA fit and of the oracle at
the tolerances of the float64 path (llf 1e-10, L 1e-9, posterior 1e-9)."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib, workloads
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu


def _problem(rows, seed):
    rng = np.random.default_rng(seed)
    P = rng.standard_normal((rows, 64)) * 10.0 ** rng.uniform(-3, 0, size=(rows, 1))
    P[:, ::7] *= 1e-4                      # wide dynamic range inside a row
    P[rows // 3] = 0.0                     # an all-zero row (padding rows of the panel look like this)
    Cm = rng.standard_normal((rows, rows))
    return P, Cm + Cm.T


@pytest.mark.parametrize("rows", [64, 128, 192, 576, 1344])
@pytest.mark.parametrize("digits", [7, 8])
def test_oz_syrk_matches_extended_precision(rows, digits):
    P, Cm = _problem(rows, rows + digits)
    eng = _lib.Engine(0)
    got, _ = eng.debug_oz_syrk(P, Cm, digits=digits)
    ref, _ = eng.debug_oz_syrk(P, Cm, digits=0)       # fp64 DMMA kernel
    exact = Cm.astype(np.longdouble) - P.astype(np.longdouble) @ P.astype(np.longdouble).T
    tril = np.tril_indices(rows)
    e_tc = np.abs(got.astype(np.longdouble) - exact)[tril].astype(np.float64)
    e_64 = np.abs(ref.astype(np.longdouble) - exact)[tril].astype(np.float64)
    scale = np.abs(P).max(1)
    # two roundings of the result (C - hi, then - lo: both terms exact) + the dropped digit pairs (< 2^-(8 digits - 6) of the
    # row-scale product per term)
    allowed = 2.0 ** -51 * np.abs(exact.astype(np.float64))[tril] + 64 * 2.0 ** -(8 * digits - 6) * (scale[:, None] * scale[None, :])[tril] + 1e-300
    assert (e_tc <= allowed).all(), float((e_tc / allowed).max())
    assert e_tc.max() <= 4.0 * e_64.max() + 1e-300     # never worse than the float64 kernel's own rounding level
    # rows of C above the diagonal tiles are untouched
    if rows >= 256:
        assert np.array_equal(got[0, 200:], Cm[0, 200:])


@pytest.mark.parametrize("digits", [7, 8])
@pytest.mark.parametrize("N,D,corr,corr_id", [(1500, 6, "matern52", go.CORR_MATERN52), (2304, 8, "squared_exponential", go.CORR_RBF)])
def test_fit_with_tensor_core_trailing_updates(digits, N, D, corr, corr_id):
    X, y, theta = workloads.canonical_problem(N, D)
    Xc = workloads.canonical_candidates(512, D)

    def fit(tc):
        gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-6)
        gp._check_data(X, y)
        gp.engine.set_chol_tc(tc, 128)     # every trailing matrix of >= 128 rows takes the tensor-core route
        return gp, gp.fit_fixed(X, y, theta, 1.0)

    a, llf_a = fit(digits)
    b, llf_b = fit(0)
    ora = go.fit_fixed(X, y, corr_id, theta, go.MODE_NOISY, sigma2=1.0, noise_var=1e-6)
    assert abs(llf_a - ora.llf) <= 1e-10 * abs(ora.llf), (llf_a, ora.llf)
    assert abs(llf_a - llf_b) <= 1e-10 * abs(llf_b)
    La, Lb = a.C, b.C
    assert np.abs(La - Lb).max() <= 1e-9 * np.abs(Lb).max()
    ya, ma = a.predict(Xc, eval_MSE=True)
    yo, mo = go.predict_chunked(ora, Xc, 256)
    assert np.abs(ya - yo).max() <= 1e-9 * max(1.0, np.abs(yo).max())
    assert np.abs(ma - mo).max() <= 1e-9 * ora.sigma2
