"""CPU: the scalar formulas the kernels evaluate (csrc/gp_math.h, compiled for the host by the test shim
csrc/hostmath.cpp) against the oracle and against the reference-generated golden vectors."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import gp_oracle as go

from conftest import load_golden


@pytest.fixture(scope="module")
def hm():
    from bayesian_optimization_b200 import build

    build.build()
    lib = C.CDLL(build.HOSTMATH)
    lib.b2h_corr.restype = C.c_double
    lib.b2h_corr.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.b2h_acq.restype = C.c_double
    lib.b2h_acq.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    lib.b2h_arg_better.restype = C.c_int
    lib.b2h_arg_better.argtypes = [C.c_double, C.c_longlong, C.c_double, C.c_longlong]
    return lib


@pytest.mark.parametrize("corr", range(6))
def test_corr_matches_oracle(hm, corr):
    rng = np.random.default_rng(corr)
    D = 7
    theta = rng.uniform(0.05, 2.0, D) * (0.3 if corr == go.CORR_CUBIC else 1.0)
    X = rng.uniform(-1, 1, (40, D))
    Y = rng.uniform(-1, 1, (40, D))
    Y[:5] = X[:5]  # zero distance
    want = go.corr_values(corr, theta, np.abs(X - Y))
    got = np.array([hm.b2h_corr(corr, theta.ctypes.data, np.ascontiguousarray(x).ctypes.data,
                                np.ascontiguousarray(y).ctypes.data, D) for x, y in zip(X, Y)])
    np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-300)
    assert np.all(got[:5] == 1.0)


MEDIUM = load_golden("medium")


@pytest.mark.parametrize("name", ["rbf_ny_ok", "m52_ny_ok", "rbf_ny_ok_max", "abs_nl_sk", "cub_ne_ok", "m32_nl_sk"])
def test_acq_matches_golden(hm, name):
    """device formulas on the golden (yhat, mse) vs the reference's one-row-at-a-time values"""
    c = MEDIUM[name]
    mn = int(bool(c["minimize"]))
    s2, pl = float(c["sigma2"]), float(c["plugin"])

    def run(acq, par):
        return np.array([hm.b2h_acq(acq, y, m, s2, pl, par, mn) for y, m in zip(c["yhat"], c["mse"])])

    tol = dict(rtol=1e-7, atol=1e-300)  # see tests/test_oracle_golden.py for why 1e-7
    np.testing.assert_allclose(run(go.ACQ_EI, 0.0), c["ei"], **tol)
    np.testing.assert_allclose(run(go.ACQ_MGFI, float(c["t"])), c["mgfi"], **tol)
    np.testing.assert_allclose(run(go.ACQ_MGFI, 30.0), c["mgfi_big_t"], **tol)
    np.testing.assert_allclose(run(go.ACQ_UCB, float(c["alpha_ucb"])), c["ucb"], rtol=1e-14)
    np.testing.assert_allclose(run(go.ACQ_PI, float(c["eps"])), c["epi"], **tol)


def test_acq_edge_cases(hm):
    # EI early-out: sd/sqrt(sigma2) < 1e-6 -> 0   (acquisition_fun.py:162-164)
    assert hm.b2h_acq(go.ACQ_EI, -5.0, 1e-13 * 0.9, 1.0, 0.0, 0.0, 1) == 0.0
    assert hm.b2h_acq(go.ACQ_EI, -5.0, 0.0, 1.0, 0.0, 0.0, 1) == 0.0
    # MGFI: isclose(sd, 0) -> 0 (:274); overflow -> 0 (:277-290)
    assert hm.b2h_acq(go.ACQ_MGFI, -5.0, 1e-17, 1.0, 0.0, 2.0, 1) == 0.0
    assert hm.b2h_acq(go.ACQ_MGFI, -1e4, 4.0, 1.0, 0.0, 22.36, 1) == 0.0
    want = go.mgfi(np.array([-1e4, 0.3]), np.array([4.0, 0.2]), 0.0, 22.36)
    assert want[0] == 0.0
    assert hm.b2h_acq(go.ACQ_MGFI, 0.3, 0.2, 1.0, 0.0, 22.36, 1) == pytest.approx(want[1], rel=1e-12)
    # maximisation flips yhat (:61-62)
    a = hm.b2h_acq(go.ACQ_UCB, 0.7, 0.04, 1.0, 0.0, 0.5, 0)
    assert a == pytest.approx(-0.7 + 0.5 * 0.2)
    # random sweep against the oracle
    rng = np.random.default_rng(0)
    y = rng.normal(size=200)
    m = rng.uniform(0, 2, 200) ** 2
    for acq, par in [(go.ACQ_EI, 0.0), (go.ACQ_PI, 0.05), (go.ACQ_PI, 0.0), (go.ACQ_UCB, 0.3), (go.ACQ_MGFI, 1.7)]:
        for mn in (True, False):
            want = go.acquisition(acq, y, m, 1.3, -0.4, par, mn)
            got = np.array([hm.b2h_acq(acq, a_, b_, 1.3, -0.4, par, int(mn)) for a_, b_ in zip(y, m)])
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-300)  # deep tails: erfc vs ndtr last bits x z^2


def test_argmax_ordering(hm):
    """numpy.argmax rule: NaN wins, then larger value, then lower index"""
    nan = float("nan")
    assert hm.b2h_arg_better(2.0, 5, 1.0, 0) == 1
    assert hm.b2h_arg_better(1.0, 0, 1.0, 5) == 1
    assert hm.b2h_arg_better(1.0, 5, 1.0, 0) == 0
    assert hm.b2h_arg_better(nan, 7, 1e300, 0) == 1
    assert hm.b2h_arg_better(1e300, 0, nan, 7) == 0
    assert hm.b2h_arg_better(nan, 3, nan, 7) == 1
    v = np.array([0.1, np.nan, 3.0, np.nan, 3.0])
    best, bi = v[0], 0
    for i in range(1, len(v)):
        if hm.b2h_arg_better(v[i], i, best, bi):
            best, bi = v[i], i
    assert bi == int(np.argmax(v))


def test_bessel_kv_and_general_matern_match_scipy(hm):
    """the device's K_nu (Temme's method, gp_math.h) against scipy.special.kv, which the reference calls (kernel.py:207)"""
    import ctypes as C

    from scipy.special import gamma, kv

    hm.b2h_kv.restype = C.c_double
    hm.b2h_kv.argtypes = [C.c_double, C.c_double]
    hm.b2h_matern.restype = C.c_double
    hm.b2h_matern.argtypes = [C.c_double, C.c_double]
    for nu in (0.1, 0.3, 0.8, 1.0, 1.49, 1.51, 2.0, 3.0, 3.5, 4.7, 7.25, 10.0):
        for x in np.r_[np.logspace(-12, 0.3, 40), np.linspace(2.0001, 60, 40), 2.0]:
            assert hm.b2h_kv(nu, float(x)) == pytest.approx(kv(nu, x), rel=5e-13), (nu, x)
        for h in np.r_[0.0, np.logspace(-10, 1.5, 60)]:
            hh = h if h > 0 else np.finfo(float).eps
            t = np.sqrt(2 * nu) * hh
            assert hm.b2h_matern(float(h), nu) == pytest.approx((2 ** (1 - nu)) / gamma(nu) * t**nu * kv(nu, t), abs=1e-13), (nu, h)


def test_thin_qr_matches_scipy_sign_for_sign(hm):
    """the host-side Householder QR of the (N, p) trend panel (csrc/host_qr.h, used by b200bo_factor for p > 1) against
    scipy.linalg.qr(mode="economic") -- what the reference calls (gpr.py:805) -- incl. LAPACK's sign convention of R"""
    import ctypes as C

    from scipy.linalg import qr, solve_triangular

    hm.b2h_thin_qr.restype = None
    hm.b2h_thin_qr.argtypes = [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p] * 3
    rng = np.random.default_rng(3)
    for N, p in ((50, 1), (50, 4), (200, 10), (64, 64), (300, 28)):
        Ft = np.ascontiguousarray(rng.standard_normal((N, p)) * rng.uniform(0.1, 10, p))
        Ft[:, 0] = np.abs(Ft[:, 0]) + 0.1 if p > 1 else -np.abs(Ft[:, 0])       # both signs of the leading pivot
        yt = rng.standard_normal(N)
        G, beta, rho = np.empty((p, p)), np.empty(p), np.empty(N)
        hm.b2h_thin_qr(Ft.ctypes.data, yt.ctypes.data, N, p, G.ctypes.data, beta.ctypes.data, rho.ctypes.data)
        Q, R = qr(Ft, mode="economic")
        np.testing.assert_allclose(G, R, rtol=1e-11, atol=1e-12 * np.abs(R).max())
        np.testing.assert_allclose(beta, solve_triangular(R, Q.T @ yt), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(rho, yt - Q @ (Q.T @ yt), rtol=1e-9, atol=1e-12)
