"""GPU parity: the CUDA path (through the C ABI, via the Python mirror) against
  (a) golden vectors produced by the REAL reference (tests/golden/*.npz), and
  (b) the numpy oracle on fresh seeded inputs.
Tolerances (fp64 path): |d yhat| <= 1e-9 max(1,|yhat|), |d MSE| <= 1e-9 sigma2,
acquisition rtol 1e-7 (values above 1e-200), arg-max index exact."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import fit_case, oracle_case

pytestmark = pytest.mark.gpu

MEDIUM = load_golden("medium")
APPB = load_golden("appendix_b")
CANON = load_golden("canonical")
FINITE = sorted(k for k, c in MEDIUM.items() if np.isfinite(c["llf"]) and int(c["trend"]) == 0)


def check_against_golden(c, gp, llf, Xc, ill=False):
    s2 = float(c["sigma2"])
    f = 100.0 if ill else 1.0  # noiseless (no nugget) cases are ill-conditioned: cond(R) * eps is larger
    assert llf == pytest.approx(float(c["llf"]), rel=1e-10 * f, abs=1e-9 * f)
    np.testing.assert_allclose(np.ravel(gp.sigma2)[0], s2, rtol=1e-10 * f)
    np.testing.assert_allclose(np.ravel(gp.noise_var)[0], float(c["noise_var"]), rtol=1e-10 * f, atol=1e-300)
    np.testing.assert_allclose(np.ravel(gp.mean.beta), c["beta"], rtol=1e-8 * f, atol=1e-11 * f)
    gmax = np.abs(c["gamma"]).max()
    np.testing.assert_allclose(gp.gamma.ravel(), c["gamma"], rtol=1e-7 * f, atol=1e-9 * f * gmax)
    yhat, mse = gp.predict(Xc, eval_MSE=True)
    assert yhat.shape == (len(Xc), 1) and mse.shape == (len(Xc), 1)
    np.testing.assert_allclose(yhat.ravel(), c["yhat"], rtol=1e-9 * f, atol=1e-9 * f)
    np.testing.assert_allclose(mse.ravel(), c["mse"], rtol=0, atol=1e-9 * f * s2)
    np.testing.assert_array_equal(gp.predict(Xc), yhat)  # eval_MSE=False returns the same mean
    mn = bool(c["minimize"])
    # acquisition: device kernel on the GOLDEN moments (isolates the elementwise kernel) ...
    eng = gp.engine
    pl = float(c["plugin"])
    tol = dict(rtol=1e-7, atol=1e-300)
    for acq, par, key in [(_lib.ACQ_EI, 0.0, "ei"), (_lib.ACQ_MGFI, float(c["t"]), "mgfi"),
                          (_lib.ACQ_MGFI, 30.0, "mgfi_big_t"), (_lib.ACQ_UCB, float(c["alpha_ucb"]), "ucb"),
                          (_lib.ACQ_PI, float(c["eps"]), "epi")]:
        bv, bi, vals = eng.acq_from_moments(c["yhat"], c["mse"], acq, mn, pl, [par])
        np.testing.assert_allclose(vals[0], c[key], **tol, err_msg=key)
        assert bi[0] == int(np.argmax(vals[0])) and bv[0] == vals[0][bi[0]]
    # ... and end to end through the acquisition classes, compared where the value is well conditioned
    kw = dict(model=gp, minimize=mn)
    for f_, key in [(b2.EI(**kw), "ei"), (b2.MGFI(t=float(c["t"]), **kw), "mgfi"),
                    (b2.UCB(alpha=float(c["alpha_ucb"]), **kw), "ucb"), (b2.EpsilonPI(epsilon=float(c["eps"]), **kw), "epi")]:
        v = f_(Xc)
        assert v.shape == (len(Xc),)
        big = np.abs(c[key]) > 1e-12
        np.testing.assert_allclose(v[big], c[key][big], rtol=1e-6 * f, err_msg=key)
        np.testing.assert_allclose(v[~big], c[key][~big], atol=1e-12, err_msg=key)
        bv, bi = f_.argmax(Xc)
        assert bi[0] == int(np.argmax(v)) and bv[0] == v[bi[0]]
        assert bi[0] == int(np.argmax(c[key])), key


@pytest.mark.parametrize("name", ["rbf_ok", "m32_sk", "m52_ok"])
def test_appendix_b(name):
    c = APPB[name]
    gp, llf = fit_case(c)
    check_against_golden(c, gp, llf, c["Xc"])
    np.testing.assert_allclose(gp.C, c["L"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(np.abs(gp.G), np.abs(c["G"]), rtol=1e-12) if bool(c["ok"]) else None


def test_appendix_b_literal():
    """SURVEY.md App. B numbers typed in by hand: RBF + ordinary kriging"""
    X = np.sin(1 + np.arange(12).reshape(6, 2))
    y = np.cos(np.arange(6))
    Xc = 0.5 * np.cos(2 + np.arange(6).reshape(3, 2))
    gp = b2.GaussianProcess(mean=b2.constant_trend(2), corr="squared_exponential", thetaL=[1e-3] * 2,
                            thetaU=[1e2] * 2, nugget=1e-2)
    llf = gp.fit_fixed(X, y, [0.7, 1.9], 0.9)
    assert llf == pytest.approx(-29.76571813633345, rel=1e-12)
    assert float(gp.mean.beta[0, 0]) == pytest.approx(-0.1635831801184949, rel=1e-11)
    yh, ms = gp.predict(Xc, eval_MSE=True)
    np.testing.assert_allclose(yh.ravel(), [-0.0212879968128635, -1.0703606826299594, 0.49462771887539203], rtol=1e-10)
    np.testing.assert_allclose(ms.ravel(), [0.1505771278371246, 0.5142866801824562, 0.5986049480634108], rtol=1e-10)
    np.testing.assert_allclose(b2.EI(model=gp)(Xc), [0.00078644264137131, 0.32807539664516655, 0.00814241449387647], rtol=1e-8)
    np.testing.assert_allclose(b2.MGFI(model=gp, t=2)(Xc), [0.00112488791535019, 0.4174286840035375, 0.0081701274655377], rtol=1e-8)
    np.testing.assert_allclose(b2.UCB(model=gp, alpha=0.5)(Xc), [0.17273334726886672, -0.7117917631101354, 0.8814755403967918], rtol=1e-9)
    np.testing.assert_allclose(b2.EpsilonPI(model=gp, epsilon=1e-10)(Xc), [0.00627329194952933, 0.544615245558464, 0.02750048759477497], rtol=1e-8)
    np.testing.assert_allclose(gp.gamma.ravel(), [21.927468519542014, 4.3528151291673565, -1.5503377138831664,
                                                  -21.700729512881814, -4.302315454867244, 1.2730990329228544], rtol=1e-10)


@pytest.mark.parametrize("name", FINITE)
def test_medium(name):
    c = MEDIUM[name]
    gp, llf = fit_case(c)
    check_against_golden(c, gp, llf, c["Xc"], ill="_nl_" in name)


def test_rejected_and_not_spd():
    c = MEDIUM["rejected"]
    gp, llf = fit_case(c)
    assert np.isneginf(llf) and not gp.is_fitted                     # llf > 0 -> -inf (gpr.py:981-982)
    with pytest.raises(b2.B200BOError):
        gp.engine.predict(c["Xc"])                                   # no stale state is served
    # exact duplicate rows, no nugget: R is singular -> Cholesky fails -> -inf (gpr.py:946)
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (50, 3))
    X[7] = X[3]
    y = rng.normal(size=50)
    gp = b2.GaussianProcess(mean=b2.constant_trend(3), thetaL=[1e-3] * 3, thetaU=[1e2] * 3, nugget=None)
    assert np.isneginf(gp.fit_fixed(X, y, [1.0, 1.0, 1.0]))
    st = gp.engine.factor(_lib.CORR_RBF, [1.0] * 3, _lib.MODE_NOISELESS)
    assert st[3] == _lib.FIT_NOT_SPD


@pytest.mark.parametrize("name", ["C2", "C5", "C3"])
def test_canonical(name):
    """BASELINE.json config shapes, canonical inputs, reference outputs on 256 candidates"""
    c = CANON[name]
    N, D = int(c["N"]), int(c["D"])
    X, y, theta = go.canonical_problem(N, D)
    Xc = go.canonical_candidates(256, D)
    gp, llf = fit_case(c, X, y)
    check_against_golden(c, gp, llf, Xc)


def test_fresh_inputs_vs_oracle_and_ragged_shapes():
    """oracle on fresh seeded inputs; N, M not multiples of any tile; M crossing the chunk size"""
    rng = np.random.default_rng(99)
    for N, D, corr, cname in [(1, 2, go.CORR_RBF, "squared_exponential"), (63, 3, go.CORR_MATERN32, "matern"),
                              (129, 1, go.CORR_RBF, "squared_exponential"), (333, 11, go.CORR_ABSEXP, "absolute_exponential")]:
        X = rng.uniform(0, 1, (N, D))
        y = np.sin(4 * X).sum(axis=1) + 0.3 * rng.standard_normal(N)
        theta = rng.uniform(0.5, 4.0, D)
        for M in (1, 7, 130):
            Xc = rng.uniform(0, 1, (M, D))
            ora = go.fit_fixed(X, y, corr, theta, go.MODE_NOISY, sigma2=1.1, noise_var=1e-3)
            gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=cname, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-3)
            llf = gp.fit_fixed(X, y, theta, 1.1)
            assert llf == pytest.approx(ora.llf, rel=1e-10, abs=1e-9)
            yo, mo = go.predict(ora, Xc)
            yd, md = gp.predict(Xc, eval_MSE=True)
            np.testing.assert_allclose(yd, yo, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(md, mo, rtol=0, atol=1e-9 * ora.sigma2)
    # M larger than one device chunk (148 * 128 = 18944): chunk seams must be invisible
    N, D, M = 200, 4, 40001
    X = rng.uniform(0, 1, (N, D))
    y = np.cos(3 * X).sum(axis=1) + 0.3 * rng.standard_normal(N)
    theta = np.full(D, 2.0)
    Xc = rng.uniform(0, 1, (M, D))
    ora = go.fit_fixed(X, y, go.CORR_MATERN52, theta, go.MODE_NOISY, sigma2=0.7, noise_var=1e-4)
    from gpu_common import matern
    import functools
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=functools.partial(matern, nu=2.5), thetaL=[1e-5] * D,
                            thetaU=[1e2] * D, nugget=1e-4)
    assert gp.fit_fixed(X, y, theta, 0.7) == pytest.approx(ora.llf, rel=1e-10)
    yo, mo = go.predict_chunked(ora, Xc, 4096)
    yd, md = gp.predict(Xc, eval_MSE=True)
    np.testing.assert_allclose(yd, yo, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(md, mo, rtol=0, atol=1e-9 * ora.sigma2)
    ts = go.mgfi_t_samples(2.0, 5)
    f = b2.MGFI(model=gp, t=2.0)
    vals = f.batch(Xc, ts)
    bv, bi = f.argmax(Xc, ts)
    pl = go.plugin_value(ora.y, True)
    for k, t in enumerate(ts):
        vo = go.mgfi(yo, mo, pl, t)
        np.testing.assert_allclose(vals[k], vo, rtol=1e-6, atol=1e-12)
        assert bi[k] == go.argmax_first(vo) == int(np.argmax(vals[k]))
        assert bv[k] == vals[k][bi[k]]


def test_state_matrices():
    """L L^T = R, L^-1 L = I, and the reference's unscaled-r quirk in noisy mode"""
    c = MEDIUM["rbf_ny_ok"]
    gp = fit_case(c)[0]
    gp.engine.set_keep_R(True)
    gp.fit_fixed(c["X"], c["y"], c["theta"], float(c["par_last"]))
    L, Linv, R = gp.C, gp.engine.state(_lib.STATE_LINV), gp.engine.state(_lib.STATE_R)
    assert np.all(np.triu(L, 1) == 0) and np.all(np.triu(Linv, 1) == 0)
    np.testing.assert_allclose(L @ L.T, R, rtol=0, atol=1e-13)
    np.testing.assert_allclose(Linv @ L, np.eye(len(L)), rtol=0, atol=1e-10)
    ora = oracle_case(c)
    s2, nv = float(c["par_last"]), float(c["nugget"])
    np.testing.assert_allclose(R, (s2 * ora.R0 + nv * np.eye(len(L))) / (s2 + nv), rtol=1e-14, atol=0)
    np.testing.assert_allclose(L, ora.L, rtol=1e-10, atol=1e-13)


def test_device_pointer_path():
    """torch CUDA tensors in / out (B200BO_DEVICE): same numbers as the host-buffer path"""
    import torch

    c = MEDIUM["m52_ny_ok"]
    gp = fit_case(c)[0]
    Xc = np.random.default_rng(5).uniform(-1, 2, (1000, c["X"].shape[1]))
    yh, ms = gp.predict(Xc, eval_MSE=True)
    xd = torch.from_numpy(Xc).cuda()
    yd = torch.empty(1000, dtype=torch.float64, device="cuda")
    md = torch.empty(1000, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    gp.engine.predict_device(xd, yd, md)
    np.testing.assert_array_equal(yd.cpu().numpy(), yh.ravel())
    np.testing.assert_array_equal(md.cpu().numpy(), ms.ravel())
    vd = torch.empty((2, 1000), dtype=torch.float64, device="cuda")
    bv, bi, _ = gp.engine.acq(xd, _lib.ACQ_MGFI, True, float(c["plugin"]), [1.0, 2.0], device_vals=vd)
    v = b2.MGFI(model=gp, t=1.0).batch(Xc, [1.0, 2.0])
    np.testing.assert_array_equal(vd.cpu().numpy(), v)
    assert list(bi) == [int(np.argmax(v[0])), int(np.argmax(v[1]))]


def test_interpolation_and_idempotence():
    """size-independent properties: noiseless GP interpolates its data with ~zero MSE; repeated calls are
    bit-identical (deterministic reductions, no atomics)"""
    c = MEDIUM["m32_nl_ok"]
    gp = fit_case(c)[0]
    yh, ms = gp.predict(c["X"], eval_MSE=True)
    np.testing.assert_allclose(yh.ravel(), c["y"], atol=1e-7)
    assert ms.max() <= 1e-8 * float(c["sigma2"])
    y2, m2 = gp.predict(c["X"], eval_MSE=True)
    np.testing.assert_array_equal(yh, y2)
    np.testing.assert_array_equal(ms, m2)
    assert np.all(b2.EI(model=gp)(c["X"]) == 0.0)  # EI early-out at s ~ 0 (acquisition_fun.py:162-164)
