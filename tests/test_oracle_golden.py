"""CPU: the numpy oracle (oracle/gp_oracle.py) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py) and against the literal known answers of SURVEY.md App. B."""
import numpy as np
import pytest

from oracle import gp_oracle as go
from oracle import ref_loader

from conftest import load_golden

MEDIUM = load_golden("medium")
FINITE = sorted(k for k, c in MEDIUM.items() if np.isfinite(c["llf"]))


def oracle_fit(c, X=None, y=None):
    X = c["X"] if X is None else X
    y = c["y"] if y is None else y
    mode = int(c["mode"])
    kw = dict(trend=int(c["trend"]))
    if not bool(c["ok"]):
        p = go.trend_basis(int(c["trend"]), X[:1]).shape[1]
        kw["beta_fixed"] = np.broadcast_to(np.asarray(c["beta_in"], float).ravel(), (p,)) if np.size(c["beta_in"]) == 1 else c["beta_in"]
    if mode == go.MODE_NOISY:
        kw.update(sigma2=float(c["par_last"]), noise_var=float(c["nugget"]))
    elif mode == go.MODE_NOISE_ESTIM:
        kw.update(alpha=float(c["par_last"]))
    return go.fit_fixed(X, y, int(c["corr"]), c["theta"], mode, **kw)


def check_case(c, gp, Xc, rtol=1e-10):
    assert gp.llf == pytest.approx(float(c["llf"]), rel=1e-11)
    np.testing.assert_allclose(gp.sigma2, c["sigma2"], rtol=1e-11)
    np.testing.assert_allclose(gp.beta.ravel(), c["beta"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gp.gamma.ravel(), c["gamma"], rtol=1e-7, atol=1e-9 * np.abs(c["gamma"]).max())
    yhat, mse = go.predict(gp, Xc)
    np.testing.assert_allclose(yhat.ravel(), c["yhat"], rtol=rtol, atol=1e-10)
    np.testing.assert_allclose(mse.ravel(), c["mse"], rtol=rtol, atol=1e-11 * gp.sigma2)
    mn = bool(c["minimize"])
    pl = go.plugin_value(gp.y, mn)
    assert pl == pytest.approx(float(c["plugin"]), rel=1e-15)
    # acquisition from the GOLDEN (yhat, mse).  The golden acquisition values come from the reference's
    # one-row-at-a-time calls, whose M=1 predict differs from the batched predict in the last bits
    # (gemv vs gemm); Phi(z) for z ~ -15..-30 amplifies that by ~z^2, hence rtol 1e-7.
    a = dict(rtol=1e-7, atol=1e-300)
    np.testing.assert_allclose(go.ei(c["yhat"], c["mse"], gp.sigma2, pl, mn), c["ei"], **a)
    np.testing.assert_allclose(go.mgfi(c["yhat"], c["mse"], pl, float(c["t"]), mn), c["mgfi"], **a)
    np.testing.assert_allclose(go.mgfi(c["yhat"], c["mse"], pl, 30.0, mn), c["mgfi_big_t"], **a)
    np.testing.assert_allclose(go.ucb(c["yhat"], c["mse"], float(c["alpha_ucb"]), mn), c["ucb"], rtol=1e-13)
    np.testing.assert_allclose(go.pi_eps(c["yhat"], c["mse"], pl, float(c["eps"]), mn), c["epi"], **a)
    for name, v in (("ei", go.ei(c["yhat"], c["mse"], gp.sigma2, pl, mn)),
                    ("mgfi", go.mgfi(c["yhat"], c["mse"], pl, float(c["t"]), mn))):
        assert go.argmax_first(v) == int(np.argmax(c[name]))


def test_appendix_b_literals(golden_appendix_b):
    """SURVEY.md App. B known answers, typed in from the survey (independent of the .npz)."""
    X = np.sin(1 + np.arange(12).reshape(6, 2))
    y = np.cos(np.arange(6))
    Xc = 0.5 * np.cos(2 + np.arange(6).reshape(3, 2))
    gp = go.fit_fixed(X, y, go.CORR_RBF, [0.7, 1.9], go.MODE_NOISY, sigma2=0.9, noise_var=1e-2)
    assert gp.llf == pytest.approx(-29.76571813633345, rel=1e-13)
    assert gp.beta[0, 0] == pytest.approx(-0.1635831801184949, rel=1e-12)
    yh, ms = go.predict(gp, Xc)
    np.testing.assert_allclose(yh.ravel(), [-0.0212879968128635, -1.0703606826299594, 0.49462771887539203], rtol=1e-11)
    np.testing.assert_allclose(ms.ravel(), [0.1505771278371246, 0.5142866801824562, 0.5986049480634108], rtol=1e-12)
    pl = go.plugin_value(gp.y, True)
    assert pl == -0.9899924966004454
    np.testing.assert_allclose(go.ei(yh, ms, 0.9, pl), [0.00078644264137131, 0.32807539664516655, 0.00814241449387647], rtol=1e-10)
    np.testing.assert_allclose(go.mgfi(yh, ms, pl, 2.0), [0.00112488791535019, 0.4174286840035375, 0.0081701274655377], rtol=1e-10)
    np.testing.assert_allclose(go.ucb(yh, ms, 0.5), [0.17273334726886672, -0.7117917631101354, 0.8814755403967918], rtol=1e-11)
    np.testing.assert_allclose(go.pi_eps(yh, ms, pl, 1e-10), [0.00627329194952933, 0.544615245558464, 0.02750048759477497], rtol=1e-10)
    gp = go.fit_fixed(X, y, go.CORR_MATERN32, [0.7, 1.9], go.MODE_NOISY, sigma2=0.9, noise_var=1e-2, beta_fixed=[0.0])
    assert gp.llf == pytest.approx(-26.50671301987987, rel=1e-13)
    yh, ms = go.predict(gp, Xc)
    np.testing.assert_allclose(ms.ravel(), [0.17340396972355407, 0.45651587932213433, 0.48058696491508407], rtol=1e-12)
    gp = go.fit_fixed(X, y, go.CORR_MATERN52, [0.7, 1.9], go.MODE_NOISY, sigma2=0.9, noise_var=1e-2)
    assert gp.llf == pytest.approx(-34.8065818516643, rel=1e-13)
    assert gp.beta[0, 0] == pytest.approx(-0.11726009971460194, rel=1e-12)


@pytest.mark.parametrize("name", ["rbf_ok", "m32_sk", "m52_ok"])
def test_appendix_b_npz(golden_appendix_b, name):
    c = golden_appendix_b[name]
    gp = oracle_fit(c)
    check_case(c, gp, c["Xc"])
    np.testing.assert_allclose(gp.L, c["L"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", FINITE)
def test_medium(name):
    c = MEDIUM[name]
    gp = oracle_fit(c)
    check_case(c, gp, c["Xc"], rtol=1e-8 if "_nl_" in name else 1e-10)
    if "llf_grad" in c:
        alpha = float(c["par_last"]) if int(c["mode"]) == go.MODE_NOISE_ESTIM else None
        g = go.llf_grad(gp, alpha)
        np.testing.assert_allclose(g, c["llf_grad"], rtol=1e-7, atol=1e-8 * np.abs(c["llf_grad"]).max())
    if "y_dx" in c:
        for i in range(c["y_dx"].shape[0]):
            ydx, mdx = go.posterior_gradient(gp, c["Xc"][i])
            ok = np.isfinite(c["y_dx"][i])
            np.testing.assert_allclose(ydx.ravel()[ok], c["y_dx"][i][ok], rtol=1e-7, atol=1e-9)
            ok = np.isfinite(c["mse_dx"][i])
            np.testing.assert_allclose(mdx.ravel()[ok], c["mse_dx"][i][ok], rtol=1e-6, atol=1e-9)


def test_rejected_likelihood():
    c = MEDIUM["rejected"]
    assert np.isneginf(c["llf"])
    gp = oracle_fit(c)
    assert np.isneginf(gp.llf)  # gpr.py:981-982: llf > 0 -> -inf


@pytest.mark.parametrize("name", ["C2", "C5", "C3"])
def test_canonical(golden_canonical, name):
    c = golden_canonical[name]
    N, D = int(c["N"]), int(c["D"])
    if N > 2048:
        pytest.importorskip("scipy")
    X, y, theta = go.canonical_problem(N, D)
    np.testing.assert_array_equal(theta, c["theta"])
    Xc = go.canonical_candidates(256, D)
    gp = oracle_fit(c, X, y)
    check_case(c, gp, Xc, rtol=1e-9)


KNOWN_LLF = {"C2": -1751.8164566295322, "C3": -4647.235533742902, "C5": -2276.4175051994885}


@pytest.mark.parametrize("name", sorted(KNOWN_LLF))
def test_canonical_known_llf(golden_canonical, name):
    """Literal likelihoods from SURVEY.md §8d / BASELINE.md §3."""
    assert float(golden_canonical[name]["llf"]) == pytest.approx(KNOWN_LLF[name], rel=1e-14)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
def test_oracle_vs_live_reference():
    """Fresh random problem, oracle vs the live reference objects (not via the .npz)."""
    ns = ref_loader.load()
    rng = np.random.default_rng(123)
    N, D, M = 150, 4, 40
    X = rng.uniform(0, 1, (N, D))
    y = np.cos(3 * X).sum(axis=1) + 0.2 * rng.standard_normal(N)
    Xc = rng.uniform(0, 1, (M, D))
    theta = rng.uniform(0.5, 3.0, D)
    ref = ns.GaussianProcess(mean=ns.constant_trend(D), corr="matern", thetaL=[1e-3] * D, thetaU=[1e2] * D, nugget=1e-3)
    llf = ref_loader.fixed_theta_fit(ref, X, y, theta, 1.3)
    gp = go.fit_fixed(X, y, go.CORR_MATERN32, theta, go.MODE_NOISY, sigma2=1.3, noise_var=1e-3)
    assert gp.llf == pytest.approx(llf, rel=1e-12)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yr, mr = ref.predict(Xc, eval_MSE=True)
        er = np.array([float(np.sum(ns.EI(model=ref, minimize=True)(x))) for x in Xc])
    yo, mo = go.predict(gp, Xc)
    np.testing.assert_allclose(yo, yr, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(mo, mr, rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(go.ei(yo, mo, gp.sigma2, go.plugin_value(gp.y, True)), er, rtol=1e-10, atol=1e-300)


def test_parameter_recipes():
    """bayes_opt.py:85, :89 samplers use the global numpy RNG; :127-130 exponential annealing."""
    np.random.seed(42)
    xi = np.random.randn(4)
    np.testing.assert_allclose(go.mgfi_t_samples(2.0, 4, 42), np.exp(np.log(2.0) + 0.5 * xi))
    np.testing.assert_allclose(go.ucb_alpha_samples(0.5, 4, 42), 1 / (1 + np.exp(0.0 + 0.6 * xi)))
    s = go.annealing_t_schedule(2.0, 0.1, 32)
    assert s[0] == 2.0 and s[-1] * (0.1 / 2.0) ** (1 / 32) == pytest.approx(0.1)


ACQ_GRAD = load_golden("acq_grad")


@pytest.mark.parametrize("name", sorted(ACQ_GRAD))
def test_acquisition_gradients(name):
    """return_dx=True of the reference's acquisition classes (one row at a time) and gp.gradient against the oracle's
    posterior_gradient + acquisition_dx; the first three candidates are training points (early-outs, Matern 0/0 quirk)."""
    c = ACQ_GRAD[name]
    gp = oracle_fit(c)
    mn = bool(c["minimize"])
    pl = go.plugin_value(gp.y, mn)
    assert pl == pytest.approx(float(c["plugin"]), rel=1e-15)
    for i in range(c["Xc"].shape[0]):
        ydx, mdx = go.posterior_gradient(gp, c["Xc"][i])
        np.testing.assert_allclose(ydx.ravel(), c["y_dx"][i], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(mdx.ravel(), c["mse_dx"][i], rtol=1e-6, atol=1e-9)
        for key, acq, par in (("ei", go.ACQ_EI, 0.0), ("ucb", go.ACQ_UCB, float(c["alpha_ucb"])),
                              ("mgfi", go.ACQ_MGFI, float(c["t"])), ("mgfi_big", go.ACQ_MGFI, 30.0),
                              ("epi", go.ACQ_PI, float(c["eps"]))):
            v, dx = go.acquisition_dx(acq, c["yhat"][i], c["mse"][i], c["y_dx"][i], c["mse_dx"][i], gp.sigma2, pl, par, mn)
            assert v == pytest.approx(float(c[key][i]), rel=1e-6, abs=1e-300), (key, i)
            ref = c[key + "_dx"][i]
            if np.all(np.isfinite(ref)):
                np.testing.assert_allclose(dx, ref, rtol=1e-6, atol=1e-9 * max(1.0, np.abs(ref).max()), err_msg=f"{key} {i}")
            else:  # sd = 0 at a training point: the reference divides by it (UCB / PI) and returns nan / inf
                assert not np.all(np.isfinite(dx)), (key, i)


RESTRICTED = load_golden("restricted")


def oracle_restricted(c, eval_grad=False):
    kw = {}
    if not bool(c["ok"]):
        kw["beta_fixed"] = [float(np.ravel(c["beta_in"])[0])]
    return go.fit_fixed_restricted(c["X"], c["y"], int(c["corr"]), c["theta"], float(c["sigma2"]), float(c["noise_var"]),
                                   eval_grad=eval_grad, n_par=c["llf_grad"].size, **kw)


@pytest.mark.parametrize("name", sorted(RESTRICTED))
def test_restricted_likelihood(name):
    """likelihood="restricted" (gpr.py:813-918): value, gradient and the fitted state against the reference"""
    c = RESTRICTED[name]
    gp, grad = oracle_restricted(c, eval_grad=True)
    # the noiseless RBF matrix (no nugget) has condition ~1e16: rho^T rho carries only a few digits
    loose = "_nl_" in name and "rbf" in name
    assert gp.llf == pytest.approx(float(c["llf"]), rel=1e-3 if loose else 1e-10)
    np.testing.assert_allclose(grad, c["llf_grad"], rtol=1e-2 if loose else 1e-7, atol=1e-7 * np.abs(c["llf_grad"]).max())
    if not loose:
        yh, ms = go.predict(gp, c["Xc"])
        np.testing.assert_allclose(yh.ravel(), c["yhat"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(ms.ravel(), c["mse"], rtol=1e-8, atol=1e-11)


TRENDS = load_golden("trends")


@pytest.mark.parametrize("name", sorted(TRENDS))
def test_trends(name):
    """linear / quadratic regression trends (trend.py:94-142) through fit state, predict and the acquisition values"""
    c = TRENDS[name]
    gp = oracle_fit(c)
    check_case(c, gp, c["Xc"], rtol=1e-8 if "_nl_" in name else 1e-10)


GENEXP = load_golden("genexp")


@pytest.mark.parametrize("name", sorted(GENEXP))
def test_generalized_exponential(name):
    """generalized_exponential (kernel.py:332-374), theta = [theta_1..n, p]"""
    c = GENEXP[name]
    gp = oracle_fit(c)
    check_case(c, gp, c["Xc"], rtol=1e-7 if "_nl_" in name else 1e-10)


MATERN_NU = load_golden("matern_nu")


@pytest.mark.parametrize("name", sorted(MATERN_NU))
def test_matern_general_nu(name):
    """matern(nu) through scipy.special.kv (kernel.py:201-207); nu rides behind theta in the oracle / device convention"""
    c = dict(MATERN_NU[name])
    c["theta"] = np.r_[c["theta"], float(c["nu"])]
    gp = oracle_fit(c)
    if not np.isfinite(c["llf"]):
        assert np.isneginf(gp.llf)       # llf > 0 is rejected (gpr.py:981-982)
        return
    check_case(c, gp, c["Xc"], rtol=1e-7 if "_nl_" in name else 1e-9)
