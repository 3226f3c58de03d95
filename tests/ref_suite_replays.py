"""Replays of the reference's own test sets for the hot path, as plain functions of the host-API module `ag` (no pytest
marks here): tests/test_api_on_fake_lib.py runs them on the CPU against the oracle-backed fake library (that pins the
API semantics and the replay logic), tests/test_gpu_variants_grad_vfecov.py runs them on a device (opt-in until they have passed
there once).  File:line citations are relative to /root/reference.  Dense Sigma_y, AD (`adjoint_test`) and
`update_posterior` cases are outside the device path and are not replayed."""
import numpy as np


def approx(a, b, rtol=1.5e-8, atol=0.0):  # Julia's isapprox, norm-wise
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= max(atol, rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


def finite_gp_statistics(ag):
    """test/finite_gp_projection.jl:26-62 (zero observation noise, N = 1 and N' = 9, Row/ColVecs sugar)."""
    rng = np.random.default_rng(123456)
    N, Np = 1, 9
    x, xp = rng.standard_normal(N), rng.standard_normal(Np)
    Xmat = rng.standard_normal((N, Np))
    f = ag.GP(np.sin, ag.SqExponentialKernel())
    fx, fxp = f(x, 0.0), f(xp, 0.0)
    for xw in (ag.RowVecs(Xmat), ag.ColVecs(Xmat)):
        assert isinstance(f(xw), ag.FiniteGP) and isinstance(f(xw, 1e-3), ag.FiniteGP)
    assert np.array_equal(ag.mean(fx), ag.mean(f, x))
    assert np.array_equal(ag.cov(fx), ag.cov(f, x))
    assert np.array_equal(ag.var(fx), np.diag(ag.cov(fx)))
    assert np.array_equal(ag.cov(fx, fxp), ag.cov(f, x, xp))
    assert np.array_equal(ag.marginals(fx).mu, ag.mean(f(x)))
    assert approx(ag.marginals(fx).sigma ** 2, ag.var(f, x), rtol=1e-15)
    m, C = ag.mean_and_cov(fx)
    assert np.array_equal(m, ag.mean(fx)) and np.array_equal(C, ag.cov(fx))
    m, c = ag.mean_and_var(fx)
    assert np.array_equal(m, ag.mean(fx)) and np.array_equal(c, ag.var(fx))
    # the same for the 9-point projection (a real Gram)
    assert np.array_equal(ag.var(fxp), np.diag(ag.cov(fxp)))


def finite_gp_rand_statistical(ag, S=100_000):
    """test/finite_gp_projection.jl:86-106: sample mean and covariance of 100 000 draws converge to 1e-2."""
    rng = np.random.default_rng(123456)
    x = np.linspace(-3.0, 3.0, 10)
    fx = ag.GP(1, ag.SqExponentialKernel())(x, 1e-12)
    F = ag.rand(rng, fx, S)
    assert F.shape == (10, S)
    mu = ag.mean(fx)
    assert np.max(np.abs(F.mean(1) - mu)) < 1e-2
    Sig = (F - mu[:, None]) @ (F - mu[:, None]).T / S
    assert np.mean(np.abs(Sig - ag.cov(fx))) < 1e-2


def finite_gp_logpdf(ag):
    """test/finite_gp_projection.jl:131-151: logpdf vs an independent multivariate normal, matrix form vs columns,
    loglikelihood = sum."""
    from scipy.stats import multivariate_normal
    rng = np.random.default_rng(123456)
    N, S, sig = 10, 11, 1e-1
    x = np.linspace(-3.0, 3.0, N)
    f = ag.GP(1, ag.SqExponentialKernel())
    y = f(x, sig ** 2)
    yh = ag.rand(rng, y)
    lp = ag.logpdf(y, yh)
    assert np.ndim(lp) == 0
    assert approx(lp, multivariate_normal(mean=ag.mean(y), cov=ag.cov(y)).logpdf(yh))
    assert ag.loglikelihood(y, yh) == lp
    Yh = ag.rand(rng, y, S)
    lps = ag.logpdf(y, Yh)
    assert lps.shape == (S,) and lps.dtype == np.float64
    assert approx(lps, [ag.logpdf(y, np.ascontiguousarray(Yh[:, n])) for n in range(S)])
    assert ag.loglikelihood(y, Yh) == np.sum(lps)


def finite_gp_type_stability(ag, T):
    """test/finite_gp_projection.jl:180-191."""
    rng = np.random.default_rng(123456)
    x = rng.standard_normal(123).astype(T)
    f = ag.GP(T(0), ag.SqExponentialKernel())
    fx = f(x, T(0.1))
    y = ag.rand(rng, fx)
    assert y.dtype == T and y.shape == (123,)
    assert ag.logpdf(fx, y).dtype == T


def sparse_approx_log_evidence(ag, Approx):
    """test/sparse_approximations.jl:87-103 for ApproxType in (VFE, DTC)."""
    rng = np.random.default_rng(123456)
    x = np.linspace(-1.0, 1.0, 3)
    f = ag.GP(ag.SqExponentialKernel())
    fx = f(x, 0.1)
    y = ag.rand(rng, fx)
    ev = ag.approx_log_evidence(Approx(f(x)), fx, y)
    assert np.ndim(ev) == 0
    assert approx(ev, ag.logpdf(fx, y))
    if Approx is ag.VFE:
        assert ag.elbo(Approx(f(x)), fx, y) == ag.approx_log_evidence(Approx(f(x)), fx, y)
        assert ag.elbo(Approx(f(x + rng.standard_normal(3))), fx, y) < ag.logpdf(fx, y)


def sparse_type_stability(ag, Approx, T):
    """test/sparse_approximations.jl:105-119 (the part the device path covers: objective and posterior marginals)."""
    rng = np.random.default_rng(123456)
    x = np.linspace(-1.0, 1.0, 3).astype(T)
    f = ag.GP(T(0), ag.SqExponentialKernel())
    fx = f(x, T(0.1))
    y = ag.rand(rng, fx)
    jitter = T(1e-12) if T == np.float64 else T(1e-6)  # K_zz of 3 points in fp32 needs a representable jitter
    assert ag.approx_log_evidence(Approx(f(x, jitter)), fx, y).dtype == T
    post = ag.posterior(Approx(f(x, jitter)), fx, y)
    m, v = ag.mean_and_var(post(x, jitter))
    assert m.dtype == T and v.dtype == T and np.array_equal(ag.inducing_points(post).a[:, 0], x)


def sparse_posterior_matches_exact(ag, Approx):
    """test/sparse_approximations.jl:3-25 restricted to marginals: with z = x the optimal approximate posterior is the
    exact posterior (mean and var at 100 random inputs)."""
    rng = np.random.default_rng(123456)
    f = ag.GP(np.sin, ag.Matern32Kernel())
    x = np.linspace(-1.0, 1.0, 3)
    fx = f(x, 1e-15)
    y = ag.rand(rng, fx)
    f_post = ag.posterior(fx, y)
    f_approx = ag.posterior(Approx(f(x, 1e-12)), fx, y)
    xt = rng.standard_normal(100)
    assert approx(ag.mean(f_post, xt), ag.mean(f_approx, xt), rtol=1e-6)
    assert approx(ag.var(f_post, xt), ag.var(f_approx, xt), rtol=1e-6, atol=1e-9)


def sparse_update_posterior(ag, Approx):
    """test/sparse_approximations.jl:27-85: online (update_posterior) vs batch, new observations and new pseudo-points.
    The reference compares its host caches field by field (atol 1e-5); the device handle exposes predictions, so the two
    routes are compared there, plus the inducing points."""
    rng = np.random.default_rng(1)
    X, y = rng.random(10), rng.random(10)
    Z = rng.random(4)
    f = ag.GP(ag.SqExponentialKernel())
    xt = np.linspace(-0.2, 1.2, 23)
    p1 = ag.posterior(Approx(f(Z)), f(X[:7], 0.1), y[:7])
    u1 = ag.update_posterior(p1, f(X[7:], 0.1), y[7:])
    p2 = ag.posterior(Approx(f(Z)), f(X, 0.1), y)
    assert approx(ag.inducing_points(u1).a, ag.inducing_points(p2).a, atol=1e-5)
    for a, b in zip(ag.mean_and_var(u1, xt), ag.mean_and_var(p2, xt)):
        assert approx(a, b, rtol=1e-6, atol=1e-5)
    assert isinstance(u1.approx, Approx)
    Z1, Z2 = rng.random(4), rng.random(3)
    q1 = ag.posterior(Approx(f(Z1)), f(X, 0.1), y)
    v1 = ag.update_posterior(q1, f(Z2))
    q2 = ag.posterior(Approx(f(np.concatenate([Z1, Z2]))), f(X, 0.1), y)
    assert approx(ag.inducing_points(v1).a, ag.inducing_points(q2).a, atol=1e-5)
    for a, b in zip(ag.mean_and_var(v1, xt), ag.mean_and_var(q2, xt)):
        assert approx(a, b, rtol=1e-6, atol=1e-5)
    assert isinstance(v1.approx, Approx)


def sparse_internal_interface(ag, Approx, testutils):
    """test/sparse_approximations.jl:12-25: the optimal approximate posterior with z = x equals the exact posterior (mean AND
    full covariance at 100 inputs), and TestUtils.test_internal_abstractgps_interface runs on it.  `testutils` is the
    module holding the TestUtils mirror (tests/test_gpu_posterior_finitegp.py).  Exact conditioning ON TOP of an
    approximate posterior (`posterior(f_approx(x, s2), y)`, the last line of the primary interface) is outside the device
    path and is skipped."""
    rng = np.random.default_rng(123456)
    f = ag.GP(np.sin, ag.Matern32Kernel())
    x = np.linspace(-1.0, 1.0, 3)
    fx = f(x, 1e-15)
    y = ag.rand(rng, fx)
    f_post = ag.posterior(fx, y)
    f_approx = ag.posterior(Approx(f(x, 1e-12)), fx, y)
    xt = rng.standard_normal(100)
    assert approx(ag.mean(f_post, xt), ag.mean(f_approx, xt), rtol=1e-6)
    assert approx(ag.cov(f_post, xt), ag.cov(f_approx, xt), rtol=1e-6, atol=1e-8)
    a, b = np.linspace(-1.0, 1.0, 5), rng.standard_normal(6)
    testutils.internal_abstractgps_interface(ag, rng, f_approx, a, b, vfe=False, conditioning=False)
