"""The reference's own jldoctests, replayed through the drop-in host API on the CUDA path.

Every test below is one ```jldoctest``` block of /root/reference/src (file:line in the docstring) with
`randn`/`rand` replaced by a seeded numpy Generator.  Where the Julia doctest asserts `==` between two
results that the reference computes by the SAME route (e.g. cov(f(x)) and kernelmatrix(k, x)) we also
assert bit equality; where the two sides take different routes we use BASELINE.json's fp64 tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-8


@pytest.fixture()
def rng():
    return np.random.default_rng(20240924)


def test_base_gp_zero_mean(ag, rng):
    """src/base_gp.jl:9-19."""
    f = ag.GP(ag.Matern32Kernel())
    x = rng.standard_normal(5)
    assert np.array_equal(ag.mean(f(x)), np.zeros(5))
    assert np.array_equal(ag.cov(f(x)), ag.kernelmatrix(ag.Matern32Kernel(), x))


def test_base_gp_const_mean(ag, rng):
    """src/base_gp.jl:26-36."""
    f = ag.GP(5.0, ag.Matern32Kernel())
    x = rng.standard_normal(5)
    assert np.array_equal(ag.mean(f(x)), 5.0 * np.ones(5))
    assert np.array_equal(ag.cov(f(x)), ag.kernelmatrix(ag.Matern32Kernel(), x))


def test_base_gp_custom_mean(ag, rng):
    """src/base_gp.jl:42-52."""
    f = ag.GP(lambda x: np.sin(x) + np.cos(x / 2), ag.Matern32Kernel())
    x = rng.standard_normal(5)
    np.testing.assert_allclose(np.ravel(ag.mean(f(x))), np.sin(x) + np.cos(x / 2), rtol=1e-14)
    assert np.array_equal(ag.cov(f(x)), ag.kernelmatrix(ag.Matern32Kernel(), x))


def test_finitegp_mean(ag, rng):
    """src/finite_gp_projection.jl:44-51."""
    f = ag.GP(ag.Matern52Kernel())
    x = rng.standard_normal(11)
    assert np.array_equal(ag.mean(f(x)), np.zeros(11))


def test_finitegp_cov_with_noise(ag, rng):
    """src/finite_gp_projection.jl:62-93: cov(f(x)), cov(f(x, 0.1)), cov(f(x, s))."""
    f = ag.GP(ag.Matern52Kernel())
    x = rng.standard_normal(11)
    K = ag.kernelmatrix(ag.Matern52Kernel(), x)
    assert np.array_equal(ag.cov(f(x)), K)
    np.testing.assert_allclose(ag.cov(f(x, 0.1)), K + 0.1 * np.eye(11), rtol=1e-14, atol=0)
    s = rng.random(11)
    np.testing.assert_allclose(ag.cov(f(x, s)), K + np.diag(s), rtol=1e-14, atol=0)


def test_finitegp_var_is_diag_cov(ag, rng):
    """src/finite_gp_projection.jl:107-112."""
    fx = ag.GP(ag.Matern52Kernel())(rng.standard_normal(10), 0.1)
    np.testing.assert_allclose(ag.var(fx), np.diag(ag.cov(fx)), rtol=1e-14)


def test_finitegp_mean_and_cov(ag):
    """src/finite_gp_projection.jl:126-131."""
    fx = ag.GP(ag.SqExponentialKernel())(np.linspace(-3.0, 3.0, 10), 0.1)
    m, c = ag.mean_and_cov(fx)
    assert np.array_equal(m, ag.mean(fx)) and np.array_equal(c, ag.cov(fx))


def test_finitegp_mean_and_var(ag):
    """src/finite_gp_projection.jl:147-152."""
    fx = ag.GP(ag.SqExponentialKernel())(np.linspace(-3.0, 3.0, 10), 0.1)
    m, v = ag.mean_and_var(fx)
    assert np.array_equal(m, ag.mean(fx)) and np.array_equal(v, ag.var(fx))


def test_finitegp_cross_cov(ag, rng):
    """src/finite_gp_projection.jl:166-175."""
    f = ag.GP(ag.Matern32Kernel())
    x1, x2 = rng.standard_normal(11), rng.standard_normal(13)
    c = ag.cov(f(x1), f(x2))
    assert c.shape == (11, 13)
    assert np.array_equal(c, ag.kernelmatrix(ag.Matern32Kernel(), x1, x2))


def test_finitegp_marginals(ag, rng):
    """src/finite_gp_projection.jl:189-201."""
    f = ag.GP(ag.Matern32Kernel())
    x = rng.standard_normal(11)
    fs = ag.marginals(f(x))
    assert np.array_equal(fs.mu, ag.mean(f(x)))
    np.testing.assert_allclose(fs.sigma, np.sqrt(np.diag(ag.cov(f(x)))), rtol=1e-14)


def test_finitegp_rand_shapes(ag, rng):
    """src/finite_gp_projection.jl:215-231 (default s2 = 1e-18, so the Matern32 Gram of 11 random points is factored
    with no jitter to speak of)."""
    f = ag.GP(ag.Matern32Kernel())
    x = rng.standard_normal(11)
    a = ag.rand(f(x))
    assert a.shape == (11,) and a.dtype == np.float64
    b = ag.rand(np.random.default_rng(123456), f(x))
    assert b.shape == (11,) and b.dtype == np.float64
    c = ag.rand(f(x), 3)
    assert c.shape == (11, 3) and c.dtype == np.float64
    d = ag.rand(np.random.default_rng(123456), f(x), 3)
    d2 = ag.rand(np.random.default_rng(123456), f(x), 3)
    assert d.shape == (11, 3) and np.array_equal(d, d2)  # same rng seed -> same draw


def test_finitegp_logpdf_types(ag, rng):
    """src/finite_gp_projection.jl:286-301."""
    f = ag.GP(ag.Matern32Kernel())
    x = rng.standard_normal(11)
    y = ag.rand(rng, f(x))
    lp = ag.logpdf(f(x), y)
    assert np.ndim(lp) == 0 and np.isfinite(lp)
    Y = ag.rand(rng, f(x), 3)
    lps = ag.logpdf(f(x), Y)
    assert np.shape(lps) == (3,) and np.all(np.isfinite(lps))
    # the matrix method is the vector method applied per column (:313-318)
    for j in range(3):
        np.testing.assert_allclose(lps[j], ag.logpdf(f(x), np.ascontiguousarray(Y[:, j])), rtol=RTOL)


def test_vfe_posterior_is_projectable(ag, rng):
    """src/sparse_approximations.jl:37-52."""
    f = ag.GP(ag.Matern52Kernel())
    x = rng.standard_normal(1000)
    z = np.linspace(-5.0, 5.0, 13)
    vfe = ag.VFE(f(z))
    y = ag.rand(rng, f(x, 0.1))
    post = ag.posterior(vfe, f(x, 0.1), y)
    pz = post(z)
    assert isinstance(pz, ag.FiniteGP)
    m, v = ag.mean_and_var(pz)
    assert m.shape == (13,) and np.all(np.isfinite(m)) and np.all(v > 0)


def test_elbo_below_logpdf(ag, rng):
    """src/sparse_approximations.jl:229-242."""
    f = ag.GP(ag.Matern52Kernel())
    x = rng.standard_normal(1000)
    z = np.linspace(-5.0, 5.0, 13)
    v = ag.VFE(f(z))
    y = ag.rand(rng, f(x, 0.1))
    assert ag.elbo(v, f(x, 0.1), y) < ag.logpdf(f(x, 0.1), y)


def test_dtc_evidence_close_to_logpdf(ag, rng):
    """src/sparse_approximations.jl:263-276: with 256 inducing points on [-5, 5] the DTC evidence equals the exact
    log marginal likelihood to atol = rtol = 1e-6.  (The reference's DTC objective is the second output of
    agp_vfe_elbo; z carries the reference's default jitter 1e-18, so the 256-point Matern52 K_zz is factored as is.)"""
    f = ag.GP(ag.Matern52Kernel())
    x = rng.standard_normal(1000)
    z = np.linspace(-5.0, 5.0, 256)
    d = ag.VFE(f(z))
    y = ag.rand(rng, f(x, 0.1))
    _, dtc = ag.approx_log_evidence(d, f(x, 0.1), y, return_dtc=True)
    exact = ag.logpdf(f(x, 0.1), y)
    assert abs(dtc - exact) <= 1e-6 + 1e-6 * abs(exact)
