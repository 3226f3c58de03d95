"""a FiniteGP over a PosteriorGP -- logpdf(f_post(x*, s2), y), rand(f_post(x*, s2)[, S]) --
through agp_post_logpdf / agp_post_rand, against the oracle; plus the reference's own TestUtils consistency
suite (/root/reference/src/util/TestUtils.jl:26-218) run on GP and PosteriorGP exactly as
/root/reference/test/base_gp.jl:4-14 and /root/reference/test/exact_gpr_posterior.jl:2-27 do."""
import numpy as np
import pytest

from oracle import agp_ref as ref
from test_gpu_parity import TOL, mk_kernel, problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,d,fam", [(300, 40, 3, ref.SE), (700, 257, 8, ref.MATERN52), (129, 128, 1, ref.MATERN32)])
def test_posterior_finitegp_logpdf_and_rand(ag, dtype, n, m, d, fam):
    ks, X, y = problem(n, d, fam, dtype, seed=11)
    rng = np.random.default_rng(5)
    Xs = rng.random((m, d)).astype(dtype)
    s2, s2s = 0.1, 0.05
    mean, noise, noise_s = ref.MeanSpec(1, 0.3), ref.NoiseSpec(0, s2), ref.NoiseSpec(0, s2s)
    post_ref = ref.posterior(ks, mean, noise, X, y)
    f = ag.GP(0.3, mk_kernel(ag, ks))
    p = ag.posterior(f(ag.RowVecs(X), s2), y)
    pfx = p(ag.RowVecs(Xs), s2s)

    Ys = np.asfortranarray(rng.standard_normal((m, 3)).astype(dtype))
    want = ref.post_logpdf(post_ref, Xs, noise_s, Ys)
    got = ag.logpdf(pfx, Ys)
    tol = TOL[dtype]
    np.testing.assert_allclose(got, want, rtol=tol["rtol"] * (1 if dtype == np.float64 else 10))
    np.testing.assert_allclose(ag.logpdf(pfx, np.ascontiguousarray(Ys[:, 1])), want[1],
                               rtol=tol["rtol"] * (1 if dtype == np.float64 else 10))

    Z = rng.standard_normal((m, 4)).astype(dtype)
    want_r = ref.post_rand_from_Z(post_ref, Xs, noise_s, Z)
    got_r = ag.rand_from_normals(pfx, Z)
    atol = 1e-8 if dtype == np.float64 else 2e-3
    np.testing.assert_allclose(got_r, want_r, rtol=0, atol=atol * max(1.0, np.abs(want_r).max()))


def test_posterior_logpdf_is_the_chain_rule_increment(ag):
    """log p(y1, y2) = log p(y1) + log p(y2 | y1): the posterior FiniteGP's logpdf is the increment (fp64)."""
    ks, X, y = problem(500, 4, ref.MATERN52, np.float64, seed=3)
    f = ag.GP(mk_kernel(ag, ks))
    n1 = 320
    lp_all = ag.logpdf(f(ag.RowVecs(X), 0.1), y)
    lp1, p1 = ag.fit(f(ag.RowVecs(X[:n1]), 0.1), y[:n1])
    lp2 = ag.logpdf(p1(ag.RowVecs(X[n1:]), 0.1), y[n1:])
    np.testing.assert_allclose(lp1 + lp2, lp_all, rtol=1e-10)


def test_sequential_conditioning_has_value_semantics(ag):
    """posterior(p1(x2, s2), y2) returns a NEW posterior (src/exact_gpr_posterior.jl:46-56 builds a new PosteriorGP);
    p1 must keep answering as before, and two different extensions of p1 must not see each other."""
    ks, X, y = problem(400, 3, ref.MATERN32, np.float64, seed=9)
    f = ag.GP(mk_kernel(ag, ks))
    n1 = 250
    Xs = np.random.default_rng(4).random((33, 3))
    p1 = ag.posterior(f(ag.RowVecs(X[:n1]), 0.1), y[:n1])
    m1, v1 = ag.mean_and_var(p1, ag.RowVecs(Xs))
    p2a = ag.posterior(p1(ag.RowVecs(X[n1:]), 0.1), y[n1:])
    p2b = ag.posterior(p1(ag.RowVecs(X[n1:330]), 0.2), y[n1:330])
    assert p1.data.C.n == n1 and p2a.data.C.n == 400 and p2b.data.C.n == 330
    m1b, v1b = ag.mean_and_var(p1, ag.RowVecs(Xs))
    assert np.array_equal(m1, m1b) and np.array_equal(v1, v1b)
    mean, z = ref.MeanSpec(), np.full
    ra = ref.posterior(ks, mean, ref.NoiseSpec(0, 0.1), X, y)
    rb = ref.posterior(ks, mean, ref.NoiseSpec(1, v=np.concatenate([z(n1, 0.1), z(80, 0.2)])), X[:330], y[:330])
    for p, r in ((p2a, ra), (p2b, rb)):
        m, v = ag.mean_and_var(p, ag.RowVecs(Xs))
        mr, vr = ref.post_mean_and_var(r, Xs)
        np.testing.assert_allclose(m, mr, rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(v, vr, rtol=1e-6, atol=1e-8)


def test_posterior_finitegp_not_posdef(ag):
    """an indefinite C* + Sigma* (negative sigma^2, as a caller's bug would produce): the reference's cholesky throws
    PosDefException at src/finite_gp_projection.jl:308 / :235; so must the device path, for logpdf and rand."""
    ks, X, y = problem(64, 2, ref.SE, np.float64, seed=1)
    f = ag.GP(mk_kernel(ag, ks))
    p = ag.posterior(f(ag.RowVecs(X), 0.1), y)
    Xs = np.random.default_rng(2).random((8, 2))
    with pytest.raises(ag.PosDefException):
        ag.logpdf(p(ag.RowVecs(Xs), -0.5), np.zeros(8))
    with pytest.raises(ag.PosDefException):
        ag.rand(p(ag.RowVecs(Xs), -0.5))
    assert np.isfinite(ag.logpdf(p(ag.RowVecs(Xs), 0.5), np.zeros(8)))  # the context stays usable


# ---- /root/reference/src/util/TestUtils.jl, line by line -------------------------------------------------
def approx(a, b, rtol=1.5e-8):  # Julia's isapprox default for Float64: rtol = sqrt(eps), norm-wise
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def finitegp_primary_public_interface(ag, rng, fx, atol=1e-12, conditioning=True):
    """TestUtils.jl:26-73."""
    y = ag.rand(rng, fx)
    assert y.ndim == 1 and len(y) == len(fx)
    y = ag.rand(fx)
    assert y.ndim == 1 and len(y) == len(fx)
    Y = ag.rand(rng, fx, 3)
    assert Y.shape == (len(fx), 3)
    Y = ag.rand(fx, 3)
    assert Y.shape == (len(fx), 3)
    ms = ag.marginals(fx)
    assert len(ms.mu) == len(fx)
    assert approx(ag.mean(fx), ms.mu)
    assert approx(ag.var(fx), ms.sigma ** 2)
    assert approx(ag.mean_and_var(fx)[0], ag.mean(fx))
    assert approx(ag.mean_and_var(fx)[1], ag.var(fx))
    assert np.all(ag.var(fx) > -atol)
    assert np.ndim(ag.logpdf(fx, y)) == 0
    if conditioning:
        assert isinstance(ag.posterior(fx, y), ag.AbstractGP)


def finitegp_primary_and_secondary_public_interface(ag, rng, fx, atol=1e-12, conditioning=True):
    """TestUtils.jl:89-108."""
    finitegp_primary_public_interface(ag, rng, fx, atol, conditioning)
    assert approx(np.diag(ag.cov(fx)), ag.var(fx))
    m, C = ag.mean_and_cov(fx)
    assert approx(m, ag.mean(fx)) and approx(C, ag.cov(fx))
    assert np.linalg.eigvalsh(ag.cov(fx)).min() > -atol
    assert approx(ag.cov(fx), ag.cov(fx).T)


def internal_abstractgps_interface(ag, rng, f, x, z, atol=1e-12, s2=1e-1, jitter=1e-18, vfe=True, conditioning=True):
    """TestUtils.jl:134-218."""
    assert len(x) != len(z)
    m = ag.mean(f, x)
    assert m.shape == (len(x),)
    C_xy = ag.cov(f, x, z)
    assert C_xy.shape == (len(x), len(z))
    assert approx(C_xy, ag.cov(f, z, x).T)
    C_xx = ag.cov(f, x)
    assert C_xx.shape == (len(x), len(x))
    assert np.linalg.eigvalsh(C_xx).min() > -atol
    assert approx(C_xx, ag.cov(f, x, x))
    C_xx_diag = ag.var(f, x)
    assert C_xx_diag.shape == (len(x),)
    assert approx(C_xx_diag, np.diag(C_xx))
    m2, C2 = ag.mean_and_cov(f, x)
    assert approx(m2, ag.mean(f, x)) and approx(C2, ag.cov(f, x))
    m3, c3 = ag.mean_and_var(f, x)
    assert approx(m3, ag.mean(f, x)) and approx(c3, ag.var(f, x))
    finitegp_primary_and_secondary_public_interface(ag, rng, f(x, s2), atol, conditioning)
    fx, fz = f(x, s2), f(z, s2)
    Sy = np.diag(np.full(len(x), s2))
    assert approx(ag.mean(fx), ag.mean(f, x))
    assert approx(ag.cov(fx), ag.cov(f, x) + Sy)
    assert approx(ag.cov(fx, fz), ag.cov(f, x, z))
    assert approx(ag.marginals(fx).mu, ag.mean(f, x))
    assert approx(ag.marginals(fx).sigma ** 2, ag.var(f, x) + np.diag(Sy))
    y = ag.rand(fx)
    assert len(y) == len(x)
    lp = ag.logpdf(fx, y)
    assert np.ndim(lp) == 0
    if vfe:  # TestUtils.jl:212-217 (VFE over a PosteriorGP prior is outside the device path: GP priors only)
        el = ag.elbo(ag.VFE(f(x, jitter)), fx, y)
        assert abs(el - lp) <= max(1e-5, 1e-5 * max(abs(el), abs(lp)))
        assert ag.elbo(ag.VFE(f(z, jitter)), fx, y) <= lp


def test_reference_base_gp_suite(ag):
    """/root/reference/test/base_gp.jl:4-14."""
    rng = np.random.default_rng(123456)
    f = ag.GP(np.sin, ag.Matern32Kernel())
    x, xp = np.linspace(-1.0, 1.0, 5), np.linspace(-1.0, 1.0, 6)
    np.testing.assert_allclose(ag.mean(f, x), np.sin(x), rtol=1e-15)  # mean_vector(m, x), test/base_gp.jl:11
    assert np.array_equal(ag.cov(f, x), ag.kernelmatrix(ag.Matern32Kernel(), x))
    internal_abstractgps_interface(ag, rng, f, x, xp)


def test_reference_exact_gpr_posterior_suite(ag):
    """/root/reference/test/exact_gpr_posterior.jl:2-27."""
    rng = np.random.default_rng(123456)
    f = ag.GP(np.sin, ag.Matern32Kernel())
    x = np.linspace(-1.0, 1.0, 3)
    fx = f(x, 1e-15)
    y = ag.rand(rng, fx)
    f_post = ag.posterior(fx, y)
    assert approx(ag.mean(f_post, x), y)
    np.testing.assert_allclose(ag.var(f_post, x), np.zeros(3), rtol=1e-14, atol=1e-13)
    a = np.linspace(-1.0, 1.0, 5)
    b = rng.standard_normal(6)
    internal_abstractgps_interface(ag, rng, f_post, a, b, vfe=False)
