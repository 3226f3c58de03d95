"""The oracle (oracle/agp_ref.py) cannot be pinned against the reference itself here (no Julia, no golden vectors in the
reference -- SURVEY s8c).  These tests pin it against two implementations that share NO code with it:

* scikit-learn's GaussianProcessRegressor / kernels (RBF, Matern nu = 1/2, 3/2, 5/2, DotProduct; fixed hyper-parameters):
  kernel matrices, log marginal likelihood (= logpdf(fx, y), /root/reference/src/finite_gp_projection.jl:306-311) and the
  predictive mean / standard deviation (= mean_and_var(f_post(x*)), /root/reference/src/exact_gpr_posterior.jl:85-90);
* 60-digit mpmath arithmetic (Gaussian elimination, no LAPACK) at N = 12 for logpdf, the posterior weights alpha and the
  Titsias bound (/root/reference/src/sparse_approximations.jl:248-254, 289-305).
"""
import numpy as np
import pytest

from oracle import agp_ref as ref


def _sk_kernel(fam, ell, var, c):
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, DotProduct, Matern
    if fam == ref.SE:
        k = RBF(length_scale=ell)
    elif fam == ref.LINEAR:
        return ConstantKernel(var) * DotProduct(sigma_0=np.sqrt(c))   # x'y + sigma_0^2
    else:
        k = Matern(length_scale=ell, nu={ref.MATERN12: 0.5, ref.MATERN32: 1.5, ref.MATERN52: 2.5}[fam])
    return ConstantKernel(var) * k


@pytest.mark.parametrize("fam", [ref.SE, ref.MATERN12, ref.MATERN32, ref.MATERN52, ref.LINEAR])
@pytest.mark.parametrize("d", [1, 5])
def test_oracle_matches_scikit_learn(fam, d):
    from sklearn.gaussian_process import GaussianProcessRegressor
    rng = np.random.default_rng(10 * fam + d)
    n, m = 60, 25
    X, Xs = rng.random((n, d)), rng.random((m, d))
    y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(n)
    ell, var, c, s2 = 0.7, 1.3, 0.4, 0.05
    if fam == ref.LINEAR:
        ks = ref.KernelSpec(ref.LINEAR, var, linear_c=c)
    else:
        ks = ref.KernelSpec(fam, var, ref.T_SCALE, scale=1.0 / ell)     # with_lengthscale(k, ell) = k o ScaleTransform(1/ell)
    sk = _sk_kernel(fam, ell, var, c)
    assert np.allclose(ref.kernelmatrix(ks, X), sk(X), rtol=1e-12, atol=1e-13)
    assert np.allclose(ref.kernelmatrix(ks, X, Xs), sk(X, Xs), rtol=1e-12, atol=1e-13)
    gpr = GaussianProcessRegressor(kernel=sk, alpha=s2, optimizer=None, normalize_y=False).fit(X, y)
    lml = gpr.log_marginal_likelihood(gpr.kernel_.theta)
    lp = ref.logpdf(ks, ref.MeanSpec(), ref.NoiseSpec(0, s2), X, y)
    assert abs(lp - lml) <= 1e-10 * abs(lml)
    mu_sk, sd_sk = gpr.predict(Xs, return_std=True)
    post = ref.posterior(ks, ref.MeanSpec(), ref.NoiseSpec(0, s2), X, y)
    mu, v = ref.post_mean_and_var(post, Xs)          # latent f*: no observation noise, as sklearn's predict
    assert np.allclose(mu, mu_sk, rtol=1e-9, atol=1e-10)
    assert np.allclose(np.sqrt(np.maximum(v, 0)), sd_sk, rtol=1e-6, atol=1e-7)
    assert np.allclose(post["alpha"], gpr.alpha_.ravel(), rtol=1e-8, atol=1e-9)


def _mp_solve(A, B):
    import mpmath as mp
    if B.cols == 1:
        return mp.lu_solve(A, B)
    return mp.inverse(A) * B   # 60 digits: the explicit inverse loses nothing that matters at these sizes


def test_oracle_matches_60_digit_arithmetic():
    import mpmath as mp
    mp.mp.dps = 60
    rng = np.random.default_rng(3)
    n, m, d = 12, 5, 2
    X, Z = rng.random((n, d)), rng.random((m, d))
    y = np.sin(2 * X[:, 0]) + 0.2 * rng.standard_normal(n)
    var, s, s2, jit, c0 = 1.7, 1.9, 0.03, 1e-6, 0.25
    ks = ref.KernelSpec(ref.MATERN52, var, ref.T_SCALE, scale=s)

    def kfun(a, b):  # Matern-5/2, sigma_f^2 * (1 + sqrt5 r + 5 r^2 / 3) exp(-sqrt5 r), r = s |a - b|
        r = mp.sqrt(sum((mp.mpf(float(u)) - mp.mpf(float(v))) ** 2 for u, v in zip(a, b))) * mp.mpf(s)
        q = mp.sqrt(5) * r
        return mp.mpf(var) * (1 + q + q * q / 3) * mp.e ** (-q)
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = kfun(X[i], X[j]) + (mp.mpf(s2) if i == j else 0)
    dlt = mp.matrix([mp.mpf(float(v)) - mp.mpf(c0) for v in y])
    alpha = _mp_solve(K, dlt)
    logdet = mp.log(mp.det(K))
    lp_mp = -(n * mp.log(2 * mp.pi) + logdet + (dlt.T * alpha)[0, 0]) / 2
    lp = ref.logpdf(ks, ref.MeanSpec(1, c0), ref.NoiseSpec(0, s2), X, y)
    assert abs(lp - float(lp_mp)) <= 1e-11 * abs(float(lp_mp))
    post = ref.posterior(ks, ref.MeanSpec(1, c0), ref.NoiseSpec(0, s2), X, y)
    assert np.allclose(post["alpha"], [float(a) for a in alpha], rtol=1e-9, atol=1e-10)
    # Titsias bound: log N(y | m, Qff + s2 I) - tr(Kff - Qff) / (2 s2), Qff = Kfu (Kuu + jit I)^-1 Kuf
    Kuu = mp.matrix(m, m)
    Kuf = mp.matrix(m, n)
    for i in range(m):
        for j in range(m):
            Kuu[i, j] = kfun(Z[i], Z[j]) + (mp.mpf(jit) if i == j else 0)
        for j in range(n):
            Kuf[i, j] = kfun(Z[i], X[j])
    Q = Kuf.T * _mp_solve(Kuu, Kuf)
    C = Q + mp.mpf(s2) * mp.eye(n)
    dtc_mp = -(n * mp.log(2 * mp.pi) + mp.log(mp.det(C)) + (dlt.T * _mp_solve(C, dlt))[0, 0]) / 2
    tr = sum(mp.mpf(var) - Q[i, i] for i in range(n))
    elbo_mp = dtc_mp - tr / (2 * mp.mpf(s2))
    el = ref.elbo(ks, ref.MeanSpec(1, c0), ref.NoiseSpec(0, s2), X, y, Z, ref.NoiseSpec(0, jit))
    dt = ref.dtc(ks, ref.MeanSpec(1, c0), ref.NoiseSpec(0, s2), X, y, Z, ref.NoiseSpec(0, jit))
    assert abs(dt - float(dtc_mp)) <= 1e-8 * abs(float(dtc_mp))
    assert abs(el - float(elbo_mp)) <= 1e-8 * abs(float(elbo_mp))
