"""T0: the oracle pinned against independent implementations and the relations the reference's own
tests assert.  CPU only."""
import numpy as np
import pytest
import scipy.stats as st

from oracle import agp_ref as ref

FAMS = [ref.SE, ref.MATERN12, ref.MATERN32, ref.MATERN52, ref.LINEAR]


def _problem(n=37, d=3, seed=0, dtype=np.float64, fam=ref.SE):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(dtype)
    y = rng.standard_normal(n).astype(dtype)
    k = ref.KernelSpec(fam, 1.3, ref.T_SCALE, scale=0.7, linear_c=0.5)
    return k, ref.MeanSpec(1, 0.3), ref.NoiseSpec(0, 0.1), X, y


@pytest.mark.parametrize("fam", FAMS)
def test_logpdf_matches_scipy_mvn(fam):
    # mirror of test/finite_gp_projection.jl:143 (logpdf vs Distributions.MvNormal)
    k, mean, noise, X, y = _problem(fam=fam)
    m, C = ref.mean_and_cov_fx(k, mean, noise, X)
    want = st.multivariate_normal(m, C).logpdf(y)
    got = ref.logpdf(k, mean, noise, X, y)
    assert np.isclose(got, want, rtol=1e-10, atol=1e-10)
    Y = np.stack([y, 2 * y, y - 1], 1)
    got = ref.logpdf(k, mean, noise, X, Y)  # multi-column consistency :147-150
    for s in range(3):
        assert np.isclose(got[s], st.multivariate_normal(m, C).logpdf(Y[:, s]), rtol=1e-10)


@pytest.mark.parametrize("fam", FAMS)
def test_kernel_closed_forms(fam):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((9, 4))
    v = rng.random(4) + 0.5
    k = ref.KernelSpec(fam, 2.0, ref.T_ARD, ard=v, linear_c=0.25)
    K = ref.kernelmatrix(k, X)
    for i in range(9):
        for j in range(9):
            a, b = X[i] * v, X[j] * v
            d = np.linalg.norm(a - b)
            want = {ref.SE: np.exp(-d * d / 2), ref.MATERN12: np.exp(-d),
                    ref.MATERN32: (1 + np.sqrt(3) * d) * np.exp(-np.sqrt(3) * d),
                    ref.MATERN52: (1 + np.sqrt(5) * d + 5 * d * d / 3) * np.exp(-np.sqrt(5) * d),
                    ref.LINEAR: a @ b + 0.25}[fam]
            assert np.isclose(K[i, j], 2.0 * want, rtol=1e-12, atol=1e-14)
    assert np.allclose(np.diag(K), ref.kernelmatrix_diag(k, X), rtol=1e-13)
    assert np.allclose(K, ref.kernelmatrix(k, X, method="gemm"), rtol=1e-10, atol=1e-12)
    Z = rng.standard_normal((5, 4))
    assert np.allclose(ref.kernelmatrix(k, X, Z), ref.kernelmatrix(k, Z, X).T)  # TestUtils.jl:157


def test_operator_identities():
    # test/util/common_covmat_ops.jl:40-105
    rng = np.random.default_rng(123456)
    B = rng.standard_normal((5, 5))
    A = B.T @ B + 1e-6 * np.eye(5)
    U = ref.cholesky_upper(A)
    X, Y = rng.standard_normal((5, 3)), rng.standard_normal((5, 4))
    assert np.allclose(ref.Xt_invA_X(U, X), X.T @ np.linalg.solve(A, X))
    assert np.allclose(ref.Xt_invA_Y(X, U, Y), X.T @ np.linalg.solve(A, Y))
    assert np.allclose(ref.diag_Xt_invA_X(U, X), np.diag(X.T @ np.linalg.solve(A, X)))
    assert np.isclose(ref.tr_Xt_invA_X(U, X), np.trace(X.T @ np.linalg.solve(A, X)))
    # update_chol vs full cholesky, atol 1e-5 (:21-37)
    Bb = rng.standard_normal((8, 8))
    Cc = Bb.T @ Bb + 1e-3 * np.eye(8)
    U11 = ref.cholesky_upper(Cc[:5, :5])
    Uu = ref.update_chol(U11, Cc[:5, 5:], Cc[5:, 5:])
    assert np.allclose(Uu, ref.cholesky_upper(Cc), atol=1e-5)


def test_posterior_collapses_on_data():
    # test/exact_gpr_posterior.jl:21-22
    rng = np.random.default_rng(3)
    X = rng.standard_normal((11, 2))
    y = rng.standard_normal(11)
    k = ref.KernelSpec(ref.SE)
    post = ref.posterior(k, ref.MeanSpec(), ref.NoiseSpec(0, 1e-15), X, y)
    m, v = ref.post_mean_and_var(post, X)
    assert np.allclose(m, y, atol=1e-7) and np.allclose(v, 0, atol=1e-7)
    mc, Cc = ref.post_mean_and_cov(post, X)
    assert np.allclose(np.diag(Cc), v, atol=1e-10)


def test_sequential_equals_batch():
    # test/exact_gpr_posterior.jl:29-43
    k, mean, noise, X, y = _problem(n=30)
    p1 = ref.posterior(k, mean, noise, X[:18], y[:18])
    p2 = ref.posterior_sequential(p1, noise, X[18:], y[18:])
    pb = ref.posterior(k, mean, noise, X, y)
    assert np.allclose(p2["U"], pb["U"], atol=1e-5)
    assert np.allclose(p2["alpha"], pb["alpha"], atol=1e-5)
    assert np.allclose(p2["delta"], pb["delta"])


def test_rand_statistics():
    # test/finite_gp_projection.jl:84-104 (statistical; S large, tolerance 1e-2 scale)
    k, mean, noise, X, _ = _problem(n=6)
    Z = np.random.default_rng(5).standard_normal((6, 200000))
    S = ref.rand_from_Z(k, mean, noise, X, Z)
    m, C = ref.mean_and_cov_fx(k, mean, noise, X)
    assert np.allclose(S.mean(1), m, atol=2e-2)
    assert np.allclose(np.cov(S), C, atol=3e-2)


def test_vfe_relations():
    # test/sparse_approximations.jl:24-25,94,99 ; src/util/TestUtils.jl:213-217
    k, mean, noise, X, y = _problem(n=40, d=2)
    jit = ref.NoiseSpec(0, 1e-12)
    lp = ref.logpdf(k, mean, noise, X, y)
    assert np.isclose(ref.elbo(k, mean, noise, X, y, X, jit), lp, rtol=1e-5, atol=1e-5)
    Z = X[:10] + 0.05
    assert ref.elbo(k, mean, noise, X, y, Z, jit) < lp
    assert ref.elbo(k, mean, noise, X, y, Z, jit) <= ref.dtc(k, mean, noise, X, y, Z, jit)
    vp = ref.vfe_posterior(k, mean, noise, X, y, X, jit)
    ep = ref.posterior(k, mean, noise, X, y)
    Xs = np.random.default_rng(9).standard_normal((7, 2))
    m1, v1 = ref.vfe_mean_and_var(vp, Xs)
    m2, v2 = ref.post_mean_and_var(ep, Xs)
    assert np.allclose(m1, m2, atol=1e-6) and np.allclose(v1, v2, atol=1e-6)
    # tr_Cf_invSigma_y vs tr(Cf / Sigma_y) (test/sparse_approximations.jl:121-135)
    nv = ref.NoiseSpec(1, v=np.linspace(0.1, 0.3, 40))
    Cf = ref.kernelmatrix(k, X)
    assert np.isclose(ref.tr_Cf_invSy(k, nv, X), np.trace(Cf @ np.diag(1 / nv.v)))
    with pytest.raises(ValueError):
        ref.elbo(k, mean, noise, X, y[:-1], Z, jit)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_type_stability(dtype):
    # test/finite_gp_projection.jl:180-191, test/sparse_approximations.jl:103-118
    k, mean, noise, X, y = _problem(n=20, dtype=dtype)
    assert ref.logpdf(k, mean, noise, X, y).dtype == dtype
    assert ref.posterior(k, mean, noise, X, y)["alpha"].dtype == dtype
    assert ref.elbo(k, mean, noise, X, y, X[:5], ref.NoiseSpec(0, 1e-3)).dtype == dtype


def test_not_posdef_raises():
    X = np.zeros((4, 1))
    with pytest.raises(np.linalg.LinAlgError):
        ref.logpdf(ref.KernelSpec(ref.SE), ref.MeanSpec(), ref.NoiseSpec(0, -1.0), X, np.zeros(4))


def test_c1_readme_toy_and_golden():
    """config C1 (README.md:31-50) on CPU, pinned to the committed fixture (tests/golden/make_golden.py)."""
    import os
    cfg = ref.make_config("C1")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1.npz"))
    assert np.array_equal(g["X"], cfg["X"]) and np.array_equal(g["y"], cfg["y"])
    lp = ref.logpdf(cfg["k"], cfg["mean"], cfg["noise"], cfg["X"], cfg["y"])
    post = ref.posterior(cfg["k"], cfg["mean"], cfg["noise"], cfg["X"], cfg["y"])
    assert np.isclose(lp, g["logpdf"], rtol=1e-12)
    assert np.allclose(post["alpha"], g["alpha"], rtol=1e-9)
    m, v = ref.post_mean_and_var(post, g["Xs"])
    assert np.allclose(m, g["mean_s"], rtol=1e-9) and np.allclose(v, g["var_s"], rtol=1e-7, atol=1e-12)


# ---- SURVEY s8(f) rank 1 (next row): the gradient oracle, pinned by central finite differences of the oracle's own
# logpdf, as the reference pins its AD rules with FiniteDifferences (test/finite_gp_projection.jl:152-178)
@pytest.mark.parametrize("fam", [ref.SE, ref.MATERN12, ref.MATERN32, ref.MATERN52, ref.LINEAR])
@pytest.mark.parametrize("transform", [ref.T_SCALE, ref.T_ARD])
def test_logpdf_grad_matches_finite_differences(fam, transform):
    import copy
    rng = np.random.default_rng(7)
    n, d = 40, 3
    X = rng.random((n, d))
    y = np.sin(3 * X[:, 0]) + 0.2 * rng.standard_normal(n)
    k = ref.KernelSpec(fam, 1.3, transform, scale=1.7, ard=np.array([1.2, 0.7, 2.1]), linear_c=0.4)
    mean = ref.MeanSpec(1, 0.25)
    noise = ref.NoiseSpec(1, v=0.05 + 0.1 * rng.random(n))
    g = ref.logpdf_grad(k, mean, noise, X, y)

    def fd(setter, h=1e-6):
        kp, mp, npz = copy.deepcopy(k), copy.deepcopy(mean), copy.deepcopy(noise)
        setter(kp, mp, npz, +h)
        up = ref.logpdf(kp, mp, npz, X, y)
        kp, mp, npz = copy.deepcopy(k), copy.deepcopy(mean), copy.deepcopy(noise)
        setter(kp, mp, npz, -h)
        return (up - ref.logpdf(kp, mp, npz, X, y)) / (2 * h)

    def chk(got, want):
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (got, want)

    chk(g["variance"], fd(lambda kp, mp, npz, h: setattr(kp, "variance", kp.variance + h)))
    chk(g["mean_c"], fd(lambda kp, mp, npz, h: setattr(mp, "c", mp.c + h)))
    if transform == ref.T_SCALE:
        chk(g["scale"], fd(lambda kp, mp, npz, h: setattr(kp, "scale", kp.scale + h)))
    else:
        for j in range(d):
            def bump(kp, mp, npz, h, j=j):
                a = np.array(kp.ard, dtype=np.float64)
                a[j] += h
                kp.ard = a
            chk(g["ard"][j], fd(bump))
    if fam == ref.LINEAR:
        chk(g["linear_c"], fd(lambda kp, mp, npz, h: setattr(kp, "linear_c", kp.linear_c + h)))
    for j in (0, 17):
        def bumpn(kp, mp, npz, h, j=j):
            v = np.array(npz.v, dtype=np.float64)
            v[j] += h
            npz.v = v
        chk(g["noise"][j], fd(bumpn))
    # scalar noise: the trace
    ns = ref.NoiseSpec(0, 0.1)
    gs = ref.logpdf_grad(k, mean, ns, X, y)
    h = 1e-6
    want = (ref.logpdf(k, mean, ref.NoiseSpec(0, 0.1 + h), X, y) - ref.logpdf(k, mean, ref.NoiseSpec(0, 0.1 - h), X, y)) / (2 * h)
    chk(gs["noise"], want)


# ---- the two distance formulations: "direct" (differences; the CUDA kernel and the parity oracle) vs "gemm"
# (||a||^2 + ||b||^2 - 2 a.b clamped at 0: the Distances.jl pairwise form KernelFunctions.kernelmatrix executes in the
# reference, call sites src/base_gp.jl:70,74).  The GPU Gram is compared with "direct"; this bounds what that choice can
# hide: the two agree to cancellation error of the gemm form, and the downstream logpdf agrees far inside rtol 1e-8.
@pytest.mark.parametrize("fam", [ref.SE, ref.MATERN12, ref.MATERN32, ref.MATERN52])
@pytest.mark.parametrize("d", [1, 8, 64])
def test_gemm_and_direct_distance_forms_agree(fam, d):
    rng = np.random.default_rng(d * 10 + fam)
    n = 400
    X = rng.random((n, d))
    k = ref.KernelSpec(fam, 1.3, ref.T_SCALE, scale=1.0 / (0.5 * np.sqrt(d)))
    Kd = ref.kernelmatrix(k, X, method="direct")
    Kg = ref.kernelmatrix(k, X, method="gemm")
    # d^2 error of the gemm form ~ eps * (||a||^2 + ||b||^2) <= eps * 8; kappa is 1-Lipschitz in d^2 for SE and
    # ~ sqrt-amplified near d = 0 for the Matern family (|d kappa| <= c |d(d)|, d(d) <= sqrt(d(d^2)))
    eps = np.finfo(np.float64).eps
    tol = 1.3 * (8 * 8 * eps if fam == ref.SE else 3.0 * np.sqrt(8 * 8 * eps))
    assert np.max(np.abs(Kd - Kg)) <= tol
    assert np.array_equal(np.diag(Kd), np.diag(Kg))  # exactly-zero self distance in both
    y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(n)
    old = ref.DEFAULT_METHOD
    try:
        ref.DEFAULT_METHOD = "direct"
        lp_d = ref.logpdf(k, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y)
        ref.DEFAULT_METHOD = "gemm"
        lp_g = ref.logpdf(k, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y)
    finally:
        ref.DEFAULT_METHOD = old
    assert abs(lp_d - lp_g) <= (1e-10 if fam == ref.SE else 2e-8) * abs(lp_d), (lp_d, lp_g)
