"""T1/T2/T4: the CUDA path (through the C ABI) against the oracle on identical seeded inputs.
Tolerances are BASELINE.json's: rtol 1e-8 (fp64), 1e-4 (fp32) on logpdf; element-wise outputs
carry an atol scaled to the data."""
import os

import numpy as np
import pytest

from oracle import agp_ref as ref

pytestmark = pytest.mark.gpu

FAM_CTOR = {ref.SE: "SqExponentialKernel", ref.MATERN12: "Matern12Kernel", ref.MATERN32: "Matern32Kernel",
            ref.MATERN52: "Matern52Kernel", ref.LINEAR: "LinearKernel"}


def mk_kernel(ag, ks: ref.KernelSpec):
    k = getattr(ag, FAM_CTOR[ks.family])() if ks.family != ref.LINEAR else ag.LinearKernel(c=ks.linear_c)
    if ks.transform == ref.T_SCALE:
        k = k.compose(ag.ScaleTransform(ks.scale))
    elif ks.transform == ref.T_ARD:
        k = k.compose(ag.ARDTransform(ks.ard))
    return ks.variance * k


def problem(n, d, fam, dtype, seed=0, transform=ref.T_SCALE):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d)).astype(dtype)
    y = (np.sin(2 * np.pi * X.mean(1)) + 0.3 * rng.standard_normal(n)).astype(dtype)
    if transform == ref.T_ARD:
        ks = ref.KernelSpec(fam, 1.7, ref.T_ARD, ard=(1.0 + rng.random(d)).astype(dtype), linear_c=0.3)
    elif transform == ref.T_SCALE:
        ks = ref.KernelSpec(fam, 1.7, ref.T_SCALE, scale=1.0 / (0.5 * np.sqrt(d)), linear_c=0.3)
    else:
        ks = ref.KernelSpec(fam, 1.7, linear_c=0.3)
    return ks, X, y


TOL = {np.float64: dict(rtol=1e-8, atol=1e-9), np.float32: dict(rtol=1e-4, atol=2e-4)}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fam", list(FAM_CTOR))
@pytest.mark.parametrize("n,m,d", [(1, 1, 1), (63, 70, 3), (129, 5, 8), (300, 257, 33)])
def test_gram(ag, fam, dtype, n, m, d):
    ks, X, _ = problem(n, d, fam, dtype, transform=ref.T_ARD if d > 1 else ref.T_SCALE)
    Z = np.random.default_rng(5).random((m, d)).astype(dtype)
    f = ag.GP(mk_kernel(ag, ks))
    K = ag.cov(f, ag.RowVecs(X))
    Kref = ref.kernelmatrix(ks, X)
    tol = dict(rtol=1e-12, atol=1e-13) if dtype == np.float64 else dict(rtol=2e-5, atol=2e-6)
    if fam == ref.MATERN12 or fam == ref.MATERN32 or fam == ref.MATERN52:
        tol["atol"] = max(tol["atol"], 1e-7 if dtype == np.float64 else 2e-3)  # sqrt near d=0 amplifies rounding
    assert np.allclose(K, Kref, **tol)
    assert np.allclose(ag.cov(f, ag.RowVecs(X), ag.RowVecs(Z)), ref.kernelmatrix(ks, X, Z), **tol)
    Kn = ag.cov(f(ag.ColVecs(X.T.copy()), 0.25))  # cov(fx) = K + Sigma_y, ColVecs layout
    assert np.allclose(Kn, Kref + 0.25 * np.eye(n, dtype=dtype), **tol)
    assert np.allclose(np.diag(K), ag.var(f, ag.RowVecs(X)), **tol)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,d,fam", [(10, 1, ref.MATERN32), (127, 2, ref.SE), (128, 4, ref.MATERN52), (129, 3, ref.SE),
                                     (700, 8, ref.SE), (1000, 5, ref.MATERN12), (1537, 16, ref.LINEAR)])
def test_logpdf_posterior(ag, dtype, n, d, fam):
    ks, X, y = problem(n, d, fam, dtype, seed=n)
    noise = ref.NoiseSpec(0, 0.1)
    if fam == ref.LINEAR:
        noise = ref.NoiseSpec(0, 0.5)
    mean = ref.MeanSpec(1, 0.25)
    f = ag.GP(0.25, mk_kernel(ag, ks))
    fx = f(ag.RowVecs(X), noise.s)
    lp, post = ag.fit(fx, y)
    lp_ref = ref.logpdf(ks, mean, noise, X, y)
    pr = ref.posterior(ks, mean, noise, X, y)
    assert lp.dtype == dtype  # type stability (test/finite_gp_projection.jl:180-191)
    rt = TOL[dtype]["rtol"]
    assert abs(lp - lp_ref) <= rt * abs(lp_ref) + (0 if dtype == np.float64 else 1e-3), (lp, lp_ref)
    assert np.isclose(ag.logpdf(fx, y), lp, rtol=1e-12)  # logpdf-only path == fused path
    scale = np.abs(pr["alpha"]).max()
    atol = (1e-7 if dtype == np.float64 else 5e-3) * scale
    assert np.allclose(post.data.alpha, pr["alpha"], rtol=1e-6 if dtype == np.float64 else 1e-2, atol=atol)
    assert np.allclose(post.data.delta, pr["delta"])
    # logdet and the exported factor
    assert np.isclose(post.data.C.logdet(), ref.logdet_chol(pr["U"]), rtol=1e-9 if dtype == np.float64 else 1e-4)
    if n <= 700:
        U = post.data.C.U
        assert np.allclose(np.tril(U, -1), 0)
        assert np.allclose(U, pr["U"], rtol=1e-7 if dtype == np.float64 else 1e-2, atol=1e-9 if dtype == np.float64 else 2e-3)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_multicolumn_and_noise_and_mean_variants(ag, dtype):
    n, d = 333, 4
    ks, X, y = problem(n, d, ref.SE, dtype, seed=11)
    rng = np.random.default_rng(2)
    Y = np.stack([y, 2 * y + 1, rng.standard_normal(n).astype(dtype)], 1)
    nv = (0.05 + 0.1 * rng.random(n)).astype(dtype)  # per-point noise
    mv = np.cos(X[:, 0]).astype(dtype)
    f = ag.GP(lambda r: np.cos(r[0]), mk_kernel(ag, ks))  # CustomMean -> host-evaluated vector
    lp = ag.logpdf(f(ag.RowVecs(X), nv), Y)
    lp_ref = ref.logpdf(ks, ref.MeanSpec(2, v=mv), ref.NoiseSpec(1, v=nv), X, Y)
    assert lp.shape == (3,)
    assert np.allclose(lp, lp_ref, rtol=TOL[dtype]["rtol"], atol=0 if dtype == np.float64 else 1e-2)
    assert np.isclose(ag.loglikelihood(f(ag.RowVecs(X), nv), Y), lp_ref.sum(), rtol=TOL[dtype]["rtol"] * 10)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,d,fam", [(10, 21, 1, ref.MATERN32), (200, 77, 3, ref.SE), (515, 300, 8, ref.MATERN32)])
def test_mean_and_var_and_cov(ag, dtype, n, m, d, fam):
    ks, X, y = problem(n, d, fam, dtype, seed=3, transform=ref.T_ARD)
    noise = ref.NoiseSpec(0, 0.05)
    Xs = np.random.default_rng(8).random((m, d)).astype(dtype)
    f = ag.GP(mk_kernel(ag, ks))
    post = ag.posterior(f(ag.RowVecs(X), 0.05), y)
    pr = ref.posterior(ks, ref.MeanSpec(), noise, X, y)
    tol = dict(rtol=1e-7, atol=1e-8) if dtype == np.float64 else dict(rtol=2e-3, atol=2e-3)
    mu, v = ag.mean_and_var(post(ag.RowVecs(Xs), 0.05))
    mu_r, v_r = ref.post_mean_and_var(pr, Xs, noise_s=noise)
    assert np.allclose(mu, mu_r, **tol) and np.allclose(v, v_r, **tol)
    mu2, v2 = ag.mean_and_var(post, ag.RowVecs(Xs))  # noise-free: f_post itself
    assert np.allclose(v2, v_r - 0.05, **tol) and np.allclose(mu2, mu_r, **tol)
    assert np.all(v2 > -1e-6)  # var > -atol (src/util/TestUtils.jl)
    mg = ag.marginals(post(ag.RowVecs(Xs), 0.05))
    assert np.allclose(mg.mu, mu) and np.allclose(mg.sigma, np.sqrt(v))
    mc, Cc = ag.mean_and_cov(post, ag.RowVecs(Xs))
    mc_r, Cc_r = ref.post_mean_and_cov(pr, Xs)
    assert np.allclose(mc, mc_r, **tol) and np.allclose(Cc, Cc_r, **tol)
    assert np.allclose(Cc, Cc.T, atol=tol["atol"])
    # cov(f_post, x, z) == cov(f_post, z, x)'  (src/util/TestUtils.jl:157)
    Zs = Xs[: m // 2] + 0.01
    assert np.allclose(ag.cov(post, ag.RowVecs(Xs), ag.RowVecs(Zs)), ag.cov(post, ag.RowVecs(Zs), ag.RowVecs(Xs)).T, **tol)


def test_posterior_collapses_on_data(ag):
    # test/exact_gpr_posterior.jl:21-22 with sigma^2 = 1e-15
    rng = np.random.default_rng(4)
    x = np.sort(rng.random(40)) * 10
    y = np.sin(x)
    f = ag.GP(ag.Matern52Kernel())
    post = ag.posterior(f(x, 1e-15), y)
    m, v = ag.mean_and_var(post, x)
    assert np.allclose(m, y, atol=1e-8) and np.allclose(v, 0, atol=1e-8)


def test_operator_api_on_device_factor(ag):
    # test/util/common_covmat_ops.jl:51-97 on the device factor (boundary #2)
    n, d = 260, 3
    ks, X, y = problem(n, d, ref.SE, np.float64, seed=21)
    f = ag.GP(mk_kernel(ag, ks))
    post = ag.posterior(f(ag.RowVecs(X), 0.1), y)
    A = ref.kernelmatrix(ks, X) + 0.1 * np.eye(n)
    rng = np.random.default_rng(6)
    B, B2 = rng.standard_normal((n, 5)), rng.standard_normal((n, 3))
    C = post.data.C
    assert np.allclose(ag.Xt_invA_X(C, B), B.T @ np.linalg.solve(A, B), rtol=1e-8, atol=1e-8)
    assert np.allclose(ag.Xt_invA_Y(B, C, B2), B.T @ np.linalg.solve(A, B2), rtol=1e-8, atol=1e-8)
    assert np.allclose(ag.diag_Xt_invA_X(C, B), np.diag(B.T @ np.linalg.solve(A, B)), rtol=1e-8)
    assert np.isclose(ag.tr_Xt_invA_X(C, B), np.trace(B.T @ np.linalg.solve(A, B)), rtol=1e-8)
    assert np.isclose(ag.Xt_invA_X(C, B[:, 0]), B[:, 0] @ np.linalg.solve(A, B[:, 0]), rtol=1e-8)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rand(ag, dtype):
    n, d = 301, 2
    ks, X, _ = problem(n, d, ref.MATERN52, dtype, seed=13)
    f = ag.GP(0.5, mk_kernel(ag, ks))
    fx = f(ag.RowVecs(X), 0.2)
    Z = np.random.default_rng(17).standard_normal((n, 6)).astype(dtype)
    S = ag.rand_from_normals(fx, Z)
    S_ref = ref.rand_from_Z(ks, ref.MeanSpec(1, 0.5), ref.NoiseSpec(0, 0.2), X, Z)
    assert S.dtype == dtype and S.shape == (n, 6)
    assert np.allclose(S, S_ref, rtol=1e-8 if dtype == np.float64 else 1e-3, atol=1e-9 if dtype == np.float64 else 1e-3)
    s1 = ag.rand(np.random.default_rng(1), fx)
    assert s1.shape == (n,)
    assert ag.rand(np.random.default_rng(1), fx, 3).shape == (n, 3)


def test_not_posdef_maps_to_exception(ag):
    x = np.zeros(200)
    f = ag.GP(ag.SqExponentialKernel())
    with pytest.raises(ag.PosDefException) as ei:
        ag.logpdf(f(x, -0.5), np.zeros(200))
    assert ei.value.info >= 1
    # the engine stays usable afterwards
    assert np.isfinite(ag.logpdf(f(np.linspace(0, 1, 20), 0.1), np.zeros(20)))


def test_dimension_mismatch(ag):
    f = ag.GP(ag.SqExponentialKernel())
    with pytest.raises(ag.DimensionMismatch):
        ag.logpdf(f(np.linspace(0, 1, 20), 0.1), np.zeros(19))
    post = ag.posterior(f(ag.RowVecs(np.zeros((5, 2)) + np.arange(5)[:, None]), 0.1), np.zeros(5))
    with pytest.raises(ag.DimensionMismatch):
        ag.mean_and_var(post, ag.RowVecs(np.zeros((3, 4))))


@pytest.mark.parametrize("name", ["c1.npz", "c2_n600.npz", "c3_n500_f32.npz"])
def test_golden_fixtures(ag, name):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    X, y, Xs = g["X"], g["y"], g["Xs"]
    dtype = X.dtype.type
    if name.startswith("c1"):
        f, s2 = ag.GP(ag.Matern32Kernel()), 0.001
    elif name.startswith("c2"):
        f, s2 = ag.GP(ag.with_lengthscale(ag.SqExponentialKernel(), np.sqrt(8) * 0.5)), 0.1
    else:
        f, s2 = ag.GP(ag.Matern32Kernel().compose(ag.ARDTransform(g["ard"]))), 0.05
    lp, post = ag.fit(f(ag.RowVecs(X), s2), y)
    rt = TOL[dtype]["rtol"]
    assert abs(lp - g["logpdf"]) <= rt * abs(g["logpdf"]) + (0 if dtype == np.float64 else 1e-3)
    sc = np.abs(g["alpha"]).max()
    assert np.allclose(post.data.alpha, g["alpha"], rtol=1e-6 if dtype == np.float64 else 2e-2,
                       atol=(1e-7 if dtype == np.float64 else 1e-2) * sc)
    m, v = ag.mean_and_var(post, ag.RowVecs(Xs))
    tol = dict(rtol=1e-6, atol=1e-8) if dtype == np.float64 else dict(rtol=5e-3, atol=5e-3)
    assert np.allclose(m, g["mean_s"], **tol) and np.allclose(v, g["var_s"], **tol)


def test_config_c2_full_size(ag):
    """BASELINE config C2: N=4096, D=8, SE, fp64 -- logpdf rtol 1e-8 vs the oracle."""
    cfg = ref.make_config("C2")
    f = ag.GP(ag.SqExponentialKernel().compose(ag.ScaleTransform(cfg["k"].scale)))
    lp, post = ag.fit(f(ag.RowVecs(cfg["X"]), 0.1), cfg["y"])
    lp_ref = ref.logpdf(cfg["k"], cfg["mean"], cfg["noise"], cfg["X"], cfg["y"])
    pr = ref.posterior(cfg["k"], cfg["mean"], cfg["noise"], cfg["X"], cfg["y"])
    assert abs(lp - lp_ref) <= 1e-8 * abs(lp_ref), (lp, lp_ref)
    assert np.allclose(post.data.alpha, pr["alpha"], rtol=1e-6, atol=1e-7 * np.abs(pr["alpha"]).max())


def test_config_c3_reduced(ag):
    """BASELINE config C3 shape at N=4096 (full N=16384 runs in bench.py): Matern32 o ARD, fp32,
    mean_and_var at 1000 test points, rtol 1e-4 on logpdf."""
    cfg = ref.make_config("C3", n=4096)
    f = ag.GP(ag.Matern32Kernel().compose(ag.ARDTransform(cfg["k"].ard)))
    lp, post = ag.fit(f(ag.RowVecs(cfg["X"]), 0.05), cfg["y"])
    X64 = cfg["X"].astype(np.float64)
    k64 = ref.KernelSpec(ref.MATERN32, 1.0, ref.T_ARD, ard=cfg["k"].ard.astype(np.float64))
    lp_ref = ref.logpdf(k64, cfg["mean"], cfg["noise"], X64, cfg["y"].astype(np.float64))
    assert abs(lp - lp_ref) <= 1e-4 * abs(lp_ref), (lp, lp_ref)
    Xs = cfg["Xs"][:1000]
    mu, v = ag.mean_and_var(post(ag.RowVecs(Xs), 0.05))
    pr = ref.posterior(k64, cfg["mean"], cfg["noise"], X64, cfg["y"].astype(np.float64))
    mu_r, v_r = ref.post_mean_and_var(pr, Xs.astype(np.float64), noise_s=cfg["noise"])
    assert np.allclose(mu, mu_r, rtol=2e-3, atol=2e-3) and np.allclose(v, v_r, rtol=2e-3, atol=2e-3)


# ---------------------------------------------------------------------------------------------
# sequential conditioning (update_chol) and VFE
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n1,n2", [(40, 30), (128, 128), (300, 77)])
def test_sequential_conditioning_equals_batch(ag, dtype, n1, n2):
    # test/exact_gpr_posterior.jl:29-43: C.U, alpha, x, delta of the sequential posterior == batch, atol 1e-5
    d = 3
    ks, X, y = problem(n1 + n2, d, ref.SE, dtype, seed=31)
    f = ag.GP(0.1, mk_kernel(ag, ks))
    nv2 = (0.05 + 0.1 * np.random.default_rng(3).random(n2)).astype(dtype)
    p1 = ag.posterior(f(ag.RowVecs(X[:n1]), 0.1), y[:n1])
    p2 = ag.posterior(p1(ag.RowVecs(X[n1:]), nv2), y[n1:])
    noise_all = ref.NoiseSpec(1, v=np.concatenate([np.full(n1, 0.1), nv2]).astype(dtype))
    pb = ref.posterior(ks, ref.MeanSpec(1, 0.1), noise_all, X, y)
    atol = 1e-5 if dtype == np.float64 else 5e-3
    assert p2.data.alpha.shape == (n1 + n2,)
    assert np.allclose(p2.data.alpha, pb["alpha"], atol=atol * max(1.0, np.abs(pb["alpha"]).max()))
    assert np.allclose(p2.data.delta, pb["delta"], atol=1e-6)
    assert np.allclose(p2.data.C.U, pb["U"], atol=atol)
    Xs = np.random.default_rng(5).random((50, d)).astype(dtype)
    m, v = ag.mean_and_var(p2, ag.RowVecs(Xs))
    m_r, v_r = ref.post_mean_and_var(pb, Xs)
    tol = dict(rtol=1e-6, atol=1e-7) if dtype == np.float64 else dict(rtol=5e-3, atol=5e-3)
    assert np.allclose(m, m_r, **tol) and np.allclose(v, v_r, **tol)
    # a third batch on top of the second
    ks3, X3, y3 = problem(25, d, ref.SE, dtype, seed=77)
    p3 = ag.posterior(p2(ag.RowVecs(X3), 0.2), y3)
    Xa, ya = np.concatenate([X, X3]), np.concatenate([y, y3])
    na = ref.NoiseSpec(1, v=np.concatenate([noise_all.v, np.full(25, 0.2)]).astype(dtype))
    pa = ref.posterior(ks, ref.MeanSpec(1, 0.1), na, Xa, ya)
    assert np.allclose(p3.data.alpha, pa["alpha"], atol=atol * max(1.0, np.abs(pa["alpha"]).max()))
    assert np.isclose(p3.data.C.logdet(), ref.logdet_chol(pa["U"]), rtol=1e-8 if dtype == np.float64 else 1e-3)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,d,fam", [(200, 20, 2, ref.SE), (1000, 130, 5, ref.MATERN52), (700, 256, 3, ref.SE)])
def test_vfe_elbo_and_posterior(ag, dtype, n, m, d, fam):
    ks, X, y = problem(n, d, fam, dtype, seed=41)
    Z = X[np.random.default_rng(1).permutation(n)[:m]].copy()
    jit = 1e-6 if dtype == np.float64 else 1e-3
    noise, jn = ref.NoiseSpec(0, 0.1), ref.NoiseSpec(0, jit)
    f = ag.GP(0.2, mk_kernel(ag, ks))
    fx = f(ag.RowVecs(X), 0.1)
    vfe = ag.VFE(f(ag.RowVecs(Z), jit))
    el, dt = ag.approx_log_evidence(vfe, fx, y, return_dtc=True)
    X64, y64, Z64 = X.astype(np.float64), y.astype(np.float64), Z.astype(np.float64)
    ks64 = ref.KernelSpec(ks.family, ks.variance, ks.transform, ks.scale, None, ks.linear_c)
    el_r = ref.elbo(ks64, ref.MeanSpec(1, 0.2), noise, X64, y64, Z64, jn)
    dt_r = ref.dtc(ks64, ref.MeanSpec(1, 0.2), noise, X64, y64, Z64, jn)
    rt = 1e-8 if dtype == np.float64 else 2e-3
    assert el.dtype == dtype
    assert abs(el - el_r) <= rt * abs(el_r), (el, el_r)
    assert abs(dt - dt_r) <= rt * abs(dt_r), (dt, dt_r)
    assert ag.elbo(vfe, fx, y) <= ag.logpdf(fx, y) + 1e-6 * abs(el_r)  # elbo <= logpdf (TestUtils.jl:214)
    # approximate posterior predictions vs the oracle (src/sparse_approximations.jl:212-217)
    vp = ag.posterior(vfe, fx, y)
    Xs = np.random.default_rng(2).random((60, d)).astype(dtype)
    mu, v = ag.mean_and_var(vp, ag.RowVecs(Xs))
    vr = ref.vfe_posterior(ks64, ref.MeanSpec(1, 0.2), noise, X64, y64, Z64, jn)
    mu_r, v_r = ref.vfe_mean_and_var(vr, Xs.astype(np.float64))
    tol = dict(rtol=1e-6, atol=1e-7) if dtype == np.float64 else dict(rtol=2e-2, atol=2e-2)
    assert np.allclose(mu, mu_r, **tol) and np.allclose(v, v_r, **tol)
    mu2, v2 = ag.mean_and_var(vp(ag.RowVecs(Xs), 0.1))
    assert np.allclose(v2, v + 0.1, atol=1e-6)


def test_vfe_with_z_equal_x_reproduces_exact(ag):
    # test/sparse_approximations.jl:24-25,94 ; src/util/TestUtils.jl:213-217 (rtol = atol = 1e-5)
    n, d = 150, 2
    ks, X, y = problem(n, d, ref.SE, np.float64, seed=43)
    f = ag.GP(mk_kernel(ag, ks))
    fx = f(ag.RowVecs(X), 0.1)
    vfe = ag.VFE(f(ag.RowVecs(X), 1e-10))
    assert np.isclose(ag.elbo(vfe, fx, y), ag.logpdf(fx, y), rtol=1e-5, atol=1e-5)
    vp, ep = ag.posterior(vfe, fx, y), ag.posterior(fx, y)
    Xs = np.random.default_rng(4).random((20, d))
    m1, v1 = ag.mean_and_var(vp, ag.RowVecs(Xs))
    m2, v2 = ag.mean_and_var(ep, ag.RowVecs(Xs))
    assert np.allclose(m1, m2, atol=1e-5) and np.allclose(v1, v2, atol=1e-5)
    with pytest.raises(ag.DimensionMismatch):
        ag.elbo(vfe, fx, y[:-1])


def test_vfe_golden_c5(ag):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c5_n3000_f64.npz"))
    X, y, Z = g["X"], g["y"], g["Z"]
    f = ag.GP(ag.with_lengthscale(ag.SqExponentialKernel(), np.sqrt(16) * 0.5))
    el, dt = ag.approx_log_evidence(ag.VFE(f(ag.RowVecs(Z), 1e-6)), f(ag.RowVecs(X), 0.1), y, return_dtc=True)
    assert abs(el - g["elbo"]) <= 1e-8 * abs(g["elbo"]) and abs(dt - g["dtc"]) <= 1e-8 * abs(g["dtc"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_logpdf_many_columns(ag, dtype):
    """logpdf(fx, Y::Matrix) with more columns than the border tile carries (src/finite_gp_projection.jl:306-311 has no
    limit): 128 ride through the factorisation, the rest are solved against the same factor in the same call"""
    n, d, S = 700, 3, 333
    ks, X, y = problem(n, d, ref.MATERN52, dtype, seed=21)
    rng = np.random.default_rng(6)
    Y = (y[:, None] * rng.random(S)[None, :] + 0.1 * rng.standard_normal((n, S))).astype(dtype)
    f = ag.GP(0.3, mk_kernel(ag, ks))
    lp = ag.logpdf(f(ag.RowVecs(X), 0.2), Y)
    lp_ref = ref.logpdf(ks, ref.MeanSpec(1, 0.3), ref.NoiseSpec(0, 0.2), X, Y)
    assert lp.shape == (S,)
    assert np.allclose(lp, lp_ref, rtol=TOL[dtype]["rtol"], atol=0 if dtype == np.float64 else 2e-2)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_chunked_prediction_rowvecs_and_colvecs(ag, dtype, monkeypatch):
    """test sets larger than one prediction chunk in BOTH storage orders (RowVecs = feature-major needs a strided gather per
    chunk): forced 256-column chunks must reproduce the single-chunk result exactly, for the exact and the VFE posterior"""
    n, d, M = 400, 3, 700
    ks, X, y = problem(n, d, ref.MATERN32, dtype, seed=9)
    Xs = np.random.default_rng(2).random((M, d)).astype(dtype)
    f = ag.GP(0.1, mk_kernel(ag, ks))
    post = ag.posterior(f(ag.RowVecs(X), 0.1), y)
    vp = ag.posterior(ag.VFE(f(ag.RowVecs(X[:60].copy()), 1e-4)), f(ag.RowVecs(X), 0.1), y)
    eng = ag.engine()
    import ctypes as C
    from agp_b200 import _cabi as cabi

    def raw_mean_var(handle_call, layout, arr):
        mu, var = np.zeros(M, dtype=dtype), np.zeros(M, dtype=dtype)
        eng.check(handle_call(layout, cabi.ptr(arr), mu, var))
        return mu, var

    def exact(layout, arr, mu, var):
        return eng.L.agp_post_mean_var(post.data.C.h, layout, arr, M, None, None, cabi.ptr(mu), cabi.ptr(var))

    def vfe(layout, arr, mu, var):
        return eng.L.agp_vfe_mean_var(vp.h, layout, arr, M, cabi.ptr(mu), cabi.ptr(var))
    pm = np.ascontiguousarray(Xs)            # point-major: [M, d] C-order
    fm = np.asfortranarray(Xs)               # feature-major: M x d column-major
    for call in (exact, vfe):
        base = raw_mean_var(call, cabi.AGP_POINT_MAJOR, pm)
        monkeypatch.setenv("AGP_PREDICT_CHUNK", "256")
        got_pm = raw_mean_var(call, cabi.AGP_POINT_MAJOR, pm)
        got_fm = raw_mean_var(call, cabi.AGP_FEATURE_MAJOR, fm)
        monkeypatch.delenv("AGP_PREDICT_CHUNK")
        for got in (got_pm, got_fm):
            tol = dict(rtol=1e-12, atol=1e-13) if dtype == np.float64 else dict(rtol=1e-5, atol=1e-6)
            assert np.allclose(got[0], base[0], **tol) and np.allclose(got[1], base[1], **tol)
