"""Host-side logic of the reference-facing mirror (no GPU): kernel algebra, input wrappers, means,
argument validation."""
import numpy as np
import pytest


def test_kernel_algebra(ag):
    k = 2.0 * ag.with_lengthscale(ag.Matern32Kernel(), 0.5)
    assert k.family == ag.api.MATERN32 and k.variance == 2.0 and np.isclose(k.transform.s, 2.0)
    k2 = ag.SqExponentialKernel().compose(ag.ARDTransform([1.0, 2.0])).compose(ag.ScaleTransform(3.0))
    assert np.allclose(k2.transform.v, [3.0, 6.0])
    k3 = ag.with_lengthscale(ag.SqExponentialKernel(), [0.5, 0.25])
    assert np.allclose(k3.transform.v, [2.0, 4.0])
    assert (ag.SqExponentialKernel() @ ag.ScaleTransform(2.0)) == ag.TransformedKernel(ag.SEKernel(), ag.ScaleTransform(2.0))
    assert ag.ScaledKernel(ag.LinearKernel(c=1.0), 3.0).variance == 3.0


def test_points_layouts(ag):
    X = np.arange(12.0).reshape(3, 4)  # 3 features x 4 points as ColVecs
    pc = ag.api._Points(ag.ColVecs(X))
    pr = ag.api._Points(ag.RowVecs(X.T))
    assert pc.n == 4 and pc.D == 3 and np.array_equal(pc.a, pr.a)
    pv = ag.api._Points(np.array([1.0, 2.0, 3.0]))
    assert pv.n == 3 and pv.D == 1
    assert len(ag.ColVecs(X)) == 4 and len(ag.RowVecs(X)) == 3
    with pytest.raises(TypeError):
        ag.api._Points(X)  # a bare matrix is ambiguous, as in the reference (needs obsdim)


def test_means(ag):
    x = ag.api._Points(np.array([0.0, 1.0, 2.0]))
    assert np.array_equal(ag.ZeroMean().vector(x, np.float64), np.zeros(3))
    assert np.array_equal(ag.ConstMean(2.5).vector(x, np.float64), np.full(3, 2.5))
    assert np.allclose(ag.CustomMean(np.sin).vector(x, np.float64), np.sin([0.0, 1.0, 2.0]))
    X = ag.api._Points(ag.RowVecs(np.ones((4, 2))))
    assert np.allclose(ag.CustomMean(lambda r: r.sum()).vector(X, np.float64), 2.0)
    f = ag.GP(3.0, ag.SqExponentialKernel())
    assert isinstance(f.mean, ag.ConstMean)
    assert isinstance(ag.GP(np.cos, ag.SqExponentialKernel()).mean, ag.CustomMean)
    assert isinstance(ag.GP(ag.SqExponentialKernel()).mean, ag.ZeroMean)


def test_finite_gp_construction(ag):
    f = ag.GP(ag.Matern52Kernel())
    fx = f(np.linspace(0, 1, 7), 0.1)
    assert len(fx) == 7 and np.allclose(fx.Sigma_y_diag, 0.1)
    fx2 = f(np.linspace(0, 1, 7))
    assert fx2.s2 == 1e-18  # default_sigma^2 (src/finite_gp_projection.jl:17)
    fx3 = f(np.linspace(0, 1, 7), np.full(7, 0.2))
    assert np.allclose(fx3.Sigma_y_diag, 0.2)
    with pytest.raises(ag.AGPError):
        f(np.linspace(0, 1, 3), np.eye(3))  # dense Sigma_y is outside the device path
    assert np.array_equal(ag.mean(fx), np.zeros(7))
    assert np.allclose(ag.var(f, np.zeros(3)), 1.0)


def test_reference_mean_function_testset(ag):
    """/root/reference/test/mean_function.jl:1-34: mean_vector for Zero / Const / Custom means on vectors, ColVecs, RowVecs."""
    rng = np.random.default_rng(123456)
    N, D = 5, 3
    x1 = rng.standard_normal(N)
    xc, xr = ag.ColVecs(rng.standard_normal((D, N))), ag.RowVecs(rng.standard_normal((N, D)))
    c = rng.standard_normal()
    foo = lambda x: np.sum(np.square(x))
    for x in (x1, xc, xr):
        assert np.array_equal(ag.mean_vector(ag.ZeroMean(), x), np.zeros(N))
        assert np.array_equal(ag.mean_vector(ag.ConstMean(c), x), np.full(N, c))
    assert np.allclose(ag.mean_vector(ag.CustomMean(foo), x1), [foo(v) for v in x1])
    assert np.allclose(ag.mean_vector(ag.CustomMean(foo), xc), [foo(col) for col in xc.X.T])   # eachcol
    assert np.allclose(ag.mean_vector(ag.CustomMean(foo), xr), [foo(row) for row in xr.X])     # eachrow


def test_reference_abstract_gp_testset_and_obsdim(ag):
    """/root/reference/test/abstract_gp.jl:1-9: mean(f) & co. without locations raise (ErrorException);
    /root/reference/test/finite_gp_projection.jl:38-43: f(Xmat; obsdim) == f(RowVecs / ColVecs)."""
    f = ag.GP(ag.SqExponentialKernel())
    for fn in (ag.mean, ag.var, ag.cov, ag.mean_and_var, ag.mean_and_cov):
        with pytest.raises(RuntimeError, match="not defined"):
            fn(f)
    Xmat = np.random.default_rng(0).standard_normal((1, 9))
    for x, obsdim in ((ag.RowVecs(Xmat), 1), (ag.ColVecs(Xmat), 2)):
        a, b = f(Xmat, obsdim=obsdim), f(x)
        assert np.array_equal(a.x.a, b.x.a) and a.s2 == b.s2
        a, b = f(Xmat, 1e-3, obsdim=obsdim), f(x, 1e-3)
        assert np.array_equal(a.x.a, b.x.a) and a.s2 == b.s2 == 1e-3
    with pytest.raises(TypeError):
        f(Xmat)  # ambiguous without obsdim
