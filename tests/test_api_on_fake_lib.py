"""The host mirror (abstractgps.jl_b200/api.py) driven end to end on the CPU against tests/fake_libagp.py, a stand-in for
libagp.so that answers the C ABI with the oracle.  What this pins: argument marshalling and layouts (point-major vs
RowVecs/ColVecs, column-major outputs), struct filling, dtype propagation, handle ownership (sequential conditioning
returns a new handle), chunking of > 128 right-hand sides, the error mapping (PosDefException / DimensionMismatch) -- and
the LOGIC of the GPU test files themselves: the replay of the reference's jldoctests and of its TestUtils suites, and
the posterior-FiniteGP cases, run here unchanged with the fake engine installed.  Nothing here exercises CUDA."""
import inspect

import numpy as np
import pytest

from oracle import agp_ref as ref
import fake_libagp
import test_gpu_posterior_finitegp as pf
import test_gpu_reference_doctests as doc


@pytest.fixture()
def fag(ag, monkeypatch):
    import ctypes as C
    eng = ag.api.Engine.__new__(ag.api.Engine)
    eng.L, eng.h, eng.device = fake_libagp.FakeLib(), C.c_void_p(1), 0
    monkeypatch.setattr(ag.api, "_engine", eng)
    yield ag


def _call(fn, fag, **extra):
    kw = {}
    for name in inspect.signature(fn).parameters:
        if name == "ag":
            kw[name] = fag
        elif name == "rng":
            kw[name] = np.random.default_rng(20240924)
        else:
            kw[name] = extra[name]
    return fn(**kw)


DOC_TESTS = [n for n in dir(doc) if n.startswith("test_")]


@pytest.mark.parametrize("name", DOC_TESTS)
def test_reference_doctest_replays(fag, name):
    _call(getattr(doc, name), fag)


def test_reference_testutils_suites(fag):
    pf.test_reference_base_gp_suite(fag)
    pf.test_reference_exact_gpr_posterior_suite(fag)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_posterior_finitegp_cases(fag, dtype):
    pf.test_posterior_finitegp_logpdf_and_rand(fag, dtype, 129, 40, 3, ref.MATERN52)
    if dtype == np.float64:
        pf.test_posterior_logpdf_is_the_chain_rule_increment(fag)
        pf.test_sequential_conditioning_has_value_semantics(fag)
        pf.test_posterior_finitegp_not_posdef(fag)


def test_fit_matches_oracle_through_every_input_wrapper(fag):
    ag = fag
    rng = np.random.default_rng(0)
    n, d = 60, 3
    X = rng.random((n, d))
    y = np.sin(X.sum(1))
    ks = ref.KernelSpec(ref.MATERN32, 1.7, ref.T_ARD, ard=np.array([1.0, 2.0, 0.5]))
    k = 1.7 * ag.Matern32Kernel().compose(ag.ARDTransform([1.0, 2.0, 0.5]))
    want = ref.logpdf(ks, ref.MeanSpec(1, 0.3), ref.NoiseSpec(0, 0.1), X, y)
    f = ag.GP(0.3, k)
    for x in (ag.RowVecs(X), ag.ColVecs(np.ascontiguousarray(X.T)), ag.ColVecs(np.asfortranarray(X.T))):
        assert np.isclose(ag.logpdf(f(x, 0.1), y), want, rtol=1e-12)
    lp, p = ag.fit(f(ag.RowVecs(X), 0.1), y)
    pr = ref.posterior(ks, ref.MeanSpec(1, 0.3), ref.NoiseSpec(0, 0.1), X, y)
    assert np.allclose(p.data.alpha, pr["alpha"]) and np.allclose(p.data.C.U, pr["U"]) and np.allclose(p.data.delta, pr["delta"])
    assert np.isclose(p.data.C.logdet(), ref.logdet_chol(pr["U"]))
    # the operator API on the factor handle
    B = rng.standard_normal((n, 4))
    assert np.allclose(ag.Xt_invA_X(p.data.C, B), ref.Xt_invA_X(pr["U"], B))
    assert np.allclose(ag.diag_Xt_invA_X(p.data.C, B), ref.diag_Xt_invA_X(pr["U"], B))
    assert np.isclose(ag.tr_Xt_invA_X(p.data.C, B[:, 0]), ref.tr_Xt_invA_X(pr["U"], B[:, 0]))
    # 1-D inputs are D = 1 points; float32 inputs stay float32 end to end
    x1 = rng.random(20).astype(np.float32)
    y1 = np.cos(x1).astype(np.float32)
    lp32 = ag.logpdf(ag.GP(ag.SqExponentialKernel())(x1, 0.1), y1)
    assert lp32.dtype == np.float32


def test_more_than_128_right_hand_sides_are_one_call(fag):
    ag = fag
    rng = np.random.default_rng(1)
    X = rng.random((30, 2))
    Y = rng.standard_normal((30, 300))
    fx = ag.GP(ag.SqExponentialKernel())(ag.RowVecs(X), 0.2)
    lps = ag.logpdf(fx, Y)
    want = ref.logpdf(ref.KernelSpec(ref.SE, 1.0), ref.MeanSpec(), ref.NoiseSpec(0, 0.2), X, Y)
    assert lps.shape == (300,) and np.allclose(lps, want)
    assert fag.api.engine().L.calls.count("agp_fit") == 1  # the library handles any number of columns with ONE factorisation
    assert np.isclose(ag.loglikelihood(fx, Y), want.sum())


def test_error_mapping(fag):
    ag = fag
    f = ag.GP(ag.SqExponentialKernel())
    x = np.linspace(0, 1, 5)
    with pytest.raises(ag.PosDefException) as e:
        ag.logpdf(f(np.array([0.0, 0.0, 1.0]), -1.0), np.zeros(3))  # K - I is indefinite
    assert e.value.info >= 1
    with pytest.raises(ag.DimensionMismatch):
        ag.logpdf(f(x, 0.1), np.zeros(4))
    p = ag.posterior(f(x, 0.1), np.zeros(5))
    with pytest.raises(ag.DimensionMismatch):
        ag.mean_and_var(p(ag.RowVecs(np.zeros((3, 2)))))
    with pytest.raises(ag.DimensionMismatch):
        ag.elbo(ag.VFE(f(x[:2])), f(x, 0.1), np.zeros(4))


def test_vfe_wrappers(fag):
    ag = fag
    rng = np.random.default_rng(2)
    X, Z = rng.random((80, 2)), rng.random((9, 2))
    y = np.sin(3 * X[:, 0])
    ks = ref.KernelSpec(ref.MATERN52, 1.0, ref.T_SCALE, scale=2.0)
    f = ag.GP(ag.Matern52Kernel().compose(ag.ScaleTransform(2.0)))
    fx = f(ag.RowVecs(X), 0.1)
    vfe = ag.VFE(f(ag.RowVecs(Z), 1e-6))
    jit = ref.NoiseSpec(0, 1e-6)
    el, dtc = ag.approx_log_evidence(vfe, fx, y, return_dtc=True)
    assert np.isclose(el, ref.elbo(ks, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y, Z, jit))
    assert np.isclose(dtc, ref.dtc(ks, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y, Z, jit))
    post = ag.posterior(vfe, fx, y)
    Xs = rng.random((7, 2))
    m, v = ag.mean_and_var(post(ag.RowVecs(Xs), 0.05))
    mr, vr = ref.vfe_mean_and_var(ref.vfe_posterior(ks, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y, Z, jit), Xs)
    assert np.allclose(m, mr) and np.allclose(v, vr + 0.05)
    vp = ref.vfe_posterior(ks, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y, Z, jit)
    Zs = rng.random((5, 2))
    assert np.allclose(ag.cov(post, ag.RowVecs(Xs)), ref.vfe_mean_and_cov(vp, Xs)[1])
    assert np.allclose(ag.cov(post, ag.RowVecs(Xs), ag.RowVecs(Zs)), ref.vfe_cov_cross(vp, Xs, Zs))
    mc, Cc = ag.mean_and_cov(post(ag.RowVecs(Xs), 0.05))
    assert np.allclose(mc, mr) and np.allclose(Cc, ref.vfe_mean_and_cov(vp, Xs)[1] + 0.05 * np.eye(7))


@pytest.mark.parametrize("transform", ["scale", "ard"])
@pytest.mark.parametrize("fam", [ref.SE, ref.LINEAR])
def test_logpdf_grad_mapping(fag, fam, transform):
    """the dict ag.logpdf_grad builds from the ABI's 5 + D vector (variance / scale|ard / linear_c / noise / mean)"""
    ag = fag
    rng = np.random.default_rng(3)
    n, D = 40, 3
    X = rng.random((n, D))
    y = np.sin(X[:, 0])
    ard = np.array([1.2, 0.7, 2.1])
    ks = ref.KernelSpec(fam, 1.3, ref.T_SCALE if transform == "scale" else ref.T_ARD, scale=1.7, ard=ard, linear_c=0.4)
    k = ag.SqExponentialKernel() if fam == ref.SE else ag.LinearKernel(c=0.4)
    k = 1.3 * k.compose(ag.ScaleTransform(1.7) if transform == "scale" else ag.ARDTransform(ard))
    nv = 0.05 + 0.1 * rng.random(n)
    lp, g = ag.logpdf_grad(ag.GP(0.25, k)(ag.RowVecs(X), nv), y)
    want = ref.logpdf_grad(ks, ref.MeanSpec(1, 0.25), ref.NoiseSpec(1, v=nv), X, y)
    assert set(g) == set(want)
    for key in want:
        assert np.allclose(g[key], want[key]), key
    lp2, g2 = ag.logpdf_grad(ag.GP(lambda r: np.sin(r[0]), k)(ag.RowVecs(X), 0.1), y)  # scalar noise, vector mean
    assert np.ndim(g2["noise"]) == 0 and g2["mean_v"].shape == (n,) and "mean_c" not in g2


# ---- the reference's own test sets for the hot path (tests/ref_suite_replays.py), on the fake library
import ref_suite_replays as rs  # noqa: E402


def test_reference_finite_gp_testsets(fag):
    rs.finite_gp_statistics(fag)
    rs.finite_gp_rand_statistical(fag, S=100_000)
    rs.finite_gp_logpdf(fag)
    for T in (np.float64, np.float32):
        rs.finite_gp_type_stability(fag, T)


@pytest.mark.parametrize("approx_name", ["VFE", "DTC"])
def test_reference_sparse_testsets(fag, approx_name):
    A = getattr(fag, approx_name)
    rs.sparse_approx_log_evidence(fag, A)
    rs.sparse_posterior_matches_exact(fag, A)
    rs.sparse_update_posterior(fag, A)
    rs.sparse_internal_interface(fag, A, pf)
    for T in (np.float64, np.float32):
        rs.sparse_type_stability(fag, A, T)


# ---- the host-side usage of the GPU parity suite (tests/test_gpu_parity.py), re-run against the fake library: a
# regression net for api.py that needs no device (numerics are the oracle's on both sides, so only the plumbing is tested)
import test_gpu_parity as gp  # noqa: E402

PARITY_CASES = [
    (gp.test_gram, dict(fam=ref.MATERN52, dtype=np.float64, n=63, m=70, d=3)),
    (gp.test_gram, dict(fam=ref.LINEAR, dtype=np.float32, n=129, m=5, d=8)),
    (gp.test_logpdf_posterior, dict(dtype=np.float64, n=129, d=3, fam=ref.SE)),
    (gp.test_logpdf_posterior, dict(dtype=np.float32, n=10, d=1, fam=ref.MATERN32)),
    (gp.test_multicolumn_and_noise_and_mean_variants, dict(dtype=np.float64)),
    (gp.test_multicolumn_and_noise_and_mean_variants, dict(dtype=np.float32)),
    (gp.test_mean_and_var_and_cov, dict(dtype=np.float64, n=200, m=77, d=3, fam=ref.SE)),
    (gp.test_posterior_collapses_on_data, {}),
    (gp.test_operator_api_on_device_factor, {}),
    (gp.test_rand, dict(dtype=np.float64)),
    (gp.test_not_posdef_maps_to_exception, {}),
    (gp.test_dimension_mismatch, {}),
    (gp.test_golden_fixtures, dict(name="c1.npz")),
    (gp.test_golden_fixtures, dict(name="c3_n500_f32.npz")),
    (gp.test_sequential_conditioning_equals_batch, dict(dtype=np.float64, n1=40, n2=30)),
    (gp.test_vfe_elbo_and_posterior, dict(dtype=np.float64, n=200, m=20, d=2, fam=ref.SE)),
    (gp.test_vfe_with_z_equal_x_reproduces_exact, {}),
]


@pytest.mark.parametrize("case", range(len(PARITY_CASES)))
def test_gpu_parity_suite_plumbing(fag, case):
    fn, kw = PARITY_CASES[case]
    _call(fn, fag, **kw)


def test_smoke_entry_plumbing(fag, capsys):
    """__graft_entry__.smoke() end to end with the fake engine installed: its host calls stay valid as api.py evolves
    (the real smoke runs the same body on cuda:0)."""
    import __graft_entry__ as ge
    ge.smoke()
    assert "smoke ok" in capsys.readouterr().out


# ---- randomized differential test of the marshalling: kernel algebra -> agp_kernel, means, noise, wrappers, dtypes
def _random_case(rng):
    fam = int(rng.integers(0, 5))
    D = int(rng.integers(1, 5))
    n = int(rng.integers(5, 40))
    dtype = [np.float64, np.float32][int(rng.integers(0, 2))]
    X = rng.random((n, D))
    var = float(0.5 + rng.random())
    c = float(rng.random())
    tkind = int(rng.integers(0, 5))
    scale, ard = 1.0, None
    return dict(fam=fam, D=D, n=n, dtype=dtype, X=X, var=var, c=c, tkind=tkind, scale=scale, ard=ard)


def _build(ag, rng, cs):
    """build the kernel through a random but equivalent sequence of the reference's constructors; return (ag kernel, KernelSpec)"""
    fam, D = cs["fam"], cs["D"]
    base = {ref.SE: ag.SqExponentialKernel, ref.MATERN12: ag.Matern12Kernel, ref.MATERN32: ag.Matern32Kernel,
            ref.MATERN52: ag.Matern52Kernel}.get(fam)
    k = base() if base else ag.LinearKernel(c=cs["c"])
    tk = cs["tkind"]
    spec = dict(transform=ref.T_NONE, scale=1.0, ard=None)
    if tk == 1:      # ScaleTransform
        s = float(0.5 + rng.random())
        k = k.compose(ag.ScaleTransform(s))
        spec = dict(transform=ref.T_SCALE, scale=s, ard=None)
    elif tk == 2:    # with_lengthscale scalar, then another ScaleTransform composed on top (applied first)
        ell, s2 = float(0.5 + rng.random()), float(0.5 + rng.random())
        k = ag.TransformedKernel(ag.with_lengthscale(k, ell), ag.ScaleTransform(s2))
        spec = dict(transform=ref.T_SCALE, scale=s2 / ell, ard=None)
    elif tk == 3:    # ARD
        v = 0.5 + rng.random(D)
        k = k @ ag.ARDTransform(v)
        spec = dict(transform=ref.T_ARD, scale=1.0, ard=v)
    elif tk == 4:    # vector lengthscale then a scalar scale
        ell, s2 = 0.5 + rng.random(D), float(0.5 + rng.random())
        k = ag.with_lengthscale(k, ell).compose(ag.ScaleTransform(s2))
        spec = dict(transform=ref.T_ARD, scale=1.0, ard=s2 / ell)
    # variance through a random split of ScaledKernel / scalar products
    a = float(0.5 + rng.random())
    k = ag.ScaledKernel(a * k, cs["var"] / a) if rng.random() < 0.5 else (cs["var"] / a) * (k * a)
    ks = ref.KernelSpec(fam, cs["var"], spec["transform"], scale=spec["scale"], ard=spec["ard"], linear_c=cs["c"])
    return k, ks


@pytest.mark.parametrize("seed", range(40))
def test_randomized_marshalling(fag, seed):
    ag = fag
    rng = np.random.default_rng(1000 + seed)
    cs = _random_case(rng)
    k, ks = _build(ag, rng, cs)
    n, D, dt = cs["n"], cs["D"], cs["dtype"]
    X = cs["X"].astype(dt)
    mk = int(rng.integers(0, 3))
    if mk == 0:
        f, mean = ag.GP(k), ref.MeanSpec()
    elif mk == 1:
        cm = dt(rng.standard_normal())
        f, mean = ag.GP(cm, k), ref.MeanSpec(1, float(cm))
    else:
        fn = (lambda r: float(np.sum(r)) ** 2) if D > 1 else (lambda r: float(r) ** 2)
        f = ag.GP(fn, k)
        mean = ref.MeanSpec(2, v=np.array([fn(r) for r in (X if D > 1 else X[:, 0])], dtype=dt))
    if rng.random() < 0.5:
        s2 = float(0.05 + rng.random())
        noise = ref.NoiseSpec(0, s2)
    else:
        s2 = (0.05 + rng.random(n)).astype(dt)
        noise = ref.NoiseSpec(1, v=s2)
    wrap = int(rng.integers(0, 3)) if D > 1 else int(rng.integers(0, 4))
    x = [ag.RowVecs(X), ag.ColVecs(np.ascontiguousarray(X.T)), ag.ColVecs(np.asfortranarray(X.T)), X[:, 0]][wrap]
    fx = f(x, s2)
    y = rng.standard_normal(n).astype(dt)
    tol = dict(rtol=1e-10, atol=1e-12) if dt == np.float64 else dict(rtol=1e-4, atol=1e-5)
    assert np.allclose(ag.logpdf(fx, y), ref.logpdf(ks, mean, noise, X, y), **tol)
    assert np.allclose(ag.cov(fx), ref.mean_and_cov_fx(ks, mean, noise, X)[1], **tol)
    assert ag.logpdf(fx, y).dtype == dt
    p = ag.posterior(fx, y)
    pr = ref.posterior(ks, mean, noise, X, y)
    m = int(rng.integers(1, 9))
    Xs = rng.random((m, D)).astype(dt)
    xs = ag.RowVecs(Xs) if D > 1 else Xs[:, 0]
    mean_s = None if mk != 2 else ref.MeanSpec(2, v=np.array([fn(r) for r in (Xs if D > 1 else Xs[:, 0])], dtype=dt))
    mm, vv = ag.mean_and_var(p, xs)
    mr, vr = ref.post_mean_and_var(pr, Xs, mean_s)
    assert np.allclose(mm, mr, **tol) and np.allclose(vv, vr, **tol) and mm.dtype == dt
    Z = rng.standard_normal((n, 2)).astype(dt)
    assert np.allclose(ag.rand_from_normals(fx, Z), ref.rand_from_Z(ks, mean, noise, X, Z), **tol)
