"""CPU model of the experimental gradient path (abstractgps.jl_b200/csrc/grad.cu + post_logpdf_grad_impl in engine.cu): the
same sums over the LOWER triangle with off-diagonal elements counted twice, computed from the TRANSFORMED points exactly
as the kernel does, and the same host-side finalisation constants -- against the finite-difference-pinned gradient oracle.
This pins the formulas; the device kernel itself is still to be validated on a GPU (tests/test_gpu_variants_grad_vfecov.py)."""
import numpy as np
import pytest

from oracle import agp_ref as ref


def kappa_pair(fam, d2):
    if fam == ref.SE:
        e = np.exp(-0.5 * d2)
        return e, -d2 * e
    if fam == ref.MATERN12:
        d = np.sqrt(d2)
        e = np.exp(-d)
        return e, -d * e
    if fam == ref.MATERN32:
        s = np.sqrt(3.0) * np.sqrt(d2)
        e = np.exp(-s)
        return (1 + s) * e, -3.0 * d2 * e
    s = np.sqrt(5.0) * np.sqrt(d2)
    e = np.exp(-s)
    return (1 + s + s * s / 3.0) * e, -(5.0 / 3.0) * d2 * (1 + s) * e


def device_model(k, mean, noise, X, y):
    n, D = X.shape
    m, C = ref.mean_and_cov_fx(k, mean, noise, X)
    U = ref.cholesky_upper(C)
    alpha = ref._U_solve(U, ref._Ut_solve(U, y - m))
    V = ref._Ut_solve(U, np.eye(n))      # L^-1
    Cinv = V.T @ V
    Xt = k.apply_transform(X)
    sums = np.zeros(5 + D)
    linear = k.family == ref.LINEAR
    gi, gj = np.tril_indices(n)
    mult = np.where(gi == gj, 1.0, 2.0)
    w = alpha[gi] * alpha[gj] - Cinv[gi, gj]
    if linear:
        acc = np.einsum("ij,ij->i", Xt[gi], Xt[gj])
        sums[0] = np.sum(mult * w * (acc + k.linear_c))
        sums[1] = np.sum(mult * w * acc)
        sums[2] = np.sum(mult * w)
        wq = mult * w
    else:
        df = Xt[gi] - Xt[gj]
        d2 = np.where(gi == gj, 0.0, np.einsum("ij,ij->i", df, df))
        kap, kr = kappa_pair(k.family, d2)
        sums[0] = np.sum(mult * w * kap)
        sums[1] = np.sum(mult * w * kr)
        with np.errstate(divide="ignore", invalid="ignore"):
            wq = np.where(d2 > 0, mult * w * kr / d2, 0.0)
    diag = gi == gj
    sums[3] = np.sum(w[diag])
    sums[4] = np.sum(alpha)
    for d in range(D):
        sums[5 + d] = np.sum(wq * Xt[gi, d] * Xt[gj, d]) if linear else np.sum(wq * (Xt[gi, d] - Xt[gj, d]) ** 2)
    # post_logpdf_grad_impl's finalisation
    var, sc = k.variance, k.scale
    g = np.zeros(5 + D)
    g[0] = 0.5 * sums[0]
    g[1] = (var * sums[1] / sc if linear else 0.5 * var * sums[1] / sc) if k.transform == ref.T_SCALE else 0.0
    g[2] = 0.5 * var * sums[2] if linear else 0.0
    g[3] = 0.5 * sums[3]
    g[4] = sums[4]
    if k.transform == ref.T_ARD:
        v = np.asarray(k.ard, dtype=np.float64)
        g[5:] = (var if linear else 0.5 * var) * sums[5:] / v
    return g, 0.5 * w[diag]


@pytest.mark.parametrize("fam", [ref.SE, ref.MATERN12, ref.MATERN32, ref.MATERN52, ref.LINEAR])
@pytest.mark.parametrize("transform", [ref.T_SCALE, ref.T_ARD])
def test_device_formulas_match_oracle(fam, transform):
    rng = np.random.default_rng(11)
    n, D = 70, 4
    X = rng.random((n, D))
    y = np.sin(3 * X[:, 0]) + 0.2 * rng.standard_normal(n)
    k = ref.KernelSpec(fam, 1.3, transform, scale=1.7, ard=np.array([1.2, 0.7, 2.1, 0.9]), linear_c=0.4)
    mean, noise = ref.MeanSpec(1, 0.25), ref.NoiseSpec(1, v=0.05 + 0.1 * rng.random(n))
    want = ref.logpdf_grad(k, mean, noise, X, y)
    g, nd = device_model(k, mean, noise, X, y)
    tol = dict(rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(g[0], want["variance"], **tol)
    np.testing.assert_allclose(g[4], want["mean_c"], **tol)
    np.testing.assert_allclose(nd, want["noise"], **tol)
    if transform == ref.T_SCALE:
        np.testing.assert_allclose(g[1], want["scale"], **tol)
    else:
        np.testing.assert_allclose(g[5:], want["ard"], **tol)
    if fam == ref.LINEAR:
        np.testing.assert_allclose(g[2], want["linear_c"], **tol)
    gs = ref.logpdf_grad(k, mean, ref.NoiseSpec(0, 0.1), X, y)
    g2, _ = device_model(k, mean, ref.NoiseSpec(0, 0.1), X, y)
    np.testing.assert_allclose(g2[3], gs["noise"], **tol)
