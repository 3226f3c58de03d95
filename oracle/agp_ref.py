"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (NumPy + SciPy/LAPACK) of the AbstractGPs.jl dense hot path:
Gram -> (+ noise) -> Cholesky -> triangular solves -> logpdf / posterior /
mean_and_var / rand, and the VFE ``elbo`` of sparse_approximations.jl.  Every
function cites the reference file:line (relative to /root/reference) whose
operation order it follows.

PARITY STATUS: **parity unpinned** against the Julia reference itself.
Julia is not installed in this image and the reference ships no golden
vectors / known-answer fixtures, so the reference cannot
be run here.  The oracle is instead pinned (tests/test_oracle.py) against
independent implementations: scipy.stats.multivariate_normal.logpdf (mirror
of test/finite_gp_projection.jl:143), naive inv/det formulas, and the
relations the reference's own tests assert (collapse-on-data, sequential ==
batch, VFE(z=x) == exact, elbo <= logpdf, operator identities); and
(tests/test_oracle_independent_pins.py) against scikit-learn's
GaussianProcessRegressor -- kernel matrices, log marginal likelihood, alpha,
predictive mean / std for RBF, Matern 1/2, 3/2, 5/2 and DotProduct -- and
against 60-digit mpmath arithmetic for logpdf, alpha, the DTC objective and the
Titsias bound.  Its two distance formulations ("direct", what the CUDA kernel
computes, and "gemm", what Distances.jl executes in the reference) are bounded
against each other in tests/test_oracle.py.

Kernel formulas are those of KernelFunctions.jl (un-vendored dependency,
compat "0.9, 0.10", Project.toml:26), restated from its documented closed
forms: they are the textbook Rasmussen & Williams forms.

Conventions: inputs are ``X[N, D]`` (row = one point).  The factor is the
UPPER ``U`` with ``Sigma = U'U`` exactly like ``cholesky(Symmetric(.))`` in
the reference.  ``dtype`` float64 or float32 -- all arithmetic is done in that
dtype (the reference is type-stable, test/finite_gp_projection.jl:180-191).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import scipy.linalg as sla

LOG2PI = math.log(2.0 * math.pi)

SE, MATERN12, MATERN32, MATERN52, LINEAR = 0, 1, 2, 3, 4
T_NONE, T_SCALE, T_ARD = 0, 1, 2


@dataclass
class KernelSpec:
    """sigma_f^2 * (kappa o transform).  Mirrors KernelFunctions' ScaledKernel /
    TransformedKernel{ScaleTransform|ARDTransform} composition used at
    src/base_gp.jl:70,72,74."""

    family: int = SE
    variance: float = 1.0
    transform: int = T_NONE
    scale: float = 1.0
    ard: Optional[np.ndarray] = None
    linear_c: float = 0.0

    def apply_transform(self, X: np.ndarray) -> np.ndarray:
        if self.transform == T_NONE:
            return X
        if self.transform == T_SCALE:
            return X * X.dtype.type(self.scale)
        return X * np.asarray(self.ard, dtype=X.dtype)[None, :]


@dataclass
class MeanSpec:
    """ZeroMean / ConstMean / CustomMean (evaluated to a vector host-side).
    src/mean_function.jl:27,40,52-55."""

    kind: int = 0  # 0 zero, 1 const, 2 vector
    c: float = 0.0
    v: Optional[np.ndarray] = None

    def vector(self, n: int, dtype) -> np.ndarray:
        if self.kind == 0:
            return np.zeros(n, dtype=dtype)
        if self.kind == 1:
            return np.full(n, self.c, dtype=dtype)
        v = np.asarray(self.v, dtype=dtype)
        assert v.shape == (n,)
        return v


@dataclass
class NoiseSpec:
    """Diagonal Sigma_y: scalar (Fill) or per-point vector.
    src/finite_gp_projection.jl:13-21."""

    kind: int = 0  # 0 scalar, 1 vector
    s: float = 1e-18  # default_sigma^2, src/finite_gp_projection.jl:17
    v: Optional[np.ndarray] = None

    def diag(self, n: int, dtype) -> np.ndarray:
        if self.kind == 0:
            return np.full(n, self.s, dtype=dtype)
        v = np.asarray(self.v, dtype=dtype)
        assert v.shape == (n,)
        return v


# ---------------------------------------------------------------------------
# Gram construction  (KernelFunctions.kernelmatrix; call sites src/base_gp.jl:70-74)
# ---------------------------------------------------------------------------

def _pairwise_sqdist(A: np.ndarray, B: np.ndarray, method: str) -> np.ndarray:
    if method == "gemm":  # Distances.jl form ||a||^2+||b||^2-2a.b clamped at 0
        d2 = (A * A).sum(1)[:, None] + (B * B).sum(1)[None, :] - 2.0 * (A @ B.T)
        return np.maximum(d2, 0).astype(A.dtype)
    # direct differences, chunked (what the CUDA kernel does)
    n, m = A.shape[0], B.shape[0]
    out = np.empty((n, m), dtype=A.dtype)
    step = max(1, int(4e6 // max(1, m * A.shape[1])))
    for i in range(0, n, step):
        diff = A[i:i + step, None, :] - B[None, :, :]
        out[i:i + step] = np.einsum("ijk,ijk->ij", diff, diff)
    return out


def _kappa(family: int, d2: np.ndarray) -> np.ndarray:
    T = d2.dtype.type
    if family == SE:
        return np.exp(-d2 / T(2))
    d = np.sqrt(d2)
    if family == MATERN12:
        return np.exp(-d)
    if family == MATERN32:
        s = T(math.sqrt(3.0)) * d
        return (T(1) + s) * np.exp(-s)
    if family == MATERN52:
        s = T(math.sqrt(5.0)) * d
        return (T(1) + s + s * s / T(3)) * np.exp(-s)
    raise ValueError(family)


# "direct" (difference form, what the CUDA kernel computes) or "gemm" (the Distances.jl formulation the
# reference's KernelFunctions actually executes; used for the timed CPU baseline in bench.py)
DEFAULT_METHOD = "direct"


def kernelmatrix(k: KernelSpec, X: np.ndarray, Z: Optional[np.ndarray] = None,
                 method: Optional[str] = None) -> np.ndarray:
    """kernelmatrix(k, x[, z]).  src/base_gp.jl:70 (one-arg), :74 (two-arg)."""
    method = method or DEFAULT_METHOD
    dtype = X.dtype
    Xt = k.apply_transform(X)
    Zt = Xt if Z is None else k.apply_transform(np.asarray(Z, dtype=dtype))
    if k.family == LINEAR:
        K = Xt @ Zt.T + dtype.type(k.linear_c)
    else:
        d2 = _pairwise_sqdist(Xt, Zt, method)
        if Z is None:
            np.fill_diagonal(d2, 0)  # exactly-zero self distance
        K = _kappa(k.family, d2)
    return (dtype.type(k.variance) * K).astype(dtype)


def kernelmatrix_diag(k: KernelSpec, X: np.ndarray) -> np.ndarray:
    """kernelmatrix_diag(k, x).  src/base_gp.jl:72."""
    dtype = X.dtype
    if k.family == LINEAR:
        Xt = k.apply_transform(X)
        return (dtype.type(k.variance) * ((Xt * Xt).sum(1) + dtype.type(k.linear_c))).astype(dtype)
    return np.full(X.shape[0], k.variance, dtype=dtype)


# ---------------------------------------------------------------------------
# FiniteGP  (src/finite_gp_projection.jl)
# ---------------------------------------------------------------------------

def mean_and_cov_fx(k, mean, noise, X):
    """mean_and_cov(::FiniteGP) src/finite_gp_projection.jl:133-136 : (m, K + Sigma_y)."""
    n = X.shape[0]
    m = mean.vector(n, X.dtype)
    C = kernelmatrix(k, X)
    C[np.diag_indices(n)] += noise.diag(n, X.dtype)
    return m, C


def cholesky_upper(C: np.ndarray) -> np.ndarray:
    """cholesky(_symmetric(C)).U  (src/util/common_covmat_ops.jl:5, LAPACK potrf('U')).
    Raises numpy.linalg.LinAlgError on a non-PD matrix (PosDefException analogue)."""
    return sla.cholesky(C, lower=False, check_finite=False)


def _Ut_solve(U, B):
    """U' \\ B  (trtrs).  src/util/common_covmat_ops.jl:54,90,101."""
    return sla.solve_triangular(U, B, trans="T", lower=False, check_finite=False)


def _U_solve(U, B):
    return sla.solve_triangular(U, B, trans="N", lower=False, check_finite=False)


def diag_At_A(A):
    """src/util/common_covmat_ops.jl:64-65."""
    if A.ndim == 1:
        return np.array([np.sum(A * A)])
    return np.sum(A * A, axis=0)


def tr_At_A(A):
    """src/util/common_covmat_ops.jl:67."""
    return np.sum(A * A)


def Xt_invA_X(U, X):
    """src/util/common_covmat_ops.jl:54-58."""
    V = _Ut_solve(U, X)
    if X.ndim == 1:
        return np.sum(V * V)
    return V.T @ V


def Xt_invA_Y(X, U, Y):
    """src/util/common_covmat_ops.jl:60."""
    return _Ut_solve(U, X).T @ _Ut_solve(U, Y)


def diag_Xt_invA_X(U, X):
    """src/util/common_covmat_ops.jl:90."""
    return diag_At_A(_Ut_solve(U, X))


def tr_Xt_invA_X(U, X):
    """src/util/common_covmat_ops.jl:101."""
    return tr_At_A(_Ut_solve(U, X))


def logdet_chol(U):
    return 2.0 * np.sum(np.log(np.diag(U).astype(np.float64)))


def logpdf(k, mean, noise, X, Y):
    """logpdf(::FiniteGP, Y) src/finite_gp_projection.jl:306-311, _sqmahal :325-326.
    Y vector -> scalar; Y matrix [N, S] -> vector of S."""
    dtype = X.dtype
    Y = np.asarray(Y, dtype=dtype)
    m, C = mean_and_cov_fx(k, mean, noise, X)
    U = cholesky_upper(C)
    n = X.shape[0]
    ld = dtype.type(logdet_chol(U))
    if Y.ndim == 1:
        sq = tr_Xt_invA_X(U, Y - m)
    else:
        sq = diag_Xt_invA_X(U, Y - m[:, None])
    return -((n * dtype.type(LOG2PI) + ld) + sq) / dtype.type(2)


def _dkappa_r(family: int, d2: np.ndarray) -> np.ndarray:
    """kappa'(r) * r for the stationary families (r = distance after the transform): the common factor of the
    derivatives of kappa(s d) with respect to the ScaleTransform s (kappa' r / s) and the ARD weights."""
    T = d2.dtype.type
    if family == SE:
        return -d2 * np.exp(-d2 / T(2))
    d = np.sqrt(d2)
    if family == MATERN12:
        return -d * np.exp(-d)
    if family == MATERN32:
        return -T(3) * d2 * np.exp(-T(math.sqrt(3.0)) * d)
    if family == MATERN52:
        s5 = T(math.sqrt(5.0)) * d
        return -(T(5) * d2 / T(3)) * (T(1) + s5) * np.exp(-s5)
    raise ValueError(family)


def logpdf_grad(k, mean, noise, X, y):
    """Gradient of logpdf(fx, y) (src/finite_gp_projection.jl:306-311) with respect to the hyper-parameters the device
    path carries -- what Zygote returns through the reference for `logpdf` in test/finite_gp_projection.jl:152-178 and
    in the training loops of examples/1-mauna-loa/script.jl:200-242.  Closed form:
    dL/dtheta = 1/2 tr((alpha alpha' - C^-1) dC/dtheta), dL/dm = alpha.  Returns a dict:
      variance (sigma_f^2), scale (ScaleTransform s) or ard (vector), linear_c, noise (scalar or per-point vector),
      mean_c (ConstMean) / mean_v (vector mean)."""
    dtype = X.dtype
    n = X.shape[0]
    m, C = mean_and_cov_fx(k, mean, noise, X)
    U = cholesky_upper(C)
    delta = np.asarray(y, dtype=dtype) - m
    alpha = _U_solve(U, _Ut_solve(U, delta))
    Vinv = _Ut_solve(U, np.eye(n, dtype=dtype))  # U^-T
    W = np.outer(alpha, alpha) - Vinv.T @ Vinv       # alpha alpha' - C^-1
    g = {}
    Kf = kernelmatrix(k, X)
    g["variance"] = 0.5 * np.sum(W * Kf) / k.variance
    Xt = k.apply_transform(X)
    if k.family == LINEAR:
        G = Xt @ Xt.T  # transformed inner products
        g["linear_c"] = 0.5 * k.variance * np.sum(W)
        if k.transform == T_SCALE:
            g["scale"] = 0.5 * k.variance * np.sum(W * G) * 2.0 / k.scale
        elif k.transform == T_ARD:
            v = np.asarray(k.ard, dtype=dtype)
            g["ard"] = np.array([0.5 * k.variance * 2.0 * v[d] * np.sum(W * np.outer(X[:, d], X[:, d])) for d in range(X.shape[1])])
    else:
        d2 = _pairwise_sqdist(Xt, Xt, "direct")
        np.fill_diagonal(d2, 0)
        KR = k.variance * _dkappa_r(k.family, d2)      # sigma_f^2 kappa'(r) r
        if k.transform == T_SCALE:
            g["scale"] = 0.5 * np.sum(W * KR) / k.scale
        elif k.transform == T_ARD:
            v = np.asarray(k.ard, dtype=dtype)
            with np.errstate(divide="ignore", invalid="ignore"):
                Q = np.where(d2 > 0, KR / d2, 0.0)       # sigma_f^2 kappa'(r) / r
            out = np.empty(X.shape[1], dtype=np.float64)
            for d in range(X.shape[1]):
                dd = X[:, d][:, None] - X[:, d][None, :]
                out[d] = 0.5 * np.sum(W * Q * (v[d] * dd * dd))
            g["ard"] = out
    g["noise"] = 0.5 * np.trace(W) if noise.kind == 0 else 0.5 * np.diag(W).copy()
    if mean.kind == 1:
        g["mean_c"] = np.sum(alpha)
    elif mean.kind == 2:
        g["mean_v"] = alpha.copy()
    return g


def posterior(k, mean, noise, X, y):
    """posterior(fx, y) src/exact_gpr_posterior.jl:29-35 -> data = (alpha, C(U), x, delta)."""
    m, C = mean_and_cov_fx(k, mean, noise, X)
    U = cholesky_upper(C)
    delta = np.asarray(y, dtype=X.dtype) - m
    alpha = _U_solve(U, _Ut_solve(U, delta))  # C \ delta  (potrs)
    return dict(alpha=alpha, U=U, x=X, delta=delta, k=k, mean=mean)


def post_mean_and_var(post, Xs, mean_s=None, noise_s: Optional[NoiseSpec] = None):
    """mean_and_var(::PosteriorGP, x*) src/exact_gpr_posterior.jl:85-90, plus the FiniteGP
    wrapper src/finite_gp_projection.jl:154-158 (adds diag(Sigma_y*)) when noise_s given."""
    k, X = post["k"], post["x"]
    Xs = np.asarray(Xs, dtype=X.dtype)
    Kxs = kernelmatrix(k, X, Xs)  # C_xcond_x  (N x M)
    ms = (mean_s or post["mean"]).vector(Xs.shape[0], X.dtype)
    m_post = ms + Kxs.T @ post["alpha"]
    v_post = kernelmatrix_diag(k, Xs) - diag_Xt_invA_X(post["U"], Kxs)
    if noise_s is not None:
        v_post = v_post + noise_s.diag(Xs.shape[0], X.dtype)
    return m_post, v_post


def post_mean_and_cov(post, Xs, mean_s=None):
    """mean_and_cov(::PosteriorGP, x*) src/exact_gpr_posterior.jl:78-83."""
    k, X = post["k"], post["x"]
    Xs = np.asarray(Xs, dtype=X.dtype)
    Kxs = kernelmatrix(k, X, Xs)
    ms = (mean_s or post["mean"]).vector(Xs.shape[0], X.dtype)
    m_post = ms + Kxs.T @ post["alpha"]
    C_post = kernelmatrix(k, Xs) - Xt_invA_X(post["U"], Kxs)
    return m_post, C_post


def post_logpdf(post, Xs, noise_s: NoiseSpec, Y, mean_s=None):
    """logpdf(f_post(x*, Sigma*), Y): the generic FiniteGP method src/finite_gp_projection.jl:306-311
    (matrix Y :313-318) with f = PosteriorGP, i.e. mean_and_cov from src/exact_gpr_posterior.jl:78-83
    plus Sigma* (src/finite_gp_projection.jl:133-136)."""
    dtype = post["x"].dtype
    Y = np.asarray(Y, dtype=dtype)
    m, C = post_mean_and_cov(post, Xs, mean_s)
    n = C.shape[0]
    C = C.copy()
    C[np.diag_indices(n)] += noise_s.diag(n, dtype)
    U = cholesky_upper(C)
    ld = dtype.type(logdet_chol(U))
    sq = tr_Xt_invA_X(U, Y - m) if Y.ndim == 1 else diag_Xt_invA_X(U, Y - m[:, None])
    return -((n * dtype.type(LOG2PI) + ld) + sq) / dtype.type(2)


def post_rand_from_Z(post, Xs, noise_s: NoiseSpec, Z, mean_s=None):
    """rand(rng, f_post(x*, Sigma*), S) src/finite_gp_projection.jl:233-237 with the caller's normals."""
    dtype = post["x"].dtype
    m, C = post_mean_and_cov(post, Xs, mean_s)
    n = C.shape[0]
    C = C.copy()
    C[np.diag_indices(n)] += noise_s.diag(n, dtype)
    U = cholesky_upper(C)
    Z = np.asarray(Z, dtype=dtype)
    return m + U.T @ Z if Z.ndim == 1 else m[:, None] + U.T @ Z


def rand_from_Z(k, mean, noise, X, Z):
    """rand(rng, fx, S) src/finite_gp_projection.jl:233-237 with the caller's normals Z[N,S]:
    m .+ C.U' * Z."""
    m, C = mean_and_cov_fx(k, mean, noise, X)
    U = cholesky_upper(C)
    Z = np.asarray(Z, dtype=X.dtype)
    if Z.ndim == 1:
        return m + U.T @ Z
    return m[:, None] + U.T @ Z


def update_chol(U11, C12, C22):
    """update_chol src/util/common_covmat_ops.jl:38-42."""
    U12 = _Ut_solve(U11, C12)
    U22 = cholesky_upper(C22 - U12.T @ U12)
    n1, n2 = U11.shape[0], U22.shape[0]
    U = np.zeros((n1 + n2, n1 + n2), dtype=U11.dtype)
    U[:n1, :n1] = U11
    U[:n1, n1:] = U12
    U[n1:, n1:] = U22
    return U


def posterior_sequential(post, noise2, X2, y2):
    """posterior(fx::FiniteGP{<:PosteriorGP}, y) src/exact_gpr_posterior.jl:46-56."""
    k, X1 = post["k"], post["x"]
    X2 = np.asarray(X2, dtype=X1.dtype)
    m2 = post["mean"].vector(X2.shape[0], X1.dtype)
    d2 = np.asarray(y2, dtype=X1.dtype) - m2
    C12 = kernelmatrix(k, X1, X2)
    C22 = kernelmatrix(k, X2)
    C22[np.diag_indices(X2.shape[0])] += noise2.diag(X2.shape[0], X1.dtype)
    U = update_chol(post["U"], C12, C22)
    delta = np.concatenate([post["delta"], d2])
    alpha = _U_solve(U, _Ut_solve(U, delta))
    return dict(alpha=alpha, U=U, x=np.concatenate([X1, X2], 0), delta=delta, k=k,
                mean=post["mean"])


# ---------------------------------------------------------------------------
# VFE / DTC  (src/sparse_approximations.jl)
# ---------------------------------------------------------------------------

def _compute_intermediates(k, mean, noise, X, y, Zind, jitter: NoiseSpec):
    """_compute_intermediates src/sparse_approximations.jl:289-305.  Returns (dtc, A)."""
    dtype = X.dtype
    n, M = X.shape[0], Zind.shape[0]
    if n != len(y):
        raise ValueError("DimensionMismatch")
    sy = noise.diag(n, dtype)
    chol_sy_U = np.sqrt(sy)  # _cholesky(Diagonal).U
    Kzz = kernelmatrix(k, Zind)
    Kzz[np.diag_indices(M)] += jitter.diag(M, dtype)
    U = cholesky_upper(Kzz)
    Kxz = kernelmatrix(k, X, Zind)  # cov(fx, fz)  N x M
    A = _Ut_solve(U, (Kxz / chol_sy_U[:, None]).T)  # M x N
    D = A @ A.T
    D[np.diag_indices(M)] += dtype.type(1)
    Lam_U = cholesky_upper(D)
    delta = (np.asarray(y, dtype=dtype) - mean.vector(n, dtype)) / chol_sy_U
    logdet_sy = dtype.type(np.sum(np.log(sy.astype(np.float64))))
    tmp = (logdet_sy + dtype.type(logdet_chol(Lam_U)) + np.sum(delta * delta)
           - np.sum(_Ut_solve(Lam_U, A @ delta) ** 2))
    dtc = -(n * dtype.type(LOG2PI) + tmp) / dtype.type(2)
    return dtc, A


def tr_Cf_invSy(k, noise, X):
    """tr_Cf_invSigma_y src/sparse_approximations.jl:307-313."""
    return np.sum(kernelmatrix_diag(k, X) / noise.diag(X.shape[0], X.dtype))


def dtc(k, mean, noise, X, y, Zind, jitter):
    """approx_log_evidence(::DTC) src/sparse_approximations.jl:282-286."""
    return _compute_intermediates(k, mean, noise, X, y, Zind, jitter)[0]


def elbo(k, mean, noise, X, y, Zind, jitter):
    """approx_log_evidence(::VFE)/elbo src/sparse_approximations.jl:248-254."""
    d, A = _compute_intermediates(k, mean, noise, X, y, Zind, jitter)
    return d - (tr_Cf_invSy(k, noise, X) - np.sum(A * A)) / X.dtype.type(2)


def vfe_posterior(k, mean, noise, X, y, Zind, jitter):
    """posterior(::VFE, fx, y) src/sparse_approximations.jl:58-75."""
    dtype = X.dtype
    n, M = X.shape[0], Zind.shape[0]
    U_y = np.sqrt(noise.diag(n, dtype))
    Kzz = kernelmatrix(k, Zind)
    Kzz[np.diag_indices(M)] += jitter.diag(M, dtype)
    U = cholesky_upper(Kzz)
    B = _Ut_solve(U, (kernelmatrix(k, X, Zind) / U_y[:, None]).T)  # B_ef  M x N
    b_y = (np.asarray(y, dtype=dtype) - mean.vector(n, dtype)) / U_y
    D = B @ B.T
    D[np.diag_indices(M)] += dtype.type(1)
    Lam_U = cholesky_upper(D)
    m_e = _U_solve(Lam_U, _Ut_solve(Lam_U, B @ b_y))
    alpha = _U_solve(U, m_e)
    return dict(m_e=m_e, Lam_U=Lam_U, U=U, alpha=alpha, b_y=b_y, z=Zind, k=k, mean=mean)


def vfe_mean_and_var(vp, Xs):
    """mean_and_var(::ApproxPosteriorGP, x*) src/sparse_approximations.jl:212-217."""
    k = vp["k"]
    Xs = np.asarray(Xs, dtype=vp["z"].dtype)
    A = _Ut_solve(vp["U"], kernelmatrix(k, vp["z"], Xs))
    m_post = vp["mean"].vector(Xs.shape[0], Xs.dtype) + A.T @ vp["m_e"]
    c_post = kernelmatrix_diag(k, Xs) - diag_At_A(A) + diag_Xt_invA_X(vp["Lam_U"], A)
    return m_post, c_post


def vfe_mean_and_cov(vp, Xs):
    """mean_and_cov(::ApproxPosteriorGP, x*) src/sparse_approximations.jl:205-210 (cov :187-190)."""
    k = vp["k"]
    Xs = np.asarray(Xs, dtype=vp["z"].dtype)
    A = _Ut_solve(vp["U"], kernelmatrix(k, vp["z"], Xs))
    m_post = vp["mean"].vector(Xs.shape[0], Xs.dtype) + A.T @ vp["m_e"]
    C_post = kernelmatrix(k, Xs) - A.T @ A + Xt_invA_X(vp["Lam_U"], A)
    return m_post, C_post


def vfe_cov_cross(vp, Xs, Ys):
    """cov(::ApproxPosteriorGP, x, y) src/sparse_approximations.jl:197-203."""
    k = vp["k"]
    A_zx = _Ut_solve(vp["U"], kernelmatrix(k, vp["z"], Xs))
    A_zy = _Ut_solve(vp["U"], kernelmatrix(k, vp["z"], Ys))
    return kernelmatrix(k, Xs, Ys) - A_zx.T @ A_zy + Xt_invA_Y(A_zx, vp["Lam_U"], A_zy)


# ---------------------------------------------------------------------------
# Synthetic workloads of SURVEY.md s8(d) -- identical bytes go to oracle and GPU
# ---------------------------------------------------------------------------

def make_config(cid: str, n: Optional[int] = None, dtype=None):
    """Returns dict(k, mean, noise, X, y[, Xs][, Z, jitter]) for configs C1..C5 (s8d)."""
    if cid == "C1":  # README.md:31-40
        rng = np.random.default_rng(20241)
        x = rng.random(10)
        return dict(k=KernelSpec(MATERN32), mean=MeanSpec(), noise=NoiseSpec(0, 0.001),
                    X=x[:, None].copy(), y=np.sin(x))
    spec = {
        "C2": (4096, 8, np.float64, SE, 0.1, 20242),
        "C3": (16384, 32, np.float32, MATERN32, 0.05, 20243),
        "C4": (65536, 64, np.float64, SE, 0.1, 20244),
        "C5": (1000000, 16, np.float32, SE, 0.1, 20245),
    }[cid]
    N, D, dt, fam, s2, seed = spec
    N = n or N
    dt = dtype or dt
    rng = np.random.default_rng(seed)
    X = rng.random((N, D))
    eps = rng.standard_normal(N)
    y = np.sin(2 * np.pi * X.mean(1)) + 0.3 * eps
    ell = math.sqrt(D) * 0.5
    if cid == "C3":
        k = KernelSpec(fam, 1.0, T_ARD, ard=((1.0 / ell) * (1.0 + 0.5 * np.arange(D) / D)).astype(dt))
    else:
        k = KernelSpec(fam, 1.0, T_SCALE, scale=1.0 / ell)
    out = dict(k=k, mean=MeanSpec(), noise=NoiseSpec(0, s2), X=X.astype(dt), y=y.astype(dt))
    if cid == "C3":
        out["Xs"] = np.random.default_rng(seed + 1).random((10000, D)).astype(dt)
    if cid == "C5":
        M = min(8192, max(8, N // 8))
        perm = np.random.default_rng(seed + 2).permutation(N)[:M]
        out["Z"] = out["X"][perm].copy()
        out["jitter"] = NoiseSpec(0, 1e-4 if dt == np.float32 else 1e-6)
    return out
