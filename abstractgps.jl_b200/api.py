"""Host-side mirror of the AbstractGPs.jl public surface for the dense hot path.

Julia is not available in this image, so the reference-facing host code is written in Python with
the reference's names and argument meaning (GP, f(x, s2)::FiniteGP, logpdf, posterior,
mean_and_var, rand, VFE/elbo ...).  Every method below forwards to ONE C-ABI entry point of
libagp.so (include/agp.h); no arithmetic on N-sized data happens on the host, and nothing falls back
to a CPU implementation.  Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import namedtuple
from typing import Optional

import numpy as np

from . import _cabi as cabi
from ._cabi import AGPError, DimensionMismatch, PosDefException  # noqa: F401

# ---------------------------------------------------------------------------------------------
# KernelFunctions.jl surface that AbstractGPs re-exports (src/AbstractGPs.jl:8)
# ---------------------------------------------------------------------------------------------
SE, MATERN12, MATERN32, MATERN52, LINEAR = range(5)


class Transform:
    pass


class ScaleTransform(Transform):
    def __init__(self, s: float):
        self.s = float(s)


class ARDTransform(Transform):
    def __init__(self, v):
        self.v = np.ascontiguousarray(v, dtype=np.float64).ravel()


class Kernel:
    """sigma_f^2 * (kappa o transform) -- the kernel set the engine implements on device."""

    def __init__(self, family, variance=1.0, transform: Optional[Transform] = None, c=0.0):
        self.family, self.variance, self.transform, self.c = family, float(variance), transform, float(c)

    def __rmul__(self, s):  # sigma^2 * k  (ScaledKernel)
        return Kernel(self.family, self.variance * float(s), self.transform, self.c)

    __mul__ = __rmul__

    def compose(self, t: Transform):  # k o t  (TransformedKernel)
        if self.transform is not None:
            t = _chain(self.transform, t)
        return Kernel(self.family, self.variance, t, self.c)

    __matmul__ = compose

    def __eq__(self, o):
        return isinstance(o, Kernel) and self._key() == o._key()

    def _key(self):
        t = self.transform
        tk = None if t is None else (("s", t.s) if isinstance(t, ScaleTransform) else ("a", tuple(t.v)))
        return (self.family, self.variance, tk, self.c)

    __hash__ = None


def _chain(outer: Transform, inner: Transform) -> Transform:
    """(k o outer) o inner applies inner first: x -> outer(inner(x)); both are diagonal scalings."""
    if isinstance(outer, ScaleTransform) and isinstance(inner, ScaleTransform):
        return ScaleTransform(outer.s * inner.s)
    ov = outer.v if isinstance(outer, ARDTransform) else outer.s
    iv = inner.v if isinstance(inner, ARDTransform) else inner.s
    return ARDTransform(np.asarray(ov) * np.asarray(iv))


def SqExponentialKernel():
    return Kernel(SE)


SEKernel = RBFKernel = GaussianKernel = SqExponentialKernel


def Matern12Kernel():
    return Kernel(MATERN12)


ExponentialKernel = Matern12Kernel


def Matern32Kernel():
    return Kernel(MATERN32)


def Matern52Kernel():
    return Kernel(MATERN52)


def LinearKernel(c: float = 0.0):
    return Kernel(LINEAR, c=c)


def TransformedKernel(k: Kernel, t: Transform):
    return k.compose(t)


def ScaledKernel(k: Kernel, s2: float):
    return s2 * k


def with_lengthscale(k: Kernel, ell):
    """with_lengthscale(k, l) = k o ScaleTransform(1/l); vector l -> ARDTransform(1 ./ l)."""
    if np.ndim(ell) == 0:
        return k.compose(ScaleTransform(1.0 / float(ell)))
    return k.compose(ARDTransform(1.0 / np.asarray(ell, dtype=np.float64)))


class ColVecs:
    """ColVecs(X): X is D x N, each COLUMN a point (KernelFunctions.ColVecs)."""

    def __init__(self, X):
        X = np.asarray(X)
        assert X.ndim == 2
        self.X = X

    def __len__(self):
        return self.X.shape[1]


class RowVecs:
    """RowVecs(X): X is N x D, each ROW a point (KernelFunctions.RowVecs)."""

    def __init__(self, X):
        X = np.asarray(X)
        assert X.ndim == 2
        self.X = X

    def __len__(self):
        return self.X.shape[0]


class _Points:
    """Engine view of an input collection: a C-contiguous [n, D] host array == D x N column-major
    == AGP_POINT_MAJOR for the ABI (so both wrappers cost at most one host transpose)."""

    def __init__(self, x):
        if isinstance(x, _Points):
            self.a = x.a
        elif isinstance(x, ColVecs):
            self.a = np.ascontiguousarray(x.X.T)
        elif isinstance(x, RowVecs):
            self.a = np.ascontiguousarray(x.X)
        else:
            v = np.asarray(x)
            if v.ndim != 1:
                raise TypeError("inputs must be a vector of reals, ColVecs(X) or RowVecs(X)")
            self.a = np.ascontiguousarray(v.reshape(-1, 1))
        if self.a.dtype not in (np.float32, np.float64):
            self.a = self.a.astype(np.float64)
        self.n, self.D = self.a.shape

    def astype(self, dt):
        p = _Points.__new__(_Points)
        p.a = np.ascontiguousarray(self.a, dtype=dt)
        p.n, p.D = self.n, self.D
        return p

    def julia_items(self):
        """what `map(f, x)` would iterate over (src/mean_function.jl:52-55)."""
        return self.a[:, 0] if self.D == 1 else self.a


def vcat(x, y):
    return _Points_from(np.concatenate([_Points(x).a, _Points(y).a], 0))


def _Points_from(a):
    p = _Points.__new__(_Points)
    p.a = np.ascontiguousarray(a)
    p.n, p.D = p.a.shape
    return p


# ---------------------------------------------------------------------------------------------
# mean functions (src/mean_function.jl)
# ---------------------------------------------------------------------------------------------
class MeanFunction:
    pass


class ZeroMean(MeanFunction):
    def vector(self, pts, dt):
        return np.zeros(pts.n, dtype=dt)

    def spec(self, pts, dt):
        return 0, 0.0, None


class ConstMean(MeanFunction):
    def __init__(self, c):
        self.c = float(c)

    def vector(self, pts, dt):
        return np.full(pts.n, self.c, dtype=dt)

    def spec(self, pts, dt):
        return 1, self.c, None


class CustomMean(MeanFunction):
    """arbitrary host closure: evaluated host-side and shipped as a vector."""

    def __init__(self, f):
        self.f = f

    def vector(self, pts, dt):
        return np.array([self.f(xi) for xi in pts.julia_items()], dtype=dt)

    def spec(self, pts, dt):
        return 2, 0.0, self.vector(pts, dt)


# ---------------------------------------------------------------------------------------------
# engine singleton
# ---------------------------------------------------------------------------------------------
class Engine:
    def __init__(self, device: Optional[int] = None):
        L = cabi.lib()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = L.agp_init(C.byref(h), device, None)
        if rc != cabi.AGP_OK:
            raise AGPError(rc, "agp_init failed on device %d (no CUDA device? there is no CPU fallback)" % device)
        self.L, self.h, self.device = L, h, device

    def check(self, rc):
        if rc == cabi.AGP_OK:
            return
        msg = self.L.agp_last_error(self.h).decode()
        if rc == cabi.AGP_ERR_NOT_POSDEF:
            raise PosDefException(int(self.L.agp_last_info(self.h)), msg)
        if rc == cabi.AGP_ERR_DIM_MISMATCH:
            raise DimensionMismatch(msg)
        raise AGPError(rc, msg)

    def timings(self):
        buf = (C.c_double * 8)()
        n = self.L.agp_last_timings(self.h, buf, 8)
        keys = ["total", "h2d", "gram", "cholesky", "solves", "d2h", "predict", "trailing"]
        return {k: buf[i] for i, k in enumerate(keys[:n])}

    def launch_count(self):
        return int(self.L.agp_launch_count(self.h))

    def set_memspace(self, m):
        self.check(self.L.agp_set_memspace(self.h, m))

    def get_config(self):
        c = cabi.agp_config()
        self.check(self.L.agp_get_config(self.h, C.byref(c)))
        return c

    def set_config(self, **kw):
        """tile_nb / fp64_mode / lookahead / ozaki_slices of the live context"""
        c = self.get_config()
        for k_, v in kw.items():
            setattr(c, k_, v)
        self.check(self.L.agp_set_config(self.h, C.byref(c)))


_engine = None


def engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


def _kernel_struct(k: Kernel, dt, keep):
    ks = cabi.agp_kernel()
    ks.family, ks.variance, ks.linear_c, ks.scale = k.family, k.variance, k.c, 1.0
    t = k.transform
    if t is None:
        ks.transform = 0
    elif isinstance(t, ScaleTransform):
        ks.transform, ks.scale = 1, t.s
    else:
        v = np.ascontiguousarray(t.v, dtype=dt)
        keep.append(v)
        ks.transform, ks.ard = 2, v.ctypes.data
    return ks


def _mean_struct(spec, keep):
    kind, c, v = spec
    ms = cabi.agp_mean()
    ms.kind, ms.c = kind, c
    if v is not None:
        keep.append(v)
        ms.v = v.ctypes.data
    return ms


def _noise_struct(s2, n, dt, keep):
    ns = cabi.agp_noise()
    if np.ndim(s2) == 0:
        ns.kind, ns.s = 0, float(s2)
    else:
        v = np.ascontiguousarray(s2, dtype=dt).ravel()
        if v.shape[0] != n:
            raise DimensionMismatch("noise vector has length %d, expected %d" % (v.shape[0], n))
        keep.append(v)
        ns.kind, ns.v = 1, v.ctypes.data
    return ns


def _vfe_mean_cov(p: "ApproxPosteriorGP", pts: _Points):
    """mean_and_cov(::ApproxPosteriorGP, x*) (src/sparse_approximations.jl:205-210) through agp_vfe_mean_cov (EXPERIMENTAL)."""
    eng = engine()
    pts = pts.astype(p.dtype)
    m = np.empty(pts.n, dtype=p.dtype)
    Cv = np.empty((pts.n, pts.n), dtype=p.dtype, order="F")
    eng.check(eng.L.agp_vfe_mean_cov(p.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, cabi.ptr(m), cabi.ptr(Cv)))
    if isinstance(p.prior.mean, CustomMean):
        m = m + p.prior.mean.vector(pts, p.dtype)
    return m, Cv


def _vfe_post_logpdf(fx: "FiniteGP", y):
    """logpdf(f_approx_post(x*, s2), y) on the device (agp_vfe_post_logpdf, EXPERIMENTAL)."""
    eng = engine()
    p = fx.f
    dt = p.dtype
    pts = fx.x.astype(dt)
    Y = np.asarray(y, dtype=dt)
    vec = Y.ndim == 1
    Yf = Y.reshape(-1, 1) if vec else Y
    if Yf.shape[0] != pts.n:
        raise DimensionMismatch("length(fx) = %d but y has %d rows" % (pts.n, Yf.shape[0]))
    if isinstance(p.prior.mean, CustomMean):  # the handle knows Zero/Const means: a closure mean is removed here
        Yf = Yf - p.prior.mean.vector(pts, dt)[:, None]
    keep = []
    ns = _noise_struct(fx.s2, pts.n, dt, keep)
    lp = np.empty(Yf.shape[1], dtype=dt)
    for s0 in range(0, Yf.shape[1], 128):
        s1 = min(Yf.shape[1], s0 + 128)
        Yc = np.asfortranarray(Yf[:, s0:s1])
        eng.check(eng.L.agp_vfe_post_logpdf(p.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, C.byref(ns), cabi.ptr(Yc),
                                            s1 - s0, cabi.ptr(lp[s0:s1])))
    return lp[0] if vec else lp


def _vfe_post_rand_from_normals(fx: "FiniteGP", Z, squeeze=False):
    eng = engine()
    p = fx.f
    dt = p.dtype
    pts = fx.x.astype(dt)
    Z = np.asfortranarray(np.asarray(Z, dtype=dt).reshape(pts.n, -1))
    keep = []
    ns = _noise_struct(fx.s2, pts.n, dt, keep)
    out = np.empty_like(Z, order="F")
    eng.check(eng.L.agp_vfe_post_rand(p.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, C.byref(ns), cabi.ptr(Z), Z.shape[1],
                                      cabi.ptr(out)))
    if isinstance(p.prior.mean, CustomMean):
        out = out + p.prior.mean.vector(pts, dt)[:, None]
    return out[:, 0] if squeeze else out


# ---------------------------------------------------------------------------------------------
# GP / FiniteGP / PosteriorGP  (src/base_gp.jl, src/finite_gp_projection.jl, src/exact_gpr_posterior.jl)
# ---------------------------------------------------------------------------------------------
default_s2 = 1e-18  # src/finite_gp_projection.jl:17


class AbstractGP:
    def __call__(self, x, s2=default_s2, obsdim=None):
        """(f::AbstractGP)(x...) src/finite_gp_projection.jl:32; (f)(X::AbstractMatrix, args...; obsdim) :33-37 --
        obsdim = 1: rows are points (RowVecs), obsdim = 2: columns are points (ColVecs); a bare matrix without obsdim
        is ambiguous (the reference deprecated its default) and is rejected."""
        if obsdim is not None:
            X = np.asarray(x)
            if X.ndim != 2 or obsdim not in (1, 2):
                raise TypeError("obsdim applies to a matrix of inputs and is 1 (rows) or 2 (columns)")
            x = RowVecs(X) if obsdim == 1 else ColVecs(X)
        return FiniteGP(self, x, s2)


class GP(AbstractGP):
    """GP(kernel) / GP(mean, kernel) / GP(c::Real, kernel)  (src/base_gp.jl:57-64)."""

    def __init__(self, *args):
        if len(args) == 1:
            mean, kernel = ZeroMean(), args[0]
        else:
            mean, kernel = args
            if isinstance(mean, (int, float, np.integer, np.floating)):  # GP(c::Real, kernel), src/base_gp.jl:64
                mean = ConstMean(mean)
            elif not isinstance(mean, MeanFunction):
                mean = CustomMean(mean)
        if not isinstance(kernel, Kernel):
            raise TypeError("kernel must be one of the device-supported KernelFunctions kernels")
        self.mean, self.kernel = mean, kernel


class FiniteGP:
    """FiniteGP(f, x, Sigma_y) (src/finite_gp_projection.jl:7-21): scalar s2 -> Fill, vector -> Diagonal."""

    def __init__(self, f, x, s2=default_s2):
        self.f, self.x = f, _Points(x)
        if np.ndim(s2) == 2:
            raise AGPError(cabi.AGP_ERR_UNSUPPORTED, "dense Sigma_y is outside the device hot path (SURVEY s8a)")
        self.s2 = s2
        self.dtype = np.result_type(self.x.a.dtype, np.float32)

    def __len__(self):
        return self.x.n

    @property
    def Sigma_y_diag(self):
        return np.full(self.x.n, self.s2, dtype=self.dtype) if np.ndim(self.s2) == 0 else np.asarray(self.s2, self.dtype)


DeviceData = namedtuple("DeviceData", "alpha C x delta")


class DeviceCholesky:
    """The `C` field of PosteriorGP.data: a handle to the device-resident factor (boundary #2,
    src/util/common_covmat_ops.jl).  `.U` exports the upper factor like `C.U` in the reference."""

    def __init__(self, eng: Engine, handle, dtype):
        self.eng, self.h, self.dtype = eng, handle, dtype

    def __del__(self):
        try:
            if self.h:
                self.eng.L.agp_post_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def n(self):
        return int(self.eng.L.agp_post_n(self.h))

    @property
    def U(self):
        out = np.empty((self.n, self.n), dtype=self.dtype, order="F")
        self.eng.check(self.eng.L.agp_post_factor_export(self.h, cabi.ptr(out)))
        return out

    def logdet(self):
        v = C.c_double()
        self.eng.check(self.eng.L.agp_post_logdet(self.h, C.byref(v)))
        return v.value

    def solve_lower(self, B):
        """U' \\ B"""
        B = np.asarray(B, dtype=self.dtype)
        vec = B.ndim == 1
        Bf = np.asfortranarray(B.reshape(self.n, -1))
        out = np.empty_like(Bf, order="F")
        self.eng.check(self.eng.L.agp_post_solve_lower(self.h, cabi.ptr(Bf), Bf.shape[1], cabi.ptr(out)))
        return out[:, 0] if vec else out


# the operator API of src/util/common_covmat_ops.jl on a device factor
def Xt_invA_X(A: DeviceCholesky, X):  # :54-58
    V = A.solve_lower(X)
    return float(np.sum(V * V)) if V.ndim == 1 else V.T @ V


def Xt_invA_Y(X, A: DeviceCholesky, Y):  # :60
    return A.solve_lower(X).T @ A.solve_lower(Y)


def diag_Xt_invA_X(A: DeviceCholesky, X):  # :90
    V = A.solve_lower(X)
    return np.array([np.sum(V * V)]) if V.ndim == 1 else np.sum(V * V, axis=0)


def tr_Xt_invA_X(A: DeviceCholesky, X):  # :101
    V = A.solve_lower(X)
    return float(np.sum(V * V))


class PosteriorGP(AbstractGP):
    """PosteriorGP(prior, data=(alpha, C, x, delta)) (src/exact_gpr_posterior.jl:1-4)."""

    def __init__(self, prior, data):
        self.prior, self.data = prior, data


def _prior_of(fx: FiniteGP) -> GP:
    f = fx.f
    if not isinstance(f, GP):
        raise TypeError("expected a FiniteGP over a prior GP")
    return f


def _fit(fx: FiniteGP, Y, want_post: bool, want_alpha: bool):
    eng = engine()
    f = _prior_of(fx)
    dt = fx.dtype
    pts = fx.x.astype(dt)
    Y = np.asarray(Y)
    vec = Y.ndim == 1
    Yf = np.asfortranarray(Y.reshape(-1, 1) if vec else Y, dtype=dt)
    if Yf.shape[0] != pts.n:
        raise DimensionMismatch("length(fx) = %d but Y has %d rows" % (pts.n, Yf.shape[0]))
    S = Yf.shape[1]
    keep = []
    ks = _kernel_struct(f.kernel, dt, keep)
    ms = _mean_struct(f.mean.spec(pts, dt), keep)
    ns = _noise_struct(fx.s2, pts.n, dt, keep)
    lp = np.empty(S, dtype=dt)
    alpha = np.empty(pts.n, dtype=dt) if want_alpha else None
    post = C.c_void_p()
    # one call whatever S is: the library carries 128 columns through the factorisation and solves the rest against the
    # same factor (ONE Gram + ONE Cholesky)
    rc = eng.L.agp_fit(eng.h, cabi.dtype_code(dt), C.byref(ks), C.byref(ms), C.byref(ns), cabi.AGP_POINT_MAJOR,
                       cabi.ptr(pts.a), pts.n, pts.D, cabi.ptr(Yf), S, cabi.ptr(lp),
                       cabi.ptr(alpha) if want_alpha else None, C.byref(post) if want_post else None)
    eng.check(rc)
    lpv = lp[0] if vec else lp
    if not want_post:
        return lpv, None
    delta = Yf[:, 0] - f.mean.vector(pts, dt)
    data = DeviceData(alpha=alpha, C=DeviceCholesky(eng, post, dt), x=fx.x, delta=delta)
    return lpv, PosteriorGP(f, data)


def logpdf(fx: FiniteGP, y):
    """logpdf(fx, y) (src/finite_gp_projection.jl:306-311); matrix y -> per-column values."""
    if isinstance(fx.f, PosteriorGP):
        return _post_logpdf(fx, y)
    if isinstance(fx.f, ApproxPosteriorGP):
        return _vfe_post_logpdf(fx, y)
    if not isinstance(fx.f, GP):
        raise AGPError(cabi.AGP_ERR_UNSUPPORTED, "logpdf of a FiniteGP over %s is outside the device hot path"
                       % type(fx.f).__name__)
    return _fit(fx, y, False, False)[0]


def _post_args(fx: FiniteGP):
    p: PosteriorGP = fx.f
    dt = p.data.C.dtype
    pts = fx.x.astype(dt)
    if pts.D != p.data.x.D:
        raise DimensionMismatch("test points have D=%d, training points D=%d" % (pts.D, p.data.x.D))
    keep = []
    ms = _mean_struct(p.prior.mean.spec(pts, dt), keep)
    ns = _noise_struct(fx.s2, pts.n, dt, keep)
    return p, dt, pts, ms, ns, keep


def _post_logpdf(fx: FiniteGP, y):
    """logpdf(f_post(x*, s2), y) (src/finite_gp_projection.jl:306-318 over src/exact_gpr_posterior.jl:78-83):
    posterior covariance, noise add, Cholesky and the quadratic form all on the device (agp_post_logpdf)."""
    eng = engine()
    p, dt, pts, ms, ns, keep = _post_args(fx)
    Y = np.asarray(y, dtype=dt)
    vec = Y.ndim == 1
    Yf = Y.reshape(-1, 1) if vec else Y
    if Yf.shape[0] != pts.n:
        raise DimensionMismatch("length(fx) = %d but y has %d rows" % (pts.n, Yf.shape[0]))
    S = Yf.shape[1]
    lp = np.empty(S, dtype=dt)
    for s0 in range(0, S, 128):
        s1 = min(S, s0 + 128)
        Yc = np.asfortranarray(Yf[:, s0:s1])
        eng.check(eng.L.agp_post_logpdf(p.data.C.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, C.byref(ms),
                                        C.byref(ns), cabi.ptr(Yc), s1 - s0, cabi.ptr(lp[s0:s1])))
    return lp[0] if vec else lp


def _post_rand_from_normals(fx: FiniteGP, Z, squeeze=False):
    """rand(f_post(x*, s2), S) (src/finite_gp_projection.jl:233-240) through agp_post_rand."""
    eng = engine()
    p, dt, pts, ms, ns, keep = _post_args(fx)
    Z = np.asfortranarray(np.asarray(Z, dtype=dt).reshape(pts.n, -1))
    out = np.empty_like(Z, order="F")
    eng.check(eng.L.agp_post_rand(p.data.C.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, C.byref(ms), C.byref(ns),
                                  cabi.ptr(Z), Z.shape[1], cabi.ptr(out)))
    return out[:, 0] if squeeze else out


def loglikelihood(fx: FiniteGP, Y):  # src/finite_gp_projection.jl:304
    return np.sum(logpdf(fx, Y))


def posterior(fx, y=None, *rest):
    """posterior(fx, y) (src/exact_gpr_posterior.jl:29-35); posterior(vfe, fx, y) (src/sparse_approximations.jl:58-75);
    posterior(fx::FiniteGP{<:PosteriorGP}, y) sequential conditioning (src/exact_gpr_posterior.jl:46-56)."""
    if isinstance(fx, VFE):
        return _vfe_posterior(fx, y, rest[0])
    if isinstance(fx.f, PosteriorGP):
        return _posterior_sequential(fx, y)
    return _fit(fx, y, True, True)[1]


def fit(fx: FiniteGP, y):
    """Fused logpdf + posterior from ONE Gram and ONE factorisation (the reference does two:
    src/finite_gp_projection.jl:307-308 and src/exact_gpr_posterior.jl:30-31)."""
    return _fit(fx, y, True, True)


def logpdf_grad(fx: FiniteGP, y):
    """EXPERIMENTAL (device path not yet validated): (logpdf, gradient dict) of logpdf(fx, y) w.r.t. the kernel variance,
    ScaleTransform s / ARDTransform v, LinearKernel c, the noise (scalar or per-point) and the mean (constant or vector)
    -- the cotangents Zygote returns through the reference (test/finite_gp_projection.jl:152-178).  One fit, then
    agp_post_logpdf_grad on its factor."""
    lp, post = _fit(fx, y, True, True)
    eng = engine()
    f = post.prior
    dt = post.data.C.dtype
    D = post.data.x.D
    g = np.zeros(5 + D, dtype=np.float64)
    per_point = np.ndim(fx.s2) != 0
    nd = np.empty(len(fx), dtype=dt) if (per_point or isinstance(f.mean, CustomMean)) else None
    eng.check(eng.L.agp_post_logpdf_grad(post.data.C.h, g.ctypes.data_as(C.POINTER(C.c_double)), cabi.ptr(nd)))
    k = f.kernel
    out = {"variance": g[0]}
    if isinstance(k.transform, ScaleTransform):
        out["scale"] = g[1]
    elif isinstance(k.transform, ARDTransform):
        out["ard"] = g[5:5 + D].copy()
    if k.family == LINEAR:
        out["linear_c"] = g[2]
    out["noise"] = nd.astype(np.float64) if per_point else g[3]
    if isinstance(f.mean, ConstMean):
        out["mean_c"] = g[4]
    elif isinstance(f.mean, CustomMean):
        out["mean_v"] = post.data.alpha.astype(np.float64)
    return lp, out


def _post_call(p: PosteriorGP, pts: _Points, s2, want_var=True, want_cov=False):
    eng = engine()
    dt = p.data.C.dtype
    pts = pts.astype(dt)
    if pts.D != p.data.x.D:
        raise DimensionMismatch("test points have D=%d, training points D=%d" % (pts.D, p.data.x.D))
    keep = []
    ms = _mean_struct(p.prior.mean.spec(pts, dt), keep)
    mean = np.empty(pts.n, dtype=dt)
    if want_cov:
        cov = np.empty((pts.n, pts.n), dtype=dt, order="F")
        eng.check(eng.L.agp_post_mean_cov(p.data.C.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, C.byref(ms),
                                          cabi.ptr(mean), cabi.ptr(cov)))
        if s2 is not None:
            cov[np.diag_indices(pts.n)] += np.asarray(s2, dtype=dt)
        return mean, cov
    var = np.empty(pts.n, dtype=dt) if want_var else None
    ns = _noise_struct(s2, pts.n, dt, keep) if s2 is not None else None
    eng.check(eng.L.agp_post_mean_var(p.data.C.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, C.byref(ms),
                                      C.byref(ns) if ns is not None else None, cabi.ptr(mean), cabi.ptr(var)))
    return mean, var


def _gram(f: GP, pts: _Points, pts2: Optional[_Points], s2, dt):
    eng = engine()
    keep = []
    pts = pts.astype(dt)
    ks = _kernel_struct(f.kernel, dt, keep)
    ns = _noise_struct(s2, pts.n, dt, keep) if s2 is not None else None
    if pts2 is None:
        K = np.empty((pts.n, pts.n), dtype=dt, order="F")
        rc = eng.L.agp_gram(eng.h, cabi.dtype_code(dt), C.byref(ks), cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, pts.D,
                            None, 0, C.byref(ns) if ns is not None else None, cabi.ptr(K))
    else:
        pts2 = pts2.astype(dt)
        if pts2.D != pts.D:
            raise DimensionMismatch("inputs have different dimensionality")
        K = np.empty((pts.n, pts2.n), dtype=dt, order="F")
        rc = eng.L.agp_gram(eng.h, cabi.dtype_code(dt), C.byref(ks), cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, pts.D,
                            cabi.ptr(pts2.a), pts2.n, None, cabi.ptr(K))
    eng.check(rc)
    return K


def kernelmatrix(k: Kernel, x, y=None):
    """KernelFunctions.kernelmatrix(k, x[, y]) -- what cov(f, x[, y]) forwards to (src/base_gp.jl:70,74)."""
    pts = _Points(x)
    dt = np.result_type(pts.a.dtype, np.float32)
    return _gram(GP(k), pts, None if y is None else _Points(y), None, dt)


def kernelmatrix_diag(k: Kernel, x):
    """KernelFunctions.kernelmatrix_diag(k, x) -- what var(f, x) forwards to (src/base_gp.jl:72)."""
    return var(GP(k), x)


def mean_vector(m: "MeanFunction", x):
    """mean_vector(m, x) (src/mean_function.jl:27,40,52-55)."""
    pts = _Points(x)
    return m.vector(pts, np.result_type(pts.a.dtype, np.float32))


def _need_x(f, x, name):
    """`mean(f::AbstractGP)` & co. are not defined on purpose (src/abstract_gp.jl:66-87): Julia's ErrorException."""
    if x is None and isinstance(f, AbstractGP):
        raise RuntimeError("`%s(f::AbstractGP)` is not defined (on purpose!).\nPlease provide an `AbstractVector` of locations "
                           "`x` at which you wish to compute your %s vector%s, and call `%s(f(x))`" %
                           (name, name, "s" if name.startswith("mean_and") else "", name))


def mean(f, x=None):
    """mean(fx) (src/finite_gp_projection.jl:53) / mean(f, x) (src/abstract_gp.jl:19)."""
    if isinstance(f, FiniteGP):
        return mean(f.f, f.x)
    _need_x(f, x, "mean")
    pts = _Points(x)
    if isinstance(f, GP):
        return f.mean.vector(pts, np.result_type(pts.a.dtype, np.float32))
    if isinstance(f, PosteriorGP):
        return _post_call(f, pts, None, want_var=False)[0]
    if isinstance(f, ApproxPosteriorGP):
        return _vfe_mean_var(f, pts)[0]
    raise TypeError(type(f))


def cov(f, x=None, z=None):
    """cov(fx) (src/finite_gp_projection.jl:96); cov(f, x[, z]) (src/base_gp.jl:70,74); cov(fx, gx) (:177-180)."""
    if isinstance(f, FiniteGP) and isinstance(x, FiniteGP):
        assert f.f is x.f
        return cov(f.f, f.x, x.x)
    if isinstance(f, FiniteGP):
        if isinstance(f.f, GP):
            return _gram(f.f, f.x, None, f.s2, f.dtype)
        return mean_and_cov(f)[1]
    _need_x(f, x, "cov")
    pts = _Points(x)
    if isinstance(f, GP):
        dt = np.result_type(pts.a.dtype, np.float32)
        return _gram(f, pts, None if z is None else _Points(z), None, dt)
    if isinstance(f, PosteriorGP) and z is None:
        return _post_call(f, pts, None, want_cov=True)[1]
    if isinstance(f, PosteriorGP):  # cov(f_post, x, z) src/exact_gpr_posterior.jl:72-76
        Cx = cov(f.prior, f.data.x, pts)
        Cz = cov(f.prior, f.data.x, _Points(z))
        return cov(f.prior, pts, _Points(z)) - Xt_invA_Y(Cx, f.data.C, Cz)
    if isinstance(f, ApproxPosteriorGP):
        if z is None:  # cov(f_approx_post, x) src/sparse_approximations.jl:187-190
            return _vfe_mean_cov(f, pts)[1]
        # cov(f_approx_post, x, z) (:197-203): the off-diagonal block of the covariance at [x; z]
        zp = _Points(z)
        Cxz = _vfe_mean_cov(f, vcat(pts, zp))[1]
        return np.asfortranarray(Cxz[:pts.n, pts.n:])
    raise TypeError(type(f))


def var(f, x=None):
    """var(fx) (src/finite_gp_projection.jl:114-117) / var(f, x) (src/base_gp.jl:72)."""
    if isinstance(f, FiniteGP):
        return mean_and_var(f)[1]
    _need_x(f, x, "var")
    pts = _Points(x)
    if isinstance(f, GP):
        dt = np.result_type(pts.a.dtype, np.float32)
        k = f.kernel
        if k.family != LINEAR:
            return np.full(pts.n, k.variance, dtype=dt)
        a = pts.a.astype(dt)
        if isinstance(k.transform, ScaleTransform):
            a = a * k.transform.s
        elif isinstance(k.transform, ARDTransform):
            a = a * k.transform.v.astype(dt)
        return (k.variance * ((a * a).sum(1) + k.c)).astype(dt)
    if isinstance(f, PosteriorGP):
        return _post_call(f, pts, None)[1]
    if isinstance(f, ApproxPosteriorGP):
        return _vfe_mean_var(f, pts)[1]
    raise TypeError(type(f))


def mean_and_var(f, x=None):
    """mean_and_var(fx) (src/finite_gp_projection.jl:154-158) -> mean_and_var(f_post, x*)
    (src/exact_gpr_posterior.jl:85-90) + diag(Sigma_y)."""
    if isinstance(f, FiniteGP):
        if isinstance(f.f, PosteriorGP):
            return _post_call(f.f, f.x, f.s2)
        if isinstance(f.f, ApproxPosteriorGP):
            m, v = _vfe_mean_var(f.f, f.x)
            return m, v + f.Sigma_y_diag.astype(v.dtype)
        return mean(f), var(f.f, f.x) + f.Sigma_y_diag
    _need_x(f, x, "mean_and_var")
    return mean(f, x), var(f, x)


def mean_and_cov(f, x=None):
    """mean_and_cov(fx) (src/finite_gp_projection.jl:133-136) / (f_post, x*) (src/exact_gpr_posterior.jl:78-83)."""
    if isinstance(f, FiniteGP):
        if isinstance(f.f, PosteriorGP):
            return _post_call(f.f, f.x, f.Sigma_y_diag, want_cov=True)
        if isinstance(f.f, ApproxPosteriorGP):
            m, Cv = _vfe_mean_cov(f.f, f.x)
            Cv[np.diag_indices(len(f))] += f.Sigma_y_diag.astype(Cv.dtype)
            return m, Cv
        return mean(f), cov(f)
    _need_x(f, x, "mean_and_cov")
    if isinstance(f, PosteriorGP):
        return _post_call(f, _Points(x), None, want_cov=True)
    if isinstance(f, ApproxPosteriorGP):
        return _vfe_mean_cov(f, _Points(x))
    return mean(f, x), cov(f, x)


Normal = namedtuple("Normal", "mu sigma")


def marginals(fx: FiniteGP):
    """marginals(fx) = Normal.(m, sqrt.(c)) (src/finite_gp_projection.jl:203-206) as arrays."""
    m, c = mean_and_var(fx)
    return Normal(m, np.sqrt(c))


def rand(*args):
    """rand([rng,] fx[, S]) (src/finite_gp_projection.jl:233-240): m .+ C.U' * randn(rng, n, S).
    The normals come from the caller's numpy Generator so the stream stays host-defined."""
    args = list(args)
    rng = args.pop(0) if not isinstance(args[0], FiniteGP) else np.random.default_rng()
    fx = args.pop(0)
    S = args.pop(0) if args else None
    if not isinstance(fx.f, (GP, PosteriorGP, ApproxPosteriorGP)):
        raise AGPError(cabi.AGP_ERR_UNSUPPORTED, "sampling from a FiniteGP over %s is outside the device hot path"
                       % type(fx.f).__name__)
    eng = engine()
    dt = fx.f.data.C.dtype if isinstance(fx.f, PosteriorGP) else (fx.f.dtype if isinstance(fx.f, ApproxPosteriorGP) else fx.dtype)
    pts = fx.x.astype(dt)
    ns_cols = 1 if S is None else int(S)
    Z = np.asfortranarray(rng.standard_normal((pts.n, ns_cols)).astype(dt))
    return rand_from_normals(fx, Z, squeeze=S is None)


def rand_from_normals(fx: FiniteGP, Z, squeeze=False):
    if isinstance(fx.f, PosteriorGP):
        return _post_rand_from_normals(fx, Z, squeeze)
    if isinstance(fx.f, ApproxPosteriorGP):
        return _vfe_post_rand_from_normals(fx, Z, squeeze)
    eng = engine()
    f = _prior_of(fx)
    dt = fx.dtype
    pts = fx.x.astype(dt)
    Z = np.asfortranarray(np.asarray(Z, dtype=dt).reshape(pts.n, -1))
    keep = []
    ks = _kernel_struct(f.kernel, dt, keep)
    ms = _mean_struct(f.mean.spec(pts, dt), keep)
    ns = _noise_struct(fx.s2, pts.n, dt, keep)
    out = np.empty_like(Z, order="F")
    eng.check(eng.L.agp_rand(eng.h, cabi.dtype_code(dt), C.byref(ks), C.byref(ms), C.byref(ns), cabi.AGP_POINT_MAJOR,
                             cabi.ptr(pts.a), pts.n, pts.D, cabi.ptr(Z), Z.shape[1], cabi.ptr(out)))
    return out[:, 0] if squeeze else out


def _posterior_sequential(fx: FiniteGP, y):
    p: PosteriorGP = fx.f
    eng = engine()
    dt = p.data.C.dtype
    pts = fx.x.astype(dt)
    y = np.ascontiguousarray(y, dtype=dt)
    if y.shape[0] != pts.n:
        raise DimensionMismatch("length(fx) != length(y)")
    keep = []
    ms = _mean_struct(p.prior.mean.spec(pts, dt), keep)
    ns = _noise_struct(fx.s2, pts.n, dt, keep)
    n1 = p.data.C.n
    alpha = np.empty(n1 + pts.n, dtype=dt)
    h = C.c_void_p()
    eng.check(eng.L.agp_post_extend(p.data.C.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, cabi.ptr(y), C.byref(ms),
                                    C.byref(ns), cabi.ptr(alpha), C.byref(h)))
    delta = np.concatenate([p.data.delta, y - p.prior.mean.vector(pts, dt)])
    # a NEW device factor: like the reference, the posterior that was conditioned on stays usable
    data = DeviceData(alpha=alpha, C=DeviceCholesky(eng, h, dt), x=vcat(p.data.x, pts), delta=delta)
    return PosteriorGP(p.prior, data)


# ---------------------------------------------------------------------------------------------
# VFE (src/sparse_approximations.jl)
# ---------------------------------------------------------------------------------------------
class VFE:
    """VFE(fz) (src/sparse_approximations.jl:1-12)."""

    def __init__(self, fz: FiniteGP):
        self.fz = fz


class DTC(VFE):
    """DTC(fz) (src/sparse_approximations.jl:14-23): the same optimal approximate posterior as VFE (`posterior(::Union{VFE,DTC}, ...)`
    :58), but `approx_log_evidence` is the DTC objective (:282-286) -- the second output of agp_vfe_elbo."""


def inducing_points(p):
    """inducing_points(f_post_approx) (src/sparse_approximations.jl:219)."""
    return p.approx.fz.x


class ApproxPosteriorGP(AbstractGP):
    """ApproxPosteriorGP(approx, prior, data) (src/sparse_approximations.jl:25-29); `data` lives on the device behind
    the handle, the observations are kept host-side for `update_posterior`."""

    def __init__(self, approx, prior, handle, dtype, x=None, y=None, s2=None):
        self.approx, self.prior, self.h, self.dtype = approx, prior, handle, dtype
        self.x, self.y, self.s2 = x, y, s2

    def __del__(self):
        try:
            if self.h:
                engine().L.agp_vfe_post_free(self.h)
                self.h = None
        except Exception:
            pass


def _vfe_args(vfe: VFE, fx: FiniteGP, y):
    if vfe.fz.f is not fx.f:
        raise AssertionError("vfe.fz.f === fx.f")  # src/sparse_approximations.jl:59,249
    f = _prior_of(fx)
    dt = fx.dtype
    pts, z = fx.x.astype(dt), vfe.fz.x.astype(dt)
    y = np.ascontiguousarray(y, dtype=dt)
    if y.shape[0] != pts.n:
        raise DimensionMismatch("the dimension of the projected GP (here: %d) must equal the number of targets "
                                "(here: %d)" % (pts.n, y.shape[0]))
    keep = [y]
    ks = _kernel_struct(f.kernel, dt, keep)
    ms = _mean_struct(f.mean.spec(pts, dt), keep)
    ns = _noise_struct(fx.s2, pts.n, dt, keep)
    js = _noise_struct(vfe.fz.s2, z.n, dt, keep)
    return f, dt, pts, z, y, ks, ms, ns, js, keep


def approx_log_evidence(vfe: VFE, fx: FiniteGP, y, return_dtc=False):
    """elbo (src/sparse_approximations.jl:248-254)."""
    eng = engine()
    f, dt, pts, z, y, ks, ms, ns, js, keep = _vfe_args(vfe, fx, y)
    out = np.empty(2, dtype=dt)
    eng.check(eng.L.agp_vfe_elbo(eng.h, cabi.dtype_code(dt), C.byref(ks), C.byref(ms), C.byref(ns), cabi.AGP_POINT_MAJOR,
                                 cabi.ptr(pts.a), pts.n, pts.D, cabi.ptr(z.a), z.n, C.byref(js), cabi.ptr(y),
                                 cabi.ptr(out[0:1]), cabi.ptr(out[1:2])))
    if return_dtc:
        return out[0], out[1]
    return out[1] if isinstance(vfe, DTC) else out[0]


def elbo(vfe: VFE, fx: FiniteGP, y):
    """elbo(vfe::VFE, fx, y) = approx_log_evidence(vfe, fx, y) (src/sparse_approximations.jl:254); VFE only."""
    if isinstance(vfe, DTC):
        raise TypeError("elbo is defined for VFE; use approx_log_evidence for DTC")
    return approx_log_evidence(vfe, fx, y)


def _vfe_posterior(vfe: VFE, fx: FiniteGP, y):
    eng = engine()
    f, dt, pts, z, y, ks, ms, ns, js, keep = _vfe_args(vfe, fx, y)
    h = C.c_void_p()
    eng.check(eng.L.agp_vfe_fit(eng.h, cabi.dtype_code(dt), C.byref(ks), C.byref(ms), C.byref(ns), cabi.AGP_POINT_MAJOR,
                                cabi.ptr(pts.a), pts.n, pts.D, cabi.ptr(z.a), z.n, C.byref(js), cabi.ptr(y), C.byref(h)))
    return ApproxPosteriorGP(vfe, f, h, dt, x=fx.x, y=np.array(y), s2=fx.s2)


def update_posterior(p: ApproxPosteriorGP, a: FiniteGP, y=None):
    """update_posterior(f_post_approx, fx, y) -- new observations, same pseudo-points (src/sparse_approximations.jl:87-119)
    -- and update_posterior(f_post_approx, fz) -- pseudo-points appended (:130-176).  The reference patches its host
    factors with rank-1 / block updates; here the posterior is re-formed by the streamed device fit on the concatenated
    observations / inducing set (one pass over the data, O(N M^2) like the first fit), which yields the same posterior
    (test/sparse_approximations.jl:27-85 compares the two routes to atol 1e-5)."""
    if a.f is not p.prior:
        raise AssertionError("f_post_approx.prior === fx.f")  # :92 / :131
    if y is not None:
        y = np.asarray(y, dtype=p.dtype)
        if y.shape[0] != len(a):
            raise DimensionMismatch("length(fx) != length(y)")
        old_fx = FiniteGP(p.prior, p.x, p.s2)
        s2 = np.concatenate([old_fx.Sigma_y_diag.astype(p.dtype), a.Sigma_y_diag.astype(p.dtype)])
        return _vfe_posterior(p.approx, FiniteGP(p.prior, vcat(p.x, a.x), s2), np.concatenate([p.y, y]))
    fz_old = p.approx.fz
    fz_new = FiniteGP(p.prior, vcat(fz_old.x, a.x), fz_old.s2 if np.ndim(fz_old.s2) == 0 else
                      np.concatenate([np.asarray(fz_old.s2), a.Sigma_y_diag]))
    return _vfe_posterior(type(p.approx)(fz_new), FiniteGP(p.prior, p.x, p.s2), p.y)


def _vfe_mean_var(p: ApproxPosteriorGP, pts: _Points):
    eng = engine()
    pts = pts.astype(p.dtype)
    m = np.empty(pts.n, dtype=p.dtype)
    v = np.empty(pts.n, dtype=p.dtype)
    eng.check(eng.L.agp_vfe_mean_var(p.h, cabi.AGP_POINT_MAJOR, cabi.ptr(pts.a), pts.n, cabi.ptr(m), cabi.ptr(v)))
    if isinstance(p.prior.mean, CustomMean):  # the handle carries Zero/Const means only: a closure is evaluated here
        m = m + p.prior.mean.vector(pts, p.dtype)
    return m, v
