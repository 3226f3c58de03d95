"""TEST INFRASTRUCTURE: a stand-in for libagp.so that answers the C ABI's entry points with the CPU oracle, so that the
host mirror (abstractgps.jl_b200/api.py: argument marshalling, layouts, dtype handling, handle ownership, error mapping)
can be exercised without a GPU.  It is installed only by tests (tests/test_api_on_fake_lib.py); the product never
imports it -- the real library has no CPU path.

Every method takes exactly what api.py passes: ctypes byref() objects for the structs, c_void_p / int addresses for the
buffers, Python ints for sizes; outputs are written through the caller's pointers like the C library does."""
import ctypes as C

import numpy as np

from oracle import agp_ref as ref

OK, NOT_POSDEF, DIM, UNSUPPORTED, CUDA, NCCL, INVALID = range(7)


def _addr(p):
    if p is None:
        return None
    if isinstance(p, int):
        return p or None
    if hasattr(p, "_obj"):      # byref(c_void_p)
        return C.addressof(p._obj)
    return p.value


def _arr(p, shape, dtype, order="C"):
    a = _addr(p)
    if a is None:
        return None
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype=dtype)
    ct = C.c_double if np.dtype(dtype) == np.float64 else C.c_float
    flat = np.ctypeslib.as_array((ct * n).from_address(a))
    return flat.reshape(shape, order=order)


def _struct(p):
    return None if p is None else p._obj


class FakeLib:
    def __init__(self):
        self.err = b""
        self.info = 0
        self.posts = {}
        self.vposts = {}
        self.next_handle = 1000
        self.calls = []
        self.cfg = None

    # ---- helpers
    def _dt(self, code):
        return np.float64 if code == 1 else np.float32

    def _kernel(self, ks, D, dt):
        ks = _struct(ks)
        ard = None
        if ks.transform == 2:
            ard = np.array(_arr(ks.ard, (D,), dt))
        return ref.KernelSpec(ks.family, ks.variance, ks.transform, scale=ks.scale, ard=ard, linear_c=ks.linear_c)

    def _mean(self, ms, n, dt):
        ms = _struct(ms)
        if ms is None:
            return ref.MeanSpec()
        return ref.MeanSpec(ms.kind, ms.c, None if ms.kind != 2 else np.array(_arr(ms.v, (n,), dt)))

    def _noise(self, ns, n, dt):
        ns = _struct(ns)
        if ns is None:
            return ref.NoiseSpec()
        return ref.NoiseSpec(ns.kind, ns.s, None if ns.kind != 1 else np.array(_arr(ns.v, (n,), dt)))

    def _points(self, layout, X, n, D, dt):
        return np.array(_arr(X, (n, D), dt)) if layout == 0 else np.array(_arr(X, (D, n), dt)).T.copy()

    def _fail(self, code, msg, info=0):
        self.err, self.info = msg.encode(), info
        return code

    def _new(self, table, obj):
        self.next_handle += 1
        table[self.next_handle] = obj
        return self.next_handle

    def _h(self, p):
        return p.value if hasattr(p, "value") else int(p)

    # ---- context
    def agp_last_error(self, h):
        return self.err

    def agp_last_info(self, h):
        return self.info

    def agp_get_config(self, h, cfg):
        return OK

    def agp_set_config(self, h, cfg):
        return OK

    def agp_launch_count(self, h):
        return 0

    def agp_last_timings(self, h, buf, n):
        return 0

    # ---- Gram
    def agp_gram(self, h, code, ks, layout, X, n, D, Z, m, ns, K_out):
        self.calls.append("agp_gram")
        dt = self._dt(code)
        k = self._kernel(ks, D, dt)
        Xa = self._points(layout, X, n, D, dt)
        if _addr(Z) is None:
            K = ref.kernelmatrix(k, Xa)
            if _struct(ns) is not None:
                K[np.diag_indices(n)] += self._noise(ns, n, dt).diag(n, dt)
            _arr(K_out, (n, n), dt, "F")[...] = K
        else:
            Za = self._points(layout, Z, m, D, dt)
            _arr(K_out, (n, m), dt, "F")[...] = ref.kernelmatrix(k, Xa, Za)
        return OK

    # ---- fit
    def agp_fit(self, h, code, ks, ms, ns, layout, X, n, D, Y, S, lp_out, alpha_out, post_out):
        self.calls.append("agp_fit")
        dt = self._dt(code)
        k, mean, noise = self._kernel(ks, D, dt), self._mean(ms, n, dt), self._noise(ns, n, dt)
        Xa = self._points(layout, X, n, D, dt)
        Ya = np.array(_arr(Y, (n, S), dt, "F"))
        try:
            lp = ref.logpdf(k, mean, noise, Xa, Ya)
            post = ref.posterior(k, mean, noise, Xa, Ya[:, 0])
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "matrix is not positive definite", 1)
        if _addr(lp_out) is not None:
            _arr(lp_out, (S,), dt)[...] = lp
        if _addr(alpha_out) is not None:
            _arr(alpha_out, (n,), dt)[...] = post["alpha"]
        if post_out is not None:
            post["noise"] = noise
            post_out._obj.value = self._new(self.posts, post)
        return OK

    def agp_post_free(self, p):
        self.posts.pop(self._h(p), None)
        return OK

    def agp_post_n(self, p):
        return self.posts[self._h(p)]["x"].shape[0]

    def agp_post_logdet(self, p, out):
        out._obj.value = ref.logdet_chol(self.posts[self._h(p)]["U"])
        return OK

    def agp_post_factor_export(self, p, U_out):
        post = self.posts[self._h(p)]
        n = post["x"].shape[0]
        _arr(U_out, (n, n), post["x"].dtype, "F")[...] = post["U"]
        return OK

    def agp_post_solve_lower(self, p, B, nrhs, V_out):
        post = self.posts[self._h(p)]
        n, dt = post["x"].shape[0], post["x"].dtype
        Bm = np.array(_arr(B, (n, nrhs), dt, "F"))
        _arr(V_out, (n, nrhs), dt, "F")[...] = ref._Ut_solve(post["U"], Bm)
        return OK

    def _post_args(self, p, layout, Xs, M, ms):
        post = self.posts[self._h(p)]
        dt = post["x"].dtype
        D = post["x"].shape[1]
        Xa = self._points(layout, Xs, M, D, dt)
        mean_s = self._mean(ms, M, dt) if _struct(ms) is not None else None
        return post, dt, Xa, mean_s

    def agp_post_mean_var(self, p, layout, Xs, M, ms, ns, mean_out, var_out):
        self.calls.append("agp_post_mean_var")
        post, dt, Xa, mean_s = self._post_args(p, layout, Xs, M, ms)
        noise_s = self._noise(ns, M, dt) if _struct(ns) is not None else None
        m, v = ref.post_mean_and_var(post, Xa, mean_s, noise_s)
        if _addr(mean_out) is not None:
            _arr(mean_out, (M,), dt)[...] = m
        if _addr(var_out) is not None:
            _arr(var_out, (M,), dt)[...] = v
        return OK

    def agp_post_mean_cov(self, p, layout, Xs, M, ms, mean_out, cov_out):
        self.calls.append("agp_post_mean_cov")
        post, dt, Xa, mean_s = self._post_args(p, layout, Xs, M, ms)
        m, Cv = ref.post_mean_and_cov(post, Xa, mean_s)
        if _addr(mean_out) is not None:
            _arr(mean_out, (M,), dt)[...] = m
        if _addr(cov_out) is not None:
            _arr(cov_out, (M, M), dt, "F")[...] = Cv
        return OK

    def agp_post_logpdf(self, p, layout, Xs, M, ms, ns, Y, S, lp_out):
        self.calls.append("agp_post_logpdf")
        post, dt, Xa, mean_s = self._post_args(p, layout, Xs, M, ms)
        Ya = np.array(_arr(Y, (M, S), dt, "F"))
        try:
            _arr(lp_out, (S,), dt)[...] = ref.post_logpdf(post, Xa, self._noise(ns, M, dt), Ya, mean_s)
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "posterior covariance is not positive definite", 1)
        return OK

    def agp_post_rand(self, p, layout, Xs, M, ms, ns, Z, S, out):
        self.calls.append("agp_post_rand")
        post, dt, Xa, mean_s = self._post_args(p, layout, Xs, M, ms)
        Za = np.array(_arr(Z, (M, S), dt, "F"))
        try:
            _arr(out, (M, S), dt, "F")[...] = ref.post_rand_from_Z(post, Xa, self._noise(ns, M, dt), Za, mean_s)
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "posterior covariance is not positive definite", 1)
        return OK

    def agp_post_extend(self, p, layout, X2, N2, y2, ms, ns, alpha_out, post_out):
        self.calls.append("agp_post_extend")
        post = self.posts[self._h(p)]
        dt, D = post["x"].dtype, post["x"].shape[1]
        Xa = self._points(layout, X2, N2, D, dt)
        ya = np.array(_arr(y2, (N2,), dt))
        mean2 = self._mean(ms, N2, dt)
        # the oracle's sequential posterior subtracts post["mean"]; a vector mean arrives per call
        pm = dict(post)
        if mean2.kind == 2:
            pm["mean"] = mean2
        try:
            new = ref.posterior_sequential(pm, self._noise(ns, N2, dt), Xa, ya)
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "extended covariance is not positive definite", post["x"].shape[0] + 1)
        new["mean"] = post["mean"]
        if _addr(alpha_out) is not None:
            _arr(alpha_out, (new["x"].shape[0],), dt)[...] = new["alpha"]
        if post_out is not None:
            post_out._obj.value = self._new(self.posts, new)
        else:
            self.posts[self._h(p)] = new
        return OK

    def agp_post_logpdf_grad(self, p, grad_out, noise_diag_out):
        self.calls.append("agp_post_logpdf_grad")
        post = self.posts[self._h(p)]
        X = post["x"].astype(np.float64)
        k = post["k"]
        y = post["delta"].astype(np.float64) + post["mean"].vector(X.shape[0], np.float64)
        g = ref.logpdf_grad(k, post["mean"], ref.NoiseSpec(1, v=post["noise"].diag(X.shape[0], np.float64)), X, y)
        D = X.shape[1]
        out = np.ctypeslib.as_array(grad_out, shape=(5 + D,))
        out[:] = 0.0
        out[0] = g["variance"]
        out[1] = g.get("scale", 0.0)
        out[2] = g.get("linear_c", 0.0)
        out[3] = np.sum(g["noise"])
        out[4] = np.sum(ref._U_solve(post["U"], ref._Ut_solve(post["U"], post["delta"])))
        if "ard" in g:
            out[5:] = g["ard"]
        if _addr(noise_diag_out) is not None:
            _arr(noise_diag_out, (X.shape[0],), post["x"].dtype)[...] = g["noise"]
        return OK

    # ---- rand
    def agp_rand(self, h, code, ks, ms, ns, layout, X, n, D, Z, S, out):
        self.calls.append("agp_rand")
        dt = self._dt(code)
        k, mean, noise = self._kernel(ks, D, dt), self._mean(ms, n, dt), self._noise(ns, n, dt)
        Xa = self._points(layout, X, n, D, dt)
        Za = np.array(_arr(Z, (n, S), dt, "F"))
        try:
            _arr(out, (n, S), dt, "F")[...] = ref.rand_from_Z(k, mean, noise, Xa, Za)
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "matrix is not positive definite", 1)
        return OK

    # ---- VFE
    def _vfe(self, code, ks, ms, ns, layout, X, n, D, Zi, M, js, y):
        dt = self._dt(code)
        k, mean, noise = self._kernel(ks, D, dt), self._mean(ms, n, dt), self._noise(ns, n, dt)
        Xa = self._points(layout, X, n, D, dt)
        Za = self._points(layout, Zi, M, D, dt)
        return dt, k, mean, noise, Xa, Za, self._noise(js, M, dt), np.array(_arr(y, (n,), dt))

    def agp_vfe_elbo(self, h, code, ks, ms, ns, layout, X, n, D, Zi, M, js, y, elbo_out, dtc_out):
        self.calls.append("agp_vfe_elbo")
        dt, k, mean, noise, Xa, Za, jit, ya = self._vfe(code, ks, ms, ns, layout, X, n, D, Zi, M, js, y)
        try:
            _arr(elbo_out, (1,), dt)[0] = ref.elbo(k, mean, noise, Xa, ya, Za, jit)
            if _addr(dtc_out) is not None:
                _arr(dtc_out, (1,), dt)[0] = ref.dtc(k, mean, noise, Xa, ya, Za, jit)
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "K_zz + jitter is not positive definite", 1)
        return OK

    def agp_vfe_fit(self, h, code, ks, ms, ns, layout, X, n, D, Zi, M, js, y, post_out):
        self.calls.append("agp_vfe_fit")
        dt, k, mean, noise, Xa, Za, jit, ya = self._vfe(code, ks, ms, ns, layout, X, n, D, Zi, M, js, y)
        post_out._obj.value = self._new(self.vposts, ref.vfe_posterior(k, mean, noise, Xa, ya, Za, jit))
        return OK

    def agp_vfe_mean_var(self, p, layout, Xs, Ms, mean_out, var_out):
        self.calls.append("agp_vfe_mean_var")
        vp = self.vposts[self._h(p)]
        dt, D = vp["z"].dtype, vp["z"].shape[1]
        Xa = self._points(layout, Xs, Ms, D, dt)
        if vp["mean"].kind == 2:  # like the device handle: a vector (closure) mean is the host's business
            vp = dict(vp, mean=ref.MeanSpec())
        m, v = ref.vfe_mean_and_var(vp, Xa)
        _arr(mean_out, (Ms,), dt)[...] = m
        _arr(var_out, (Ms,), dt)[...] = v
        return OK

    def _vp(self, p):
        vp = self.vposts[self._h(p)]
        if vp["mean"].kind == 2:  # like the device handle: Zero/Const means only
            vp = dict(vp, mean=ref.MeanSpec())
        return vp

    def agp_vfe_mean_cov(self, p, layout, Xs, M, mean_out, cov_out):
        self.calls.append("agp_vfe_mean_cov")
        vp = self._vp(p)
        dt, D = vp["z"].dtype, vp["z"].shape[1]
        m, Cv = ref.vfe_mean_and_cov(vp, self._points(layout, Xs, M, D, dt))
        if _addr(mean_out) is not None:
            _arr(mean_out, (M,), dt)[...] = m
        if _addr(cov_out) is not None:
            _arr(cov_out, (M, M), dt, "F")[...] = Cv
        return OK

    def _vfe_factor(self, p, layout, Xs, M, ns):
        vp = self._vp(p)
        dt, D = vp["z"].dtype, vp["z"].shape[1]
        m, Cv = ref.vfe_mean_and_cov(vp, self._points(layout, Xs, M, D, dt))
        Cv = Cv.copy()
        Cv[np.diag_indices(M)] += self._noise(ns, M, dt).diag(M, dt)
        return dt, m, ref.cholesky_upper(Cv)

    def agp_vfe_post_logpdf(self, p, layout, Xs, M, ns, Y, S, lp_out):
        self.calls.append("agp_vfe_post_logpdf")
        try:
            dt, m, U = self._vfe_factor(p, layout, Xs, M, ns)
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "approximate posterior covariance is not positive definite", 1)
        Ya = np.array(_arr(Y, (M, S), dt, "F"))
        sq = ref.diag_Xt_invA_X(U, Ya - m[:, None])
        _arr(lp_out, (S,), dt)[...] = -((M * ref.LOG2PI + ref.logdet_chol(U)) + sq) / 2.0
        return OK

    def agp_vfe_post_rand(self, p, layout, Xs, M, ns, Z, S, out):
        self.calls.append("agp_vfe_post_rand")
        try:
            dt, m, U = self._vfe_factor(p, layout, Xs, M, ns)
        except np.linalg.LinAlgError:
            return self._fail(NOT_POSDEF, "approximate posterior covariance is not positive definite", 1)
        Za = np.array(_arr(Z, (M, S), dt, "F"))
        _arr(out, (M, S), dt, "F")[...] = m[:, None] + U.T @ Za
        return OK

    def agp_vfe_post_free(self, p):
        self.vposts.pop(self._h(p), None)
        return OK
