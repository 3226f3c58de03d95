"""Run `reps` fused fits (agp_fit: Gram + Cholesky + solves) of a synthetic N x D SqExponential fp64 problem through the
C ABI without importing torch -- the short command ncu wraps for per-kernel captures.
Usage: python tools/fit_once.py N D [reps] [dtype f64|f32]"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("agp_cabi", os.path.join(ROOT, "abstractgps.jl_b200", "_cabi.py"))
cabi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cabi)

N, D = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
f32 = len(sys.argv) > 4 and sys.argv[4] == "f32"
for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        C.CDLL(name, mode=C.RTLD_GLOBAL)
        break
    except OSError:
        continue
L = C.CDLL(cabi.LIB_PATH, mode=C.RTLD_GLOBAL)
for name, (res, args) in cabi.SIGNATURES.items():
    fn = getattr(L, name)
    fn.restype, fn.argtypes = res, args
h = C.c_void_p()
assert L.agp_init(C.byref(h), 0, None) == 0
rng = np.random.default_rng(0)
dt = np.float32 if f32 else np.float64
X = rng.random((N, D)).astype(dt)
y = (np.sin(2 * np.pi * X.mean(1)) + 0.3 * rng.standard_normal(N)).astype(dt)
ks = cabi.agp_kernel(0, 1, 1.0, 1.0 / (0.5 * np.sqrt(D)), 0.0, None)
ms = cabi.agp_mean(0, 0.0, None)
ns = cabi.agp_noise(0, 0.1, None)
lp = np.zeros(1, dtype=dt)
alpha = np.zeros(N, dtype=dt)
for _ in range(reps):
    rc = L.agp_fit(h, cabi.AGP_F32 if f32 else cabi.AGP_F64, C.byref(ks), C.byref(ms), C.byref(ns), cabi.AGP_POINT_MAJOR,
                   X.ctypes.data, N, D, y.ctypes.data, 1, lp.ctypes.data, alpha.ctypes.data, None)
    assert rc == 0, L.agp_last_error(h).decode()
tm = (C.c_double * 8)()
L.agp_last_timings(h, tm, 8)
print("logpdf %.12g total_ms %.3f chol_ms %.3f" % (float(lp[0]), tm[0], tm[3]))
