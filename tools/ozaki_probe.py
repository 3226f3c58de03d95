"""One-shot probe of the tcgen05 int8-slice SYRK kernel's epilogue variants and of the outer panel width.
Runs in a few seconds on one B200 (no torch import): python tools/ozaki_probe.py > gpurun_out/ozaki_probe.json

Part A: agp_debug_ozaki_syrk on a synthetic M x 512 panel (C = -P P', lower) under AGP_OZAKI_EPI = 0..4
        (0 shipped epilogue, 1 int32 pair pre-combination, 2-4 timing-only variants that skip parts of the epilogue).
Part B: a full C4h fit (N = 32768, D = 64) with tile_nb = 512 and 1024, and with AGP_OZAKI_EPI = 0 / 1.
Timings are wall clock around synchronous library calls (ms scale), best of `reps`."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util

spec = importlib.util.spec_from_file_location("agp_cabi", os.path.join(ROOT, "abstractgps.jl_b200", "_cabi.py"))
cabi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cabi)

out = {}
t00 = time.perf_counter()


class _Mock:  # PROBE_MOCK=1: exercise the script's own logic without a GPU (every call returns 0)
    def __getattr__(self, name):
        def f(*a):
            return b"mock" if name == "agp_last_error" else 0
        return f


MOCK = os.environ.get("PROBE_MOCK") == "1"
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = C.CDLL(name, mode=C.RTLD_GLOBAL)
        break
    except OSError:
        continue
if rt is None:
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so", mode=C.RTLD_GLOBAL)
L = C.CDLL(cabi.LIB_PATH, mode=C.RTLD_GLOBAL)
for name, (res, args) in cabi.SIGNATURES.items():
    fn = getattr(L, name)
    fn.restype, fn.argtypes = res, args
if MOCK:
    L, rt = _Mock(), _Mock()
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]
H2D, D2H = 1, 2


def save():
    out["total_s"] = time.perf_counter() - t00
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ozaki_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


def ck(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %d" % (what, rc))


h = C.c_void_p()
ck(L.agp_init(C.byref(h), 0, None), "agp_init")


def err():
    return L.agp_last_error(h).decode()


# ---------------- part A
M, K, S = int(os.environ.get("PROBE_M", 24576)), 512, 7
reps = 2
rng = np.random.default_rng(0)
P = np.asfortranarray(rng.standard_normal((M, K)) * np.exp(rng.uniform(-3, 3, (M, 1))))
dP, dC = C.c_void_p(), C.c_void_p()
ck(rt.cudaMalloc(C.byref(dP), P.nbytes), "cudaMalloc P")
ck(rt.cudaMalloc(C.byref(dC), M * M * 8), "cudaMalloc C")
if MOCK:
    dP.value = dC.value = 4096
ck(rt.cudaMemcpy(dP, P.ctypes.data, P.nbytes, H2D), "H2D")
out["setup_s"] = time.perf_counter() - t00


def syrk(mode, n_cols):
    os.environ["AGP_OZAKI_EPI"] = str(mode)
    best = 1e9
    for _ in range(reps):
        ck(rt.cudaMemset(dC, 0, M * M * 8), "memset")
        rt.cudaDeviceSynchronize()
        t0 = time.perf_counter()
        rc = L.agp_debug_ozaki_syrk(h, dC, M, dP, M, M, n_cols, K, S, 1)
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError("syrk mode %d: %s" % (mode, err()))
        best = min(best, dt)
    return best * 1e3


def sample():
    cols = [0, M // 3, M - 129]
    res = []
    for c in cols:
        buf = np.empty(M - c, dtype=np.float64)
        ck(rt.cudaMemcpy(buf.ctypes.data, C.c_void_p(dC.value + (c * M + c) * 8), buf.nbytes, D2H), "D2H")
        res.append(buf)
    return cols, res


fixed = syrk(0, 128)  # slicing + workspace + a 1-strip update: the mode-independent part
out["partA"] = {"M": M, "K": K, "S": S, "fixed_ms": fixed, "modes": {}}
t_full = syrk(0, M)
cols, ref = sample()
want = [-(P[c:, :] @ P[c, :]) for c in cols]
scale = [np.abs(P[c:, :]) @ np.abs(P[c, :]) for c in cols]
out["partA"]["modes"]["0"] = {"ms": t_full, "syrk_ms": t_full - fixed,
                              "max_rel_err_vs_fp64": float(max(np.max(np.abs(r - w) / s) for r, w, s in zip(ref, want, scale)))}
flops = 2.0 * K * (M * (M + 1) / 2)
out["partA"]["modes"]["0"]["fp64_equiv_tflops"] = flops / ((t_full - fixed) * 1e-3) / 1e12
save()
for mode in (1, 3, 2, 4, 7, 5, 6):  # 5: MMA-only main loop, 6: operand traffic only, 7: TMEM reads without conversions
    t = syrk(mode, M)
    d = {"ms": t, "syrk_ms": t - fixed}
    if mode == 1:
        _, got = sample()
        d["max_rel_err_vs_fp64"] = float(max(np.max(np.abs(r - w) / s) for r, w, s in zip(got, want, scale)))
        d["max_rel_diff_vs_mode0"] = float(max(np.max(np.abs(r - w) / s) for r, w, s in zip(got, ref, scale)))
    d["fp64_equiv_tflops"] = flops / (d["syrk_ms"] * 1e-3) / 1e12
    out["partA"]["modes"][str(mode)] = d
    save()

def cluster_part():
    """variants of the default (v3) kernel: CTA pairs with multicast A, 4 / 8 epilogue warps; timing-only modes 3 (no
    epilogue), 5 (MMA only), 6 (operand traffic only).  Run LAST: a protocol bug traps and kills the CUDA context,
    everything above is already saved."""
    os.environ["AGP_OZAKI_EPI"] = "1"
    os.environ["AGP_OZAKI_CLUSTER"] = "1"
    os.environ.pop("AGP_OZAKI_EPIWARPS", None)
    syrk(1, M)
    _, base = sample()
    res = {}
    for tag, cl, ew in (("epiwarps4", "1", "4"), ("cluster2", "2", "8"), ("cluster2_epiwarps4", "2", "4")):
        for epi in (1, 3, 5, 6):
            os.environ["AGP_OZAKI_CLUSTER"], os.environ["AGP_OZAKI_EPIWARPS"] = cl, ew
            t = syrk(epi, M)
            d = {"ms": t, "syrk_ms": t - fixed, "fp64_equiv_tflops": flops / ((t - fixed) * 1e-3) / 1e12}
            if epi == 1:
                _, got = sample()
                d["max_abs_diff_vs_default"] = float(max(np.max(np.abs(r - w)) for r, w in zip(got, base)))
            res["%s_epi%d" % (tag, epi)] = d
            out["partA"]["variants"] = res
            save()
    os.environ["AGP_OZAKI_CLUSTER"] = "1"
    os.environ.pop("AGP_OZAKI_EPIWARPS", None)


run_cluster_last = os.environ.get("PROBE_CLUSTER") == "1"
if not run_cluster_last:
    rt.cudaFree(dC)
    rt.cudaFree(dP)

# ---------------- part B: full fit, C4h
if os.environ.get("PROBE_FIT", "1") == "1":
    N, D = 32768, 64
    X = rng.random((N, D))
    y = np.sin(2 * np.pi * X.mean(1)) + 0.3 * rng.standard_normal(N)
    ks = cabi.agp_kernel(0, 1, 1.0, 1.0 / (0.5 * np.sqrt(D)), 0.0, None)  # SE o ScaleTransform
    ms = cabi.agp_mean(0, 0.0, None)
    ns = cabi.agp_noise(0, 0.1, None)
    lp = np.zeros(1)
    out["partB"] = {}
    for tag, nb, epi in (("nb512_epi0", 512, 0), ("nb1024_epi0", 1024, 0), ("nb512_epi1", 512, 1), ("nb256_epi0", 256, 0)):
        os.environ["AGP_OZAKI_EPI"] = str(epi)
        cfg = cabi.agp_config()
        ck(L.agp_get_config(h, C.byref(cfg)), "get_config")
        cfg.tile_nb = nb
        ck(L.agp_set_config(h, C.byref(cfg)), "set_config")
        best = 1e9
        for it in range(3):
            t0 = time.perf_counter()
            rc = L.agp_fit(h, cabi.AGP_F64, C.byref(ks), C.byref(ms), C.byref(ns), cabi.AGP_POINT_MAJOR, X.ctypes.data, N, D,
                           y.ctypes.data, 1, lp.ctypes.data, None, None)
            dt = time.perf_counter() - t0
            if rc != 0:
                out["partB"][tag] = {"error": err()}
                break
            if it > 0:
                best = min(best, dt)
        else:
            tm = (C.c_double * 8)()
            L.agp_last_timings(h, tm, 8)
            out["partB"][tag] = {"wall_ms": best * 1e3, "device_ms": tm[0], "cholesky_ms": tm[3], "logpdf": float(lp[0])}
        save()
save()
if run_cluster_last:
    cluster_part()
print(json.dumps(out, indent=1))
