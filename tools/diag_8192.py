"""which of {persistent, bounded-CTA} v3 runs is right at N = 8192 (M = 8320), K = 512, S = 7?  compare both with fp64 matmul"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import agp_b200 as ag
eng = ag.engine()
for N in (8192, 4224, 6144, 8192):
    K, S = 512, 7
    M = N + 128
    g = torch.Generator(device="cuda").manual_seed(1)
    P = (torch.rand((K, M), generator=g, device="cuda", dtype=torch.float64) * 2 - 1).t()
    P = P * torch.logspace(-3, 2, M, device="cuda", dtype=torch.float64)[:, None]
    Pc = P.t().contiguous()
    C0 = torch.rand((N, M), generator=g, device="cuda", dtype=torch.float64)
    want = C0.t() - P @ P[:N].t()
    i = torch.arange(M, device="cuda")[:, None]; j = torch.arange(N, device="cuda")[None, :]
    low = (j // 64) * 64 < (i // 128) * 128 + 128
    for tag, env in (("persistent", {}), ("chunk4", {"AGP_OZAKI_CHUNK_TEST": "4"}), ("chunk16", {"AGP_OZAKI_CHUNK_TEST": "16"}), ("persistent2", {})):
        for k_, v_ in env.items(): os.environ[k_] = v_
        Cc = C0.clone()
        torch.cuda.synchronize()
        eng.check(eng.L.agp_debug_ozaki_syrk(eng.h, C.c_void_p(Cc.data_ptr()), M, C.c_void_p(Pc.data_ptr()), M, M, N, K, S, 1))
        torch.cuda.synchronize()
        for k_ in env: os.environ.pop(k_, None)
        got = Cc.t()
        err = (got - want).abs()
        bad = (err > 1e-6 * (want.abs() + 1)) & low
        untouched_ok = bool((got[~low] == C0.t()[~low]).all())
        nb = int(bad.sum())
        msg = ""
        if nb:
            idx = bad.nonzero()
            ti = torch.unique(idx[:, 0] // 128); tj = torch.unique(idx[:, 1] // 64)
            msg = " bad row tiles %s col strips %s" % (ti[:12].tolist(), tj[:12].tolist())
        print("N=%d %-11s max_err_low %.3e bad %d untouched_ok %s%s" % (N, tag, float(err[low].max()), nb, untouched_ok, msg), flush=True)
