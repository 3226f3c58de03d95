"""tcgen05 int8-sliced (Ozaki) fp64 trailing update vs. a float64 reference computed with torch on the
same device (a floating-point kernel, so the reference is fp64 matmul; the exact-integer part of the
scheme is additionally checked on integer-valued inputs, where the result must be bit-exact)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(ag, M, N, K, S, lower, seed=0, integer=False):
    import torch
    eng = ag.engine()
    g = torch.Generator(device="cuda").manual_seed(seed)
    if integer:
        P = torch.randint(-60, 61, (K, M), generator=g, device="cuda").to(torch.float64).t()  # column-major M x K view
    else:
        P = (torch.rand((K, M), generator=g, device="cuda", dtype=torch.float64) * 2 - 1).t()
        P = P * torch.logspace(-3, 2, M, device="cuda", dtype=torch.float64)[:, None]  # very different row scales
    lda = M
    Cm = torch.rand((N, M + 5), generator=g, device="cuda", dtype=torch.float64).t()  # column-major, ldc = M + 5
    ldc = M + 5
    C0 = Cm.clone()
    Pc = P.t().contiguous()      # storage of the column-major M x K matrix
    Cc = Cm.t().contiguous()     # storage of the column-major (M+5) x N matrix
    rc = eng.L.agp_debug_ozaki_syrk(eng.h, C.c_void_p(Cc.data_ptr()), ldc, C.c_void_p(Pc.data_ptr()), lda, M, N, K, S, int(lower))
    eng.check(rc)
    got = Cc.t()[:M, :N]
    want = C0[:M, :N] - P @ P[:N].t()
    return got.cpu().numpy(), want.cpu().numpy(), C0[:M, :N].cpu().numpy(), (P.abs().max(1).values).cpu().numpy()


@pytest.mark.parametrize("M,N,K,S", [(128, 64, 64, 8), (256, 256, 128, 8), (384, 320, 256, 7), (1000, 704, 256, 8), (640, 640, 512, 6)])
def test_ozaki_syrk_matches_fp64(ag, M, N, K, S):
    got, want, c0, rmax = _run(ag, M, N, K, S, lower=False)
    # error model: slicing truncation is relative to the row-scale products (2^-7S), plus the fp64
    # rounding of C itself and of the reference matmul
    scale = np.outer(rmax, rmax[:N]) * K
    tol = {8: 1e-15, 7: 2e-13, 6: 3e-11}[S]
    bound = tol * scale + 4e-16 * (np.abs(c0) + np.abs(want) + scale)
    assert np.all(np.abs(got - want) <= bound), float((np.abs(got - want) / bound).max())


def test_ozaki_exact_on_integers(ag):
    got, want, _, _ = _run(ag, 256, 256, 128, 8, lower=False, integer=True)
    assert np.array_equal(got, want)  # every product and sum is exact in int32 / fp64


def test_ozaki_lower_only_leaves_upper_tiles(ag):
    got, want, c0, _ = _run(ag, 512, 512, 128, 8, lower=True)
    i, j = np.indices(got.shape)
    low = (j // 64) * 64 < (i // 128) * 128 + 128   # tiles the kernel owns
    assert np.allclose(got[low], want[low], rtol=0, atol=1e-9)
    assert np.array_equal(got[~low], c0[~low])       # tiles entirely above the diagonal are untouched


@pytest.mark.parametrize("N,K,S", [(128, 128, 7), (1024, 256, 7), (2048, 512, 8), (2176, 512, 6)])
def test_ozaki_persistent_lower_with_border(ag, N, K, S):
    """the persistent (v2) kernel on the shape the Cholesky uses: M = N + 128 border rows, lower tiles only"""
    got, want, c0, rmax = _run(ag, N + 128, N, K, S, lower=True, seed=3)
    i, j = np.indices(got.shape)
    low = (j // 64) * 64 < (i // 128) * 128 + 128
    scale = np.outer(rmax, rmax[:N]) * K
    tol = {8: 1e-15, 7: 2e-13, 6: 3e-11}[S]
    bound = tol * scale + 4e-16 * (np.abs(c0) + np.abs(want) + scale)
    assert np.all(np.abs(got - want)[low] <= bound[low]), float((np.abs(got - want) / bound)[low].max())
    assert np.array_equal(got[~low], c0[~low])
