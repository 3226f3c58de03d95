"""CPU model of the arithmetic of the tcgen05 int8-slice SYRK (abstractgps.jl_b200/csrc/umma_ozaki.cu): row-exponent
scaling, error-free 7-bit slicing (ozaki_slice_kernel), exact int32 accumulation per diagonal d = s + t (what the
stacked-B MMAs leave in TMEM), and the two fp64 recombinations of the epilogue -- plain Horner (AGP_OZAKI_EPI=0) and the
int32 pair pre-combination that is the default for S = 7, K <= 512 (AGP_OZAKI_EPI=1).  Checks the bounds the kernel relies
on: slices stay in [-64, 64], accumulators and pairs fit int32 for K <= 512 (and the pair does NOT fit at K = 1024, which is
why the host falls back), both recombinations agree with exact integer arithmetic to fp64 rounding, and the final update
meets the 2^-7S truncation bound."""
from fractions import Fraction

import numpy as np
import pytest


def slice_rows(P, S):
    """ozaki_rowscale_kernel + ozaki_slice_kernel: returns (q[S, m, K] int64, scale[m])."""
    mx = np.abs(P).max(1)
    e = np.where(mx > 0, np.frexp(mx)[1], 0)
    scale = np.ldexp(1.0, e)
    r = P / scale[:, None]
    q = np.empty((S,) + P.shape, dtype=np.int64)
    up = 64.0
    for s in range(S):
        qs = np.rint(r * up)
        r = r - qs / up          # exact in fp64 (fma(-q, dn, r) in the kernel)
        q[s] = qs.astype(np.int64)
        up *= 128.0
    return q, scale, r


def accumulators(qa, qb, S):
    """ACC_d[i, j] = sum_{s + t = d} sum_k qa[s, i, k] qb[t, j, k], d = 0 .. S-1 (only s + t <= S-1 is computed)."""
    acc = np.zeros((S, qa.shape[1], qb.shape[1]), dtype=np.int64)
    for s in range(S):
        for t in range(S - s):
            acc[s + t] += qa[s] @ qb[t].T
    return acc


def horner_plain(acc):
    S = acc.shape[0]
    v = acc[S - 1].astype(np.float64)
    for d in range(S - 2, -1, -1):
        v = v * (1.0 / 128.0) + acc[d].astype(np.float64)  # fma in the kernel; the product by 2^-7 is exact
    return v


def horner_pairs(acc):
    """the default epilogue: t_j = 128 ACC_2j + ACC_2j+1 in int32, then Horner over the pairs (+ the odd term)."""
    S = acc.shape[0]
    a32 = acc.astype(np.int32)
    pair = lambda d: (a32[d] * np.int32(128) + a32[d + 1])  # int32 arithmetic, wraps like the device would
    if S & 1:
        v = a32[S - 1].astype(np.float64)
        v = v * (1.0 / 128.0) + pair(S - 3).astype(np.float64)
        for d in range(S - 5, -1, -2):
            v = v * (1.0 / 16384.0) + pair(d).astype(np.float64)
    else:
        v = pair(S - 2).astype(np.float64)
        for d in range(S - 4, -1, -2):
            v = v * (1.0 / 16384.0) + pair(d).astype(np.float64)
    return v * (1.0 / 128.0)


def exact_value(acc, i, j):
    S = acc.shape[0]
    return sum(Fraction(int(acc[d, i, j]), 128 ** d) for d in range(S))


@pytest.mark.parametrize("S", [5, 6, 7, 8])
def test_slices_accumulators_and_both_epilogues(S):
    rng = np.random.default_rng(S)
    m, K = 48, 512
    P = rng.standard_normal((m, K)) * np.exp(rng.uniform(-6, 6, (m, 1)))
    q, scale, resid = slice_rows(P, S)
    assert np.abs(q).max() <= 64
    assert np.abs(resid).max() <= 2.0 ** (-7 * S)  # what the S slices leave behind, relative to the row scale
    acc = accumulators(q, q, S)
    for d in range(S):
        assert np.abs(acc[d]).max() <= (d + 1) * K * 64 * 64 < 2 ** 31
    v0, v1 = horner_plain(acc), horner_pairs(acc)
    for (i, j) in [(0, 0), (3, 17), (47, 1), (20, 20), (5, 44)]:
        ex = exact_value(acc, i, j)
        for v in (v0, v1):
            assert abs(Fraction(float(v[i, j])) - ex) <= abs(ex) * Fraction(1, 2 ** 51) + Fraction(1, 2 ** 60)
    # the update itself: scale_i scale_j / 4096 * v  vs  fp64 P P'
    got = v1 * (scale[:, None] / 4096.0) * scale[None, :]
    want = P @ P.T
    bound = (2.0 ** (-7 * S + 3)) * np.outer(scale, scale) * K + 4e-16 * np.abs(want) + 1e-300
    assert np.all(np.abs(got - want) <= bound)


@pytest.mark.parametrize("S", [5, 6, 7, 8])
def test_pair_bound_holds_at_k512_and_fails_at_k1024(S):
    """worst case: every slice entry +-64 with aligned signs.  K = 512: every pair fits int32 (exact result);
    K = 1024: the pair of diagonals (4, 5) wraps for S >= 6 -- the launcher must not use the pair epilogue there."""
    for K, must_fit in ((512, True), (1024, False)):
        q = np.full((S, 2, K), 64, dtype=np.int64)
        acc = accumulators(q, q, S)
        assert np.abs(acc).max() < 2 ** 31  # the accumulators themselves always fit
        v0, v1 = horner_plain(acc), horner_pairs(acc)
        fits = np.array_equal(v0, v1) or np.allclose(v0, v1, rtol=1e-15, atol=0)
        if must_fit:
            assert fits
        elif S >= 6:
            assert not fits


def test_umma_noswizzle_chunk_layout_halves():
    """ozaki_slice_kernel's bulk layout: byte offset of (row, k) inside a 4096-byte chunk; the cluster variant multicasts
    bytes [0, 2048) from CTA 0 and [2048, 4096) from CTA 1 and relies on these being rows 0-63 and 64-127."""
    def off(row, kbyte):
        g, r8, h = (row & 127) >> 3, row & 7, (kbyte >> 4) & 1
        return ((g * 2 + h) * 8 + r8) * 16 + (kbyte & 15)
    seen = set()
    for row in range(128):
        for kb in range(32):
            o = off(row, kb)
            assert (o < 2048) == (row < 64)
            seen.add(o)
    assert seen == set(range(4096))
