"""N>1 host logic on CPU: two gloo processes run the SAME schedule the CUDA path uses (block-column-cyclic
ownership, owner factors + solves the panel, panel broadcast, local trailing update with the gathered-row
mapping, all-reduce of logdet, distributed backward substitution with alpha broadcast) in NumPy and
must reproduce the dense Cholesky / logpdf of the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TILE = 16  # small tile so the test is quick; the mapping code is tile-size agnostic


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import scipy.linalg as sla
    import agp_b200  # noqa: F401
    from agp_b200.dist import owner_of_block, local_blocks, gather_row_for_local_col
    from agp_b200 import _cabi
    from oracle import agp_ref as ref

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lib = _cabi.lib()
    rng = np.random.default_rng(0)
    X = rng.random((n, 2))
    y = np.sin(X.sum(1))
    ks = ref.KernelSpec(ref.SE, 1.0, ref.T_SCALE, scale=2.0)
    K = ref.kernelmatrix(ks, X) + 0.1 * np.eye(n)
    nt = n // TILE
    mine = local_blocks(nt, rank, world)
    for j in mine:  # the C helper agrees with the Python mirror (1 x Q grid)
        assert lib.agp_bc_owner(0, j, 1, world) == rank == owner_of_block(j, world)
    # local storage: all rows (+ one border row = delta') of my column blocks
    Lloc = {j: np.vstack([K[:, j * TILE:(j + 1) * TILE], y[None, j * TILE:(j + 1) * TILE]]) for j in mine}
    logdet = 0.0
    for kk in range(nt):
        owner = owner_of_block(kk, world)
        rows_below = n + 1 - (kk + 1) * TILE
        panel = torch.zeros(rows_below, TILE, dtype=torch.float64)
        if owner == rank:
            blk = Lloc[kk]
            Lkk = np.linalg.cholesky(blk[kk * TILE:(kk + 1) * TILE])
            blk[kk * TILE:(kk + 1) * TILE] = Lkk
            logdet += 2 * np.log(np.diag(Lkk)).sum()
            below = sla.solve_triangular(Lkk, blk[(kk + 1) * TILE:].T, lower=True).T  # A21 L11^-T
            blk[(kk + 1) * TILE:] = below
            panel = torch.from_numpy(below.copy())
        dist.broadcast(panel, src=owner)
        P = panel.numpy()
        trailing = [j for j in mine if j > kk]
        if not trailing:
            continue
        j0 = trailing[0]
        for n_local in range(len(trailing) * TILE):  # column-by-column to exercise the gather mapping
            j = trailing[n_local // TILE]
            prow = gather_row_for_local_col(n_local, j0, kk, world, TILE)
            assert prow == (j - (kk + 1)) * TILE + n_local % TILE
            col = Lloc[j][:, n_local % TILE]
            col[(kk + 1) * TILE:] -= P @ P[prow]
    # logdet all-reduce, v from the border rows, distributed backward substitution
    t = torch.tensor([logdet])
    dist.all_reduce(t)
    v = np.zeros(n)
    for j in mine:
        v[j * TILE:(j + 1) * TILE] = Lloc[j][n]
    tv = torch.from_numpy(v.copy())
    dist.all_reduce(tv)
    sq = float((tv.numpy() ** 2).sum())
    r = v.copy()
    alpha = np.zeros(n)
    for i in range(nt - 1, -1, -1):
        a = torch.zeros(TILE, dtype=torch.float64)
        if owner_of_block(i, world) == rank:
            Lii = Lloc[i][i * TILE:(i + 1) * TILE]
            a = torch.from_numpy(sla.solve_triangular(Lii, r[i * TILE:(i + 1) * TILE], lower=True, trans="T"))
        dist.broadcast(a, src=owner_of_block(i, world))
        alpha[i * TILE:(i + 1) * TILE] = a.numpy()
        for j in mine:
            if j < i:
                r[j * TILE:(j + 1) * TILE] -= Lloc[j][i * TILE:(i + 1) * TILE].T @ a.numpy()
    lp = -0.5 * (n * np.log(2 * np.pi) + t.item() + sq)
    lp_ref = ref.logpdf(ks, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y)
    pr = ref.posterior(ks, ref.MeanSpec(), ref.NoiseSpec(0, 0.1), X, y)
    ok = abs(lp - lp_ref) <= 1e-10 * abs(lp_ref) and np.allclose(alpha, pr["alpha"], rtol=1e-8, atol=1e-10)
    out.put((rank, bool(ok), float(lp), float(lp_ref)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_block_cyclic_schedule_world(world):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 7 * TILE, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res


def _vfe_worker(rank, world, port, out):
    """VFE with the data dimension sharded over ranks and ONE all-reduce of (D, b, scalars) -- the schedule of
    vfe_core in csrc/engine.cu -- must reproduce the oracle's elbo."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import scipy.linalg as sla
    from oracle import agp_ref as ref

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    n, m, d = 501, 40, 3
    X = rng.random((n, d))
    y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n)
    Z = X[:m].copy()
    ks = ref.KernelSpec(ref.SE, 1.2, ref.T_SCALE, scale=1.5)
    s2, jit = 0.1, 1e-8
    per = (n + world - 1) // world
    lo, hi = min(n, rank * per), min(n, rank * per + per)       # same split as vfe_core
    Kzz = ref.kernelmatrix(ks, Z) + jit * np.eye(m)
    U = ref.cholesky_upper(Kzz)
    Xl, yl = X[lo:hi], y[lo:hi]
    A = sla.solve_triangular(U, (ref.kernelmatrix(ks, Xl, Z) / np.sqrt(s2)).T, trans="T", lower=False)  # M x n_local
    delta = yl / np.sqrt(s2)
    D = torch.from_numpy(A @ A.T)
    b = torch.from_numpy(A @ delta)
    sc = torch.tensor([len(yl) * np.log(s2), float(delta @ delta), float(ks.variance * len(yl) / s2), float((A * A).sum())])
    for t in (D, b, sc):
        dist.all_reduce(t)                                       # the one exchange step
    Lam = ref.cholesky_upper(D.numpy() + np.eye(m))
    sq = float((sla.solve_triangular(Lam, b.numpy(), trans="T", lower=False) ** 2).sum())
    dtc = -0.5 * (n * np.log(2 * np.pi) + sc[0].item() + ref.logdet_chol(Lam) + sc[1].item() - sq)
    elbo = dtc - 0.5 * (sc[2].item() - sc[3].item())
    want = ref.elbo(ks, ref.MeanSpec(), ref.NoiseSpec(0, s2), X, y, Z, ref.NoiseSpec(0, jit))
    out.put((rank, bool(abs(elbo - want) <= 1e-9 * abs(want)), float(elbo), float(want)))
    dist.destroy_process_group()


def test_vfe_data_sharding_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_vfe_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
