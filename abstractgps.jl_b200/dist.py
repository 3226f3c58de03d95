"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL inside libagp.so.

torch.distributed (gloo) is used ONLY to ship rank 0's 128-byte ncclUniqueId to the other ranks and for
host-side barriers / max-reductions of timings; every collective on the data path (panel broadcast,
all-reduce of logdet / sqmahal, alpha broadcast) is issued by libagp.so on its own CUDA streams.

Tile ownership (mirrors agp_bc_owner / fit_dist_impl in csrc/engine.cu): block column j of the padded
matrix lives on rank j mod Q of a 1 x Q grid, as local block j div Q."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi as cabi
from .api import AGPError, Engine


def owner_of_block(j: int, world: int) -> int:
    return j % world


def local_blocks(nt: int, rank: int, world: int):
    return [j for j in range(nt) if j % world == rank]


def gather_row_for_local_col(n_local: int, first_global_block: int, kk: int, world: int, tile: int = 128) -> int:
    """row of the packed panel (origin = global row (kk+1)*tile) that local trailing column n_local reads:
    the b_tile_stride / b_off mapping of the trailing GEMM."""
    return (n_local // tile) * world * tile + n_local % tile + (first_global_block - (kk + 1)) * tile


def init_distributed_engine(backend: str = "gloo") -> Engine:
    """Create the per-rank engine for a torchrun launch (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return Engine(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    L = cabi.lib()
    idbuf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        rc = L.agp_nccl_unique_id(cabi.ptr(idbuf))
        if rc != cabi.AGP_OK:
            raise AGPError(rc, "agp_nccl_unique_id failed")
    t = torch.from_numpy(idbuf)
    dist.broadcast(t, src=0)
    eng = Engine.__new__(Engine)
    h = C.c_void_p()
    rc = L.agp_init_dist(C.byref(h), local, rank, world, 1, world, cabi.ptr(idbuf), None)
    if rc != cabi.AGP_OK:
        raise AGPError(rc, "agp_init_dist failed (rank %d of %d, device %d)" % (rank, world, local))
    eng.L, eng.h, eng.device = L, h, local
    eng.rank, eng.world = rank, world
    return eng
