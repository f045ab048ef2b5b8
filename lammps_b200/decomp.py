"""Spatial domain decomposition: one brick sub-domain per GPU / process, chosen and numbered
like the reference does for MPI ranks.

  proc_grid     ProcMap::onelevel_grid -> factor() + best_factors()   procmap.cpp:48,725,836
  rank_to_loc   MPI_Cart rank order (last dimension fastest)          procmap.cpp:361-374
  sub_box       Domain::set_local_box, uniform xsplit = i/P           domain.cpp, comm.cpp
  owned_mask    ownership test sublo <= x < subhi                     comm_brick.cpp:655,702
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def proc_grid(nprocs: int, prd=(1.0, 1.0, 1.0)):
    xprd, yprd, zprd = prd
    area = (xprd * yprd, xprd * zprd, yprd * zprd)
    best, bestsurf = (1, 1, nprocs), 2.0 * sum(area)
    for i in range(1, nprocs + 1):
        if nprocs % i:
            continue
        nyz = nprocs // i
        for j in range(1, nyz + 1):
            if nyz % j:
                continue
            k = nyz // j
            surf = area[0] / i / j + area[1] / i / k + area[2] / j / k
            if surf < bestsurf:  # strict: the first minimum in (x,y) enumeration order wins
                best, bestsurf = (i, j, k), surf
    return best


def rank_to_loc(rank: int, grid):
    px, py, pz = grid
    return (rank // (py * pz), (rank // pz) % py, rank % pz)


def loc_to_rank(loc, grid):
    px, py, pz = grid
    return (loc[0] * py + loc[1]) * pz + loc[2]


def sub_box(lo, hi, grid, loc):
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    prd = hi - lo
    sublo, subhi = np.empty(3), np.empty(3)
    for d in range(3):
        P, me = grid[d], loc[d]
        sublo[d] = lo[d] + prd[d] * (me * 1.0 / P)
        subhi[d] = lo[d] + prd[d] * ((me + 1) * 1.0 / P) if me < P - 1 else hi[d]
    return sublo, subhi


def owned_mask(x, lo, hi, grid, loc):
    sublo, subhi = sub_box(lo, hi, grid, loc)
    return np.all((x >= sublo) & (x < subhi), axis=1)


def init_comm(engine, dist, rank: int, world: int):
    """Create the NCCL communicator of the engine: rank 0 makes the ncclUniqueId, the host
    (torch.distributed) broadcasts its 128 bytes, every rank joins."""
    import torch
    buf = (C.c_char * 128)()
    if rank == 0:
        engine._chk(engine.L.b200_comm_unique_id(buf))
    t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
    dev = t.cuda() if dist.get_backend() == "nccl" else t
    dist.broadcast(dev, src=0)
    raw = bytes(dev.cpu().numpy().tobytes())
    idbuf = (C.c_char * 128).from_buffer_copy(raw)
    engine._chk(engine.L.b200_comm_init(engine.h, C.c_int(world), C.c_int(rank), idbuf))
