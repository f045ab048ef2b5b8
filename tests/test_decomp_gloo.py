"""Host-side logic of the multi-GPU path on CPU: the brick decomposition and the neighbour /
message schedule of the 26-direction halo (b200_neighbor_ranks, a pure host function of the
C ABI), exercised by two gloo processes that exchange their send schedules."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from lammps_b200 import decomp  # noqa: E402
from lammps_b200.engine import load_library  # noqa: E402


def neighbor_ranks(grid, loc, periodic=(1, 1, 1)):
    L = load_library()
    g = (C.c_int * 3)(*grid)
    m = (C.c_int * 3)(*loc)
    p = (C.c_int * 3)(*periodic)
    out = (C.c_int * 27)()
    assert L.b200_neighbor_ranks(g, m, p, out) == 0
    return list(out)


def schedule(grid, rank, periodic=(1, 1, 1)):
    """(sends, recvs): per peer, the ordered list of directions this rank sends / expects,
    walking directions in ascending order like halo_exchange() in engine.cu does."""
    nbr = neighbor_ranks(grid, decomp.rank_to_loc(rank, grid), periodic)
    sends, recvs = {}, {}
    for d in range(27):
        if d == 13:
            continue
        to, frm = nbr[d], nbr[26 - d]
        if to >= 0 and to != rank:
            sends.setdefault(to, []).append(d)
        if frm >= 0 and frm != rank:
            recvs.setdefault(frm, []).append(d)
    return sends, recvs


def test_neighbor_ranks_match_python_decomp():
    for grid in [(1, 1, 2), (1, 2, 2), (2, 2, 2), (3, 2, 1), (4, 1, 2)]:
        n = grid[0] * grid[1] * grid[2]
        for r in range(n):
            loc = decomp.rank_to_loc(r, grid)
            assert decomp.loc_to_rank(loc, grid) == r
            nbr = neighbor_ranks(grid, loc)
            for d in range(27):
                dv = (d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1)
                l2 = tuple((loc[k] + dv[k]) % grid[k] for k in range(3))
                assert nbr[d] == decomp.loc_to_rank(l2, grid)
            open_nbr = neighbor_ranks(grid, loc, (0, 0, 0))
            for d in range(27):
                dv = (d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1)
                inside = all(0 <= loc[k] + dv[k] < grid[k] for k in range(3))
                assert (open_nbr[d] >= 0) == inside


def test_message_order_matches_between_every_pair_of_ranks():
    """NCCL matches several messages between one pair of ranks in issue order: what A sends to
    B (ascending direction) must be what B posts as receives from A, in the same order."""
    for grid in [(1, 1, 2), (1, 2, 2), (2, 2, 2), (3, 3, 3), (4, 2, 1)]:
        n = grid[0] * grid[1] * grid[2]
        sched = [schedule(grid, r) for r in range(n)]
        for a in range(n):
            for b, dirs in sched[a][0].items():
                assert sched[b][1].get(a) == dirs, (grid, a, b)


def test_owned_masks_partition_the_box():
    rng = np.random.default_rng(1)
    lo, hi = np.array([0.0, -1.0, 2.0]), np.array([10.0, 7.0, 11.0])
    x = lo + rng.random((5000, 3)) * (hi - lo)
    x[0] = lo
    for nprocs in (2, 4, 8, 6):
        grid = decomp.proc_grid(nprocs, tuple(hi - lo))
        assert grid[0] * grid[1] * grid[2] == nprocs
        owner_count = np.zeros(len(x), int)
        for r in range(nprocs):
            owner_count += decomp.owned_mask(x, lo, hi, grid, decomp.rank_to_loc(r, grid))
        assert (owner_count == 1).all()


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np, torch, torch.distributed as dist
from test_decomp_gloo import schedule
from lammps_b200 import decomp
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2,
                        init_method="tcp://127.0.0.1:" + os.environ["PORT"])
rank = dist.get_rank()
grid = decomp.proc_grid(2, (1.0, 1.0, 1.0))
sends, recvs = schedule(grid, rank)
peer = 1 - rank
# halo with fake payloads: message for direction d carries d repeated (d+1) times
out = [torch.full((d + 1,), float(d + 100 * rank)) for d in sends[peer]]
inn = [torch.zeros(d + 1) for d in recvs[peer]]
ops = [dist.P2POp(dist.isend, t, peer) for t in out] + [dist.P2POp(dist.irecv, t, peer) for t in inn]
for w in dist.batch_isend_irecv(ops):
    w.wait()
for d, t in zip(recvs[peer], inn):
    assert t.shape[0] == d + 1 and (t == float(d + 100 * peer)).all(), (rank, d, t)
# rebuild vote / thermo sums: MAX and SUM all-reduce as decide() / fetch_ev() use them
v = torch.tensor([float(rank)]); dist.all_reduce(v, op=dist.ReduceOp.MAX); assert v.item() == 1.0
e = torch.tensor([1.5 + rank]); dist.all_reduce(e); assert e.item() == 4.0
dist.barrier(); dist.destroy_process_group()
print("GLOO_HALO_OK", rank)
"""


def test_two_process_gloo_halo_schedule(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT), tests=str(ROOT / "tests")))
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True,
                              env={**os.environ, "RANK": str(r), "PORT": port}) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"GLOO_HALO_OK {r}" in o, o[-2000:]
