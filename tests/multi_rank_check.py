"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_rank_check.py [lj|eam] [cells] [steps]

Every rank owns one brick sub-domain (decomp.proc_grid / sub_box, the reference's rule) and
drives its own b200 context; the halo, borders and migration run over NCCL.  Rank 0 runs the
single-box oracle on the same atoms and checks:
  * the union of the per-rank half lists == the oracle's pair multiset (bit-exact keys)
  * forces by tag <= 1e-12 (relative to max|f|), energy and virial <= 1e-12
  * after `steps` timesteps (rebuilds, migration): atoms conserved, positions by tag and the
    global thermo tallies agree with the oracle to 1e-9, same number of list builds.
Prints MULTI_RANK_CHECK OK on success (tests/test_gpu_multi.py greps for it).
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from common import eam_system, lj_system, make_oracle, melted  # noqa: E402
from lammps_b200 import decomp  # noqa: E402
from lammps_b200.engine import Engine  # noqa: E402


def pair_keys(pi, pj, tag, x, xown_by_tag, prd):
    """canonical (tag_a, tag_b, sx, sy, sz) keys; image shifts relative to the owned copies"""
    tag = np.asarray(tag, np.int64)
    shift = np.rint((x - xown_by_tag[tag]) / prd).astype(np.int64)
    s = shift[pj] - shift[pi]
    ta, tb = tag[pi], tag[pj]
    swap = ta > tb
    a, b = np.where(swap, tb, ta), np.where(swap, ta, tb)
    s = np.where(swap[:, None], -s, s)
    same = ta == tb
    if same.any():
        sgn = np.sign(s[:, 0] * 9 + s[:, 1] * 3 + s[:, 2])
        s = np.where((same & (sgn < 0))[:, None], -s, s)
    return np.stack([a, b, s[:, 0], s[:, 1], s[:, 2]], axis=1)


def sort_keys(k):
    return k[np.lexsort(k.T[::-1])]


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "lj"
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    s = (lj_system if kind == "lj" else eam_system)((cells, cells, cells))
    s = melted(s, 40)  # O(1) forces, atoms already off-lattice
    n = len(s["x"])
    prd = np.asarray(s["hi"]) - np.asarray(s["lo"])
    grid = decomp.proc_grid(world, tuple(prd))
    myloc = decomp.rank_to_loc(rank, grid)
    # ownership by the periodically wrapped position (the melted state carries atoms that
    # drifted out of the box since the last rebuild; Verlet::setup wraps them before exchange)
    lo = np.asarray(s["lo"], float)
    xw = s["x"] - np.floor((s["x"] - lo) / prd) * prd
    xw = np.where(xw >= np.asarray(s["hi"], float), lo, xw)
    sel = decomp.owned_mask(xw, s["lo"], s["hi"], grid, myloc)

    e = Engine(local, "double", s["units"])
    decomp.init_comm(e, dist, rank, world)
    e.set_box(s["lo"], s["hi"])
    e.set_decomposition(grid, myloc)
    e.set_atoms(s["x"][sel], s["v"][sel], s["type"][sel], s["tag"][sel], s["mass"],
                image=s["image"][sel] if "image" in s else None, natoms_total=n)
    e.neighbor(s["skin"], every=s["every"], delay=s["delay"], check=s["check"])
    e.fix_nve(s["dt"])
    (e.pair_lj_cut if kind == "lj" else e.pair_eam)(s["tables"])
    e.setup(1, 1)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # ---- static parity
    a = e.get_atoms(ghosts=True, fields=("x", "tag"))
    nl, ng = e.counts()
    nn, pi, pj = e.neighbor_list()
    own = gather((a["tag"][:nl], a["x"][:nl]))
    xown = np.zeros((n + 1, 3))
    for t, x in own:
        xown[t] = x
    keys = gather(pair_keys(pi, pj, a["tag"], a["x"], xown, prd))
    fa = e.get_atoms(fields=("f", "tag"))
    fall = gather((fa["tag"], fa["f"]))
    eng, vir = e.tallies()
    ok = True
    msg = []
    if rank == 0:
        o = make_oracle(s)
        o.setup(1, 1)
        opi, opj = o.pairs()
        xo_own = np.zeros((n + 1, 3))
        xo_own[o.tag()] = o.x()
        ko = sort_keys(pair_keys(opi, opj, o.tag(True), o.x(True), xo_own, prd))
        ke = sort_keys(np.concatenate(keys))
        same = ke.shape == ko.shape and np.array_equal(ke, ko)
        msg.append(f"pairs: {len(ke)} vs oracle {len(ko)} identical={same}")
        ok &= bool(same)
        f = np.zeros((n + 1, 3))
        cnt = np.zeros(n + 1, int)
        for t, ff in fall:
            f[t] = ff
            cnt[t] += 1
        ok &= bool((cnt[1:] == 1).all())
        fo = np.zeros((n + 1, 3))
        fo[o.tag()] = o.f()
        ferr = np.abs(f - fo).max() / np.abs(fo).max()
        eerr = abs(eng - o.eng_vdwl) / abs(o.eng_vdwl)
        verr = np.abs(vir - o.virial).max() / np.abs(o.virial).max()
        msg.append(f"static: force err {ferr:.2e} energy err {eerr:.2e} virial err {verr:.2e}")
        ok &= ferr <= 1e-12 and eerr <= 1e-12 and verr <= 1e-12

    # ---- dynamics: rebuilds + migration
    th = e.run(steps, 0)
    b = e.get_atoms(fields=("x", "tag"))
    xall = gather((b["tag"], b["x"]))
    counts = gather(e.counts())
    builds = e.stats()["nbuilds"]
    if rank == 0:
        to = o.run(steps, 0, 0)
        x = np.full((n + 1, 3), np.nan)
        cnt = np.zeros(n + 1, int)
        for t, xx in xall:
            x[t] = xx
            cnt[t] += 1
        conserved = bool((cnt[1:] == 1).all()) and sum(c[0] for c in counts) == n
        xo = np.zeros((n + 1, 3))
        xo[o.tag()] = o.x()
        d = x[1:] - xo[1:]
        d -= np.rint(d / prd) * prd
        xerr = np.abs(d).max()
        terr = np.abs(th[-1][1:9] - to[-1][1:9]) / np.maximum(np.abs(to[-1][1:9]), 1e-300)
        msg.append(f"dynamics {steps} steps: atoms conserved={conserved} owned per rank="
                   f"{[c[0] for c in counts]} ghosts={[c[1] for c in counts]} max|dx|={xerr:.2e} "
                   f"tally rel err max={terr.max():.2e} builds {builds} vs oracle {o.ncalls}")
        ok &= conserved and xerr < 1e-9 and terr.max() < 1e-9 and builds == o.ncalls
        print("\n".join(msg))
        print(f"MULTI_RANK_CHECK {'OK' if ok else 'FAILED'} kind={kind} ranks={world} grid={grid} "
              f"natoms={n}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
