"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches
tests/multi_rank_check.py under torchrun, one rank per GPU, NCCL halo between sub-domains."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(nproc, kind, cells, steps, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "multi_rank_check.py"), kind, str(cells), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_RANK_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("kind,cells,steps", [("lj", 12, 100), ("eam", 8, 60)])
def test_two_ranks_match_oracle(kind, cells, steps):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, kind, cells, steps, 29531)


def test_eight_ranks_match_oracle():
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8, "lj", 16, 100, 29532)


def test_peer_memory_halo_gives_up_instead_of_hanging():
    """the spin-wait of the peer-memory halo (k_p2p_unpack_forward) has a limit (P2P_SPIN_LIMIT,
    ~2 s): a sub-domain whose neighbour never sends reports B200_ECUDA "timed out" at the next
    sync instead of hanging the GPU.  Two sub-domains on two devices in one process; only one of
    them is stepped."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import ctypes as C
    sys.path.insert(0, str(ROOT / "tests"))
    from common import lj_system
    from lammps_b200.engine import B200Error, EngineGroup
    s = lj_system((10, 10, 10))
    g = EngineGroup([0, 1], "double", s["units"])
    g.set_box(s["lo"], s["hi"])
    g.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"])
    g.neighbor(s["skin"], every=20, delay=0, check=False)
    g.fix_nve(s["dt"])
    g.pair_lj_cut(s["tables"])
    g.setup(1, 1)
    e0 = g.sub[0]
    e0.sync()                                                      # (selects device 0 in this thread)
    e0._chk(e0.L.b200_step(e0.h, C.c_int(0), C.c_int(0), None))   # the neighbour stays silent
    with pytest.raises(B200Error, match="timed out"):
        e0.sync()


def test_lmp_b200_package_drives_two_gpus_from_one_process():
    """`lmp_b200 -sf b200 -pk b200 gpus 2 -in bench/in.lj` (unmodified input): one LAMMPS process,
    one brick sub-domain per GPU, peer access instead of MPI.  The thermo output must be the golden
    log's (tests/golden/ref_lj_32k.json) and the list statistics the reference's."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import re
    sys.path.insert(0, str(ROOT / "tests"))
    from test_gpu_lammps_pkg import BENCH, GOLDEN, close_to_printed, run_lmp, thermo_rows
    out = run_lmp(["-sf", "b200", "-pk", "b200", "gpus", "2", "-in", "in.lj"], cwd=BENCH)
    assert "2 sub-domains on 2 GPU(s)" in out
    g = json.loads((GOLDEN / "ref_lj_32k.json").read_text())["published_log"]
    rows = thermo_rows(out)
    assert [int(r[0]) for r in rows] == [int(r[0]) for r in g["thermo"]]
    for got, ref in zip(rows, g["thermo"]):
        for a, b in zip(got[1:], ref[1:]):
            assert close_to_printed(a, b), (got, ref)
    m = re.search(r"Total # of neighbors = (\d+)", out)
    assert m and int(m.group(1)) == g["neighbors"]
    m = re.search(r"Neighbor list builds = (\d+)", out)
    assert m and int(m.group(1)) == g["builds"]


@pytest.mark.parametrize("variant", ["triclinic-langevin-exclude", "newton-off-eam-triclinic"])
def test_round2_features_on_two_gpus_match_reference_executable(tmp_path, variant):
    """the round-2 additions across two real devices (`package b200 gpus 2`: peer stores over
    NVLink instead of copies inside one GPU): a prism box cut into two lamda bricks, group masks
    of remote ghosts, fix langevin with zero yes (the reference's random stream drawn on the host,
    summed random force over both sub-domains), newton off, eam -- thermo every step against lmp_ref"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import re
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    from test_gpu_triclinic import EAM_TRI, LJ_TRI, REF, EXE, _run, _script
    if variant == "triclinic-langevin-exclude":
        body = "atom_modify sort 0 0\n" + _script(LJ_TRI, 10, "2.0 -1.0 3.0", "every 2 delay 0 check yes").replace(
            "fix 1 all nve", "group odd id 1:4000:2\ngroup even id 2:4000:2\nneigh_modify exclude group odd even\n"
                             "fix 1 all nve\nfix 2 all langevin 1.2 0.9 0.5 33 zero yes")
        extra = ["-pk", "b200", "gpus", "2", "langevin_rng", "host"]
    else:
        body = "newton off\n" + _script(EAM_TRI, 8, "1.5 -2.0 1.0", "every 1 delay 5 check yes")
        extra = ["-pk", "b200", "gpus", "2"]
    body += """
thermo 1
thermo_style custom step temp pe etotal press pxy pxz pyz
thermo_modify format float %.12g
dump 1 all custom 60 f.dump id x y z vx fx fy fz
dump_modify 1 sort id format float %.10g
run 60
"""
    ta, da, oa = _run(REF, [], tmp_path / "ref", body, 8)
    tb, db, ob = _run(EXE, ["-sf", "b200", *extra], tmp_path / "b200", body, 8)
    assert "2 sub-domains on 2 GPU(s)" in ob
    assert ta.shape == tb.shape == (61, 8)
    scale = np.maximum(np.abs(ta).max(axis=0), 1e-3)
    assert (np.abs(ta - tb).max(axis=0) <= 1e-9 * scale).all(), np.abs(ta - tb).max(axis=0) / scale
    assert np.abs(da[:, 5:8] - db[:, 5:8]).max() <= 1e-8 * np.abs(da[:, 5:8]).max()
    m = re.search(r"Neighbor list builds = (\d+)", oa)
    assert m and m.group(0) in ob
