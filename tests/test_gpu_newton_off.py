"""`newton off` (Force::newton_pair = 0; list rule NPairBin<HALF,!NEWTON>, npair_bin.cpp:126-131:
owned pairs once, owned-ghost pairs on both owners; no force returns from ghosts).  The product
runs it on the bin-tile rows that hold every ghost partner.  Checker: the compiled reference
(oracle/_ref) with `newton off` -- md_oracle.c restates the newton-on rule only."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from test_gpu_triclinic import EAM_TRI, EXE, LJ_TRI, POT, REF, _pair_keys, _reference_state, _run, _script

pytestmark = pytest.mark.gpu

LJ_ORTHO = LJ_TRI.replace("region box prism 0 CELLS 0 CELLS 0 CELLS TILT", "region box block 0 CELLS 0 CELLS 0 CELLS")
EAM_ORTHO = EAM_TRI.replace("region box prism 0 CELLS 0 CELLS 0 CELLS TILT", "region box block 0 CELLS 0 CELLS 0 CELLS")


def _off(script):
    return "newton off\n" + script


@pytest.mark.parametrize("base,tilt,subdomains", [(LJ_ORTHO, "", 1), (LJ_TRI, "2.0 -1.0 3.0", 1), (LJ_ORTHO, "", 8)],
                         ids=["orthogonal", "triclinic", "8-subdomains"])
def test_newton_off_list_forces_energy_equal_the_reference(base, tilt, subdomains):
    from lammps_b200 import pair_lj
    from lammps_b200.engine import Engine, EngineGroup
    st = _reference_state(_off(_script(base, 8, tilt, "every 1 delay 0 check yes")), 50)
    nl = st["nlocal"]
    e = Engine(0, "double", "lj") if subdomains == 1 else EngineGroup([0] * subdomains, "double", "lj")
    if tilt:
        e.set_box_triclinic(st["lo"], st["hi"], st["xy"], st["xz"], st["yz"])
    else:
        e.set_box(st["lo"], st["hi"])
    e.set_atoms(st["x"][:nl], st["v"], st["type"], st["tag"][:nl], np.array([0.0, 1.0]), image=st["image"])
    e.neighbor(0.3, every=1, delay=0, check=True)
    e.fix_nve(0.005)
    e.pair_lj_cut(pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5))
    subs = [e] if subdomains == 1 else e.sub
    for sub in subs:
        sub.set_newton(False)
    e.setup(1, 1)
    want = _pair_keys(*st["pairs"], st["tag"], st["x"])
    keys = []
    for sub in subs:
        a = sub.get_atoms(ghosts=True, fields=("x", "tag"))
        nn, pi, pj = sub.neighbor_list()
        assert nn.sum() == len(pi) == sub.stats()["npairs"]
        keys.append(_pair_keys(pi, pj, a["tag"], a["x"]))
    got = np.concatenate(keys)
    got = got[np.lexsort(got.T[::-1])]
    if subdomains == 1:
        # one box: every boundary pair appears twice (once from each of its two owned atoms)
        assert got.shape == want.shape and np.array_equal(got, want)
        assert len(np.unique(got, axis=0)) < len(got)
    else:
        # pairs that cross an inner sub-domain face are owned-ghost here and owned-owned in the
        # reference's single box: stored twice instead of once; as a SET the lists agree
        assert np.array_equal(np.unique(got, axis=0), np.unique(want, axis=0))
    a = e.get_atoms(fields=("f", "tag"))
    o, ro = np.argsort(a["tag"]), np.argsort(st["tag"][:nl])
    ferr = np.abs(a["f"][o] - st["f"][ro]).max() / np.abs(st["f"]).max()
    assert ferr <= 1e-12, ferr
    eng, _ = e.tallies()
    assert abs(eng / st["natoms"] - st["pe"]) <= 1e-12 * abs(st["pe"])
    e.close()


@pytest.mark.parametrize("kind,base,cells,tilt,neigh,extra", [
    ("lj", LJ_ORTHO, 10, "", "every 20 delay 0 check no", []),
    ("lj", LJ_TRI, 10, "2.0 -1.0 3.0", "every 1 delay 0 check yes", []),
    ("lj", LJ_ORTHO, 10, "", "every 2 delay 0 check yes", ["-pk", "b200", "subdomains", "8"]),
    ("eam", EAM_ORTHO, 8, "", "every 1 delay 5 check yes", []),
], ids=["lj", "lj-triclinic", "lj-8-subdomains", "eam"])
def test_newton_off_run_matches_reference_executable(tmp_path, kind, base, cells, tilt, neigh, extra):
    body = _off(_script(base, cells, tilt, neigh)) + """
compute pea all pe/atom
thermo 1
thermo_style custom step temp pe etotal press pxy pxz pyz
thermo_modify format float %.12g
dump 1 all custom 60 f.dump id x y z vx fx fy fz c_pea
dump_modify 1 sort id format float %.10g
run 60
"""
    ta, da, oa = _run(REF, [], tmp_path / "ref", body, 8)
    tb, db, ob = _run(EXE, ["-sf", "b200", *extra], tmp_path / "b200", body, 8)
    assert ta.shape == tb.shape == (61, 8)
    scale = np.maximum(np.abs(ta).max(axis=0), 1e-3)
    assert (np.abs(ta - tb).max(axis=0) <= 1e-9 * scale).all(), np.abs(ta - tb).max(axis=0) / scale
    assert np.array_equal(da[:, 0], db[:, 0])
    assert np.abs(da[:, 1:4] - db[:, 1:4]).max() <= 1e-8
    assert np.abs(da[:, 5:8] - db[:, 5:8]).max() <= 1e-8 * np.abs(da[:, 5:8]).max()
    assert np.abs(da[:, 8] - db[:, 8]).max() <= 1e-8 * np.abs(da[:, 8]).max()
    m = re.search(r"Neighbor list builds = (\d+)", oa)
    assert m and m.group(0) in ob
    if not extra:   # (with sub-domains inner-face pairs are owned-ghost: counted twice)
        m = re.search(r"Total # of neighbors = (\d+)", oa)
        assert m and m.group(0) in ob, (m.group(0), re.search(r"Total # of neighbors = (\d+)", ob).group(0))


def test_newton_off_without_full_ghost_rows_is_refused(tmp_path):
    body = _off(_script(LJ_ORTHO, 6, "", "every 1 delay 0 check yes")) + "run 5\n"
    d = tmp_path / "flat"
    d.mkdir()
    (d / "in.t").write_text(body)
    r = subprocess.run([str(EXE), "-sf", "b200", "-pk", "b200", "list", "flat", "-in", "in.t"], cwd=d,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "newton off needs the bin-tile list" in (r.stdout + r.stderr)
