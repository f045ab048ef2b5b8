"""Round-2 list rules against committed fixtures generated from the reference
(tests/golden/ref_r2_*.npz, made by tests/golden/make_golden_round2.py): triclinic boxes (tag rule),
newton off, neigh_modify exclude group, lj/cut and eam.  The oracle is checked on the CPU, the CUDA
path (through the C ABI) on the GPU; neither needs oracle/_ref at run time."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

from common import eam_tables
from lammps_b200 import pair_lj, units

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["lj_tri_a", "lj_tri_b", "lj_newtoff", "lj_tri_newtoff", "lj_exclude_group", "eam_tri", "eam_newtoff"]


def pair_keys(pi, pj, tag, x):
    ta, tb = tag[pi].astype(np.int64), tag[pj].astype(np.int64)
    d = x[pj] - x[pi]
    swap = ta > tb
    a, b = np.where(swap, tb, ta), np.where(swap, ta, tb)
    d = np.where(swap[:, None], -d, d)
    q = np.rint(d * 1e6).astype(np.int64)
    same = ta == tb
    if same.any():
        sgn = np.sign(q[:, 2] * 4 + q[:, 1] * 2 + q[:, 0])
        q = np.where((same & (sgn < 0))[:, None], -q, q)
    key = np.stack([a, b, q[:, 0], q[:, 1], q[:, 2]], axis=1)
    return key[np.lexsort(key.T[::-1])]


def configure(o, d, is_engine):
    kind = str(d["kind"])
    u = units.get("lj" if kind == "lj" else "metal")
    if bool(d["triclinic"]):
        o.set_box_triclinic(d["lo"], d["hi"], *[float(t) for t in d["tilt"]])
    else:
        o.set_box(d["lo"], d["hi"])
    mass = np.array([0.0, 1.0]) if kind == "lj" else eam_tables().mass
    o.set_atoms(d["x"], d["v"], d["type"], d["tag"], mass, mask=d["mask"], image=d["image"])
    skin = 0.3 if kind == "lj" else 1.0
    if is_engine:
        o.neighbor(skin, every=int(d["every"]), delay=int(d["delay"]), check=bool(d["check"]))
        o.fix_nve(0.005)
    else:
        o.set_neighbor(skin, every=int(d["every"]), delay=int(d["delay"]), check=bool(d["check"]))
        o.fix_nve(0.005, u.ftm2v)
    if kind == "lj":
        o.pair_lj_cut(pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5))
    else:
        o.pair_eam(eam_tables().as_dict())
    if not int(d["newton"]):
        o.set_newton(False)
    if bool(d["exclude"]):
        o.neigh_modify_groups([(2, 4), (8, 8)])     # odd x even, slab x slab (group bits 1, 2, 3)
    return u, mass


def check(d, keys, nlocal, nghost, tag, f, eng, virial_xy, u, mass):
    assert (nlocal, nghost) == (len(d["x"]), int(d["nghost"]))
    assert len(keys) == int(d["npairs"])
    assert hashlib.sha256(np.ascontiguousarray(keys, dtype=np.int64).tobytes()).hexdigest() == str(d["pair_hash"])
    if len(d["pair_keys"]):
        assert np.array_equal(keys, d["pair_keys"].astype(np.int64))
    o, r = np.argsort(tag), np.argsort(d["tag"])
    assert np.abs(f[o] - d["f"][r]).max() <= 1e-12 * np.abs(d["f"]).max()
    assert abs(eng - float(d["pe"])) <= 1e-12 * abs(float(d["pe"]))
    mv = (mass[d["type"]] * d["v"][:, 0] * d["v"][:, 1]).sum() * u.mvv2e
    pxy = (mv + virial_xy) / float(d["vol"]) * u.nktv2p
    assert abs(pxy - float(d["pxy"])) <= 1e-9 * max(abs(float(d["press"])), abs(float(d["pxy"])))


def check_after(d, tag, x, f, eng):
    o = np.argsort(tag)
    assert np.array_equal(tag[o], d["tag40"])
    bad = np.abs(x[o] - d["x40"]).max(axis=1) > 1e-9   # an atom within rounding of a face may wrap either way
    assert bad.sum() <= 2
    assert np.abs(f[o] - d["f40"]).max() <= 1e-8 * np.abs(d["f40"]).max()
    assert abs(eng - float(d["pe40"])) <= 1e-10 * abs(float(d["pe40"]))


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_fixture(case):
    from oracle.oracle import Oracle
    d = np.load(GOLDEN / f"ref_r2_{case}.npz")
    o = Oracle()
    u, mass = configure(o, d, False)
    o.setup(1, 1)
    keys = pair_keys(*o.pairs(), o.tag(True), o.x(True))
    check(d, keys, o.nlocal, o.nghost, o.tag(), o.f(), o.eng_vdwl, o.virial[3], u, mass)
    o.run(40, 0, 40)
    check_after(d, o.tag(), o.x(), o.f(), o.eng_vdwl)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_engine_reproduces_reference_fixture(case):
    from lammps_b200.engine import Engine
    d = np.load(GOLDEN / f"ref_r2_{case}.npz")
    e = Engine(0, "double", "lj" if str(d["kind"]) == "lj" else "metal")
    u, mass = configure(e, d, True)
    e.setup(1, 1)
    a = e.get_atoms(ghosts=True, fields=("x", "tag"))
    _, pi, pj = e.neighbor_list()
    keys = pair_keys(pi, pj, a["tag"], a["x"])
    g = e.get_atoms(fields=("f", "tag"))
    eng, vir = e.tallies()
    check(d, keys, *e.counts(), g["tag"], g["f"], eng, vir[3], u, mass)
    e.run(40, 40)
    g = e.get_atoms(fields=("x", "f", "tag"))
    check_after(d, g["tag"], g["x"], g["f"], e.tallies()[0])
    e.close()
