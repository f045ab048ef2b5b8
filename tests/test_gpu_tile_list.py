"""The two list layouts of the CUDA engine must describe the same physics: the bin-tile list
(kernels_tile.cuh, default for lj/cut) and the flat int32 half list (B200_LIST=flat) give the
same half-list pair set (bit-exact), the same forces and tallies, for lj/cut and eam, in both
precisions; odd tile sizes and a non-cubic box exercise the tile edge cases."""
import os

import numpy as np
import pytest

from common import by_tag, eam_system, lj_system, make_engine, make_oracle, melted
from oracle.oracle import canonical_pairs_box

pytestmark = pytest.mark.gpu


def _engine(s, mode, precision="double", tile=None):
    old = {k: os.environ.get(k) for k in ("B200_LIST", "B200_TILE")}
    os.environ["B200_LIST"] = mode
    if tile:
        os.environ["B200_TILE"] = tile
    else:
        os.environ.pop("B200_TILE", None)
    try:
        e = make_engine(s, precision)   # the environment is read by b200_create
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return e


def _keys(e, s):
    a = e.get_atoms(ghosts=True, fields=("x", "tag"))
    nn, pi, pj = e.neighbor_list()
    return canonical_pairs_box(pi, pj, a["tag"], a["x"], s["lo"], s["hi"], nlocal=e.counts()[0])


def _state(e):
    a = e.get_atoms(fields=("x", "f", "tag"))
    x, f = by_tag(a["tag"], a["x"], a["f"])
    eng, vir = e.tallies()
    return x, f, eng, np.asarray(vir)


CASES = [("lj", (9, 9, 9), None), ("lj", (10, 7, 6), "3,2,1"), ("lj", (8, 8, 8), "8,8,4"),
         ("eam", (7, 7, 7), None), ("eam", (8, 6, 7), "4,2,2")]


@pytest.mark.parametrize("kind,cells,tile", CASES)
def test_tile_list_equals_flat_list(kind, cells, tile):
    s = melted((lj_system if kind == "lj" else eam_system)(cells), 40)
    et, ef = _engine(s, "tile", tile=tile), _engine(s, "flat")
    for e in (et, ef):
        e.setup(1, 1)
    assert et.stats()["list_kind"] == 1 and ef.stats()["list_kind"] == 0
    assert et.stats()["npairs"] == ef.stats()["npairs"]
    assert np.array_equal(_keys(et, s), _keys(ef, s)), "half-list pair sets differ"
    xt, ft, engt, virt = _state(et)
    xf, ff, engf, virf = _state(ef)
    fmax = np.abs(ff).max()
    assert np.abs(ft - ff).max() <= 1e-12 * fmax
    assert abs(engt - engf) <= 1e-12 * abs(engf)
    assert np.abs(virt - virf).max() <= 1e-12 * np.abs(virf).max()
    # 60 steps incl. rebuilds: same trajectory
    et.run(60, 0)
    ef.run(60, 0)
    xt, ft, *_ = _state(et)
    xf, ff, *_ = _state(ef)
    assert np.abs(xt - xf).max() < 1e-9
    assert et.stats()["nbuilds"] == ef.stats()["nbuilds"]


@pytest.mark.parametrize("kind,cells", [("lj", (9, 9, 9)), ("eam", (7, 7, 7))])
def test_tile_list_mixed_precision(kind, cells):
    """mixed pair math on the tile list against the FP64 oracle: forces <= 1e-5, energy <= 1e-6"""
    s = melted((lj_system if kind == "lj" else eam_system)(cells), 40)
    o = make_oracle(s)
    o.setup(1, 1)
    e = _engine(s, "tile", "mixed")
    e.setup(1, 1)
    assert e.stats()["list_kind"] == 1
    a = e.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(a["tag"], a["f"])
    (fo,) = by_tag(o.tag(), o.f())
    assert np.abs(fe - fo).max() <= 1e-5 * np.abs(fo).max()
    eng, vir = e.tallies()
    assert abs(eng - o.eng_vdwl) <= 1e-6 * abs(o.eng_vdwl)


@pytest.mark.parametrize("env", [{"B200_MIXED_FX": "1"}, {"B200_MIXED_FX": "0"}, {"B200_LIST": "flat"}],
                         ids=["tile-fixedpoint", "tile-fp64-staged", "flat"])
def test_lj_mixed_variants_meet_the_mixed_tolerances(env):
    """every mixed lj/cut kernel (fixed-point staged tile, FP64-staged tile, flat list) against the
    FP64 oracle: same pair set, forces <= 1e-5 (norm-wise), energy <= 1e-6, and a 60-step run
    that stays on the oracle's trajectory to 1e-4 sigma"""
    s = melted(lj_system((11, 9, 10)), 40)
    o = make_oracle(s)
    o.setup(1, 1)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        e = make_engine(s, "mixed")
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    e.setup(1, 1)
    assert e.stats()["npairs"] == o.nneigh
    a = e.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(a["tag"], a["f"])
    (fo,) = by_tag(o.tag(), o.f())
    assert np.abs(fe - fo).max() <= 1e-5 * np.abs(fo).max()
    eng, _ = e.tallies()
    assert abs(eng - o.eng_vdwl) <= 1e-6 * abs(o.eng_vdwl)
    e.run(60, 0)
    o.run(60, 0, 0)
    a = e.get_atoms(fields=("x", "tag"))
    (xe,) = by_tag(a["tag"], a["x"])
    (xo,) = by_tag(o.tag(), o.x())
    prd = np.asarray(s["hi"]) - np.asarray(s["lo"])
    d = xe - xo
    d -= prd * np.rint(d / prd)
    assert np.abs(d).max() < 1e-4


def test_two_type_lj_on_tiles_matches_oracle_and_flat_list():
    """the per-type-pair table path of the tile kernels (cutsq, lj1..lj4, offset looked up by
    (itype, jtype)), FP64: forces <= 1e-12 against the oracle, same pair set as the flat list"""
    from lammps_b200 import pair_lj
    s = melted(lj_system((9, 8, 10)), 40)
    s["type"] = (1 + (np.arange(len(s["x"])) % 2)).astype(np.int32)
    s["mass"] = np.array([0.0, 1.0, 1.5])
    s["tables"] = pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2),
                                            (1, 2): (0.9, 1.05, 2.4)}, 2.5)
    o = make_oracle(s)
    o.setup(1, 1)
    et, ef = _engine(s, "tile"), _engine(s, "flat")
    for e in (et, ef):
        e.setup(1, 1)
    assert et.stats()["list_kind"] == 1
    assert et.stats()["npairs"] == o.nneigh == ef.stats()["npairs"]
    assert np.array_equal(_keys(et, s), _keys(ef, s))
    a = et.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(a["tag"], a["f"])
    (fo,) = by_tag(o.tag(), o.f())
    assert np.abs(fe - fo).max() <= 1e-12 * np.abs(fo).max()
    eng, vir = et.tallies()
    assert abs(eng - o.eng_vdwl) <= 1e-12 * abs(o.eng_vdwl)
    assert np.abs(vir - o.virial).max() <= 1e-12 * np.abs(o.virial).max()
