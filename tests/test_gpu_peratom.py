"""Per-atom energy and virial of the pair style on the device (b200_pair_peratom: Pair::ev_tally's
eatom / vatom, pair.cpp:1087-1182, what compute pe/atom and stress/atom read) against the oracle's
orc_pair_peratom -- itself pinned to the compiled reference's compute pe/atom and compute
stress/atom in tests/test_oracle_vs_ref_live.py.  Every list flavour: FULLGHOST tile rows (lj/cut
default, eam on small systems: no scatter), flat half lists (RED onto ghosts + reverse halo), and
sub-domains sharing a GPU (ghost shares cross sub-domain boundaries).  Tolerance 1e-12 relative to
the largest per-atom value; the per-atom values must also sum to the global tallies."""
import numpy as np
import pytest

from common import by_tag, eam_system, lj_system, make_engine, make_oracle, melted

pytestmark = pytest.mark.gpu


def _check(s, monkeypatch, env, list_kind):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    o = make_oracle(s)
    o.setup(1, 1)
    eo, vo = o.pair_peratom()
    eo, vo = by_tag(o.tag(), eo, vo)
    e = make_engine(s)
    e.setup(1, 1)
    assert e.stats()["list_kind"] == list_kind
    ee, ve = e.pair_peratom()
    a = e.get_atoms(fields=("tag",))
    ee, ve = by_tag(a["tag"], ee, ve)
    assert np.abs(ee - eo).max() <= 1e-12 * np.abs(eo).max()
    assert np.abs(ve - vo).max() <= 1e-12 * np.abs(vo).max()
    eng, vir = e.tallies()
    assert abs(ee.sum() - eng) <= 1e-11 * abs(eng)
    assert np.abs(ve.sum(axis=0) - np.asarray(vir)).max() <= 1e-10 * np.abs(vir).max()
    return e, o


@pytest.mark.parametrize("env,kind", [({}, 1), ({"B200_LIST": "flat"}, 0)], ids=["tile", "flat"])
def test_lj_per_atom_tallies(monkeypatch, env, kind):
    _check(melted(lj_system((9, 8, 10)), 40), monkeypatch, env, kind)


def test_two_type_lj_per_atom_tallies(monkeypatch):
    from lammps_b200 import pair_lj
    s = melted(lj_system((9, 9, 8)), 40)
    s["type"] = (1 + (np.arange(len(s["x"])) % 2)).astype(np.int32)
    s["mass"] = np.array([0.0, 1.0, 1.5])
    s["tables"] = pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2),
                                            (1, 2): (0.9, 1.05, 2.4)}, 2.5, offset_flag=True)
    _check(s, monkeypatch, {}, 1)


@pytest.mark.parametrize("env,kind", [({}, 1), ({"B200_EAM2": "0"}, 0)], ids=["tile-eam2", "flat"])
def test_eam_per_atom_tallies(monkeypatch, env, kind):
    _check(melted(eam_system((8, 7, 8)), 60), monkeypatch, env, kind)


def test_per_atom_tallies_after_a_run_and_only_after_a_tally_step(monkeypatch):
    from lammps_b200.engine import B200Error
    s = lj_system((10, 10, 10))
    o = make_oracle(s)
    o.setup(1, 1)
    e = make_engine(s)
    e.setup(1, 1)
    o.run(45, 0, 0)
    e.run(45, 0)                       # the last step of a run tallies
    eo, vo = by_tag(o.tag(), *o.pair_peratom())
    a = e.get_atoms(fields=("tag",))
    ee, ve = by_tag(a["tag"], *e.pair_peratom())
    assert np.abs(ee - eo).max() <= 1e-9 * np.abs(eo).max()
    assert np.abs(ve - vo).max() <= 1e-9 * np.abs(vo).max()
    import ctypes as C
    e._chk(e.L.b200_step(e.h, C.c_int(0), C.c_int(0), None))   # a plain step leaves final_integrate pending
    with pytest.raises(B200Error):
        e.pair_peratom()


@pytest.mark.parametrize("kind,env", [("lj", {}), ("eam", {"B200_EAM2": "0"})], ids=["lj-tile", "eam-flat"])
def test_per_atom_tallies_between_subdomains(monkeypatch, kind, env):
    """8 (lj) / 2 (eam, half list: ghost shares return through the reverse halo) sub-domains on one GPU"""
    from test_gpu_subdomains import _make_group
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    s = melted((lj_system if kind == "lj" else eam_system)((10, 10, 10)), 40)
    o = make_oracle(s)
    o.setup(1, 1)
    eo, vo = by_tag(o.tag(), *o.pair_peratom())
    g = _make_group(s, 8 if kind == "lj" else 2)
    g.setup(1, 1)
    ee, ve = g.pair_peratom()
    a = g.get_atoms(fields=("tag",))
    ee, ve = by_tag(a["tag"], ee, ve)
    assert np.abs(ee - eo).max() <= 1e-12 * np.abs(eo).max()
    assert np.abs(ve - vo).max() <= 1e-12 * np.abs(vo).max()
    g.close()
