"""neigh_modify options on the device build: exclude type i j (NPair::exclusion, npair.cpp:244-248 --
excluded type pairs never enter the list; here they make the build's cutoff test fail, see
engine.cu setup_geometry) and once yes (neighbor.cpp:2420: no rebuild after setup).  Against the
oracle, whose versions are pinned to the compiled reference in tests/test_oracle_vs_ref_live.py."""
import numpy as np
import pytest

from common import by_tag, lj_system, make_engine, make_oracle, melted
from oracle.oracle import canonical_pairs_box

pytestmark = pytest.mark.gpu


def _two_types():
    from lammps_b200 import pair_lj
    s = melted(lj_system((9, 8, 9)), 40)
    s["type"] = (1 + (np.arange(len(s["x"])) % 3 == 0)).astype(np.int32)
    s["mass"] = np.array([0.0, 1.0, 1.5])
    s["tables"] = pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2),
                                            (1, 2): (0.9, 1.05, 2.4)}, 2.5)
    s.update(every=5, delay=0, check=False)
    return s


@pytest.mark.parametrize("mode", ["tile", "flat"])
@pytest.mark.parametrize("excl", [[(1, 2)], [(2, 2)], [(1, 1), (1, 2)]], ids=["1-2", "2-2", "1-1+1-2"])
def test_exclude_type_pairs(monkeypatch, mode, excl):
    monkeypatch.setenv("B200_LIST", mode)
    s = _two_types()
    o = make_oracle(s)
    o.neigh_modify(exclude_types=excl, ntypes=2)
    o.setup(1, 1)
    e = make_engine(s)
    e.neigh_modify(exclude_types=excl, ntypes=2)
    e.setup(1, 1)
    assert e.stats()["npairs"] == o.nneigh
    a = e.get_atoms(ghosts=True, fields=("x", "tag", "type"))
    nn, pi, pj = e.neighbor_list()
    ex = {tuple(sorted(p)) for p in excl}
    assert not any(tuple(sorted(t)) in ex for t in set(zip(a["type"][pi].tolist(), a["type"][pj].tolist())))
    ke = canonical_pairs_box(pi, pj, a["tag"], a["x"], s["lo"], s["hi"], nlocal=e.counts()[0])
    opi, opj = o.pairs()
    ko = canonical_pairs_box(opi, opj, o.tag(True), o.x(True), s["lo"], s["hi"], nlocal=o.nlocal)
    assert np.array_equal(ke, ko)
    b = e.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(b["tag"], b["f"])
    (fo,) = by_tag(o.tag(), o.f())
    assert np.abs(fe - fo).max() <= 1e-12 * np.abs(fo).max()
    eng, vir = e.tallies()
    assert abs(eng - o.eng_vdwl) <= 1e-12 * abs(o.eng_vdwl)
    # rebuilds keep the exclusion
    to, te = o.run(20, 0, 0), e.run(20, 0)
    assert e.stats()["nbuilds"] == o.ncalls == 4
    assert e.stats()["npairs"] == o.nneigh
    assert np.allclose(to[-1][1:9], te[-1][1:9], rtol=1e-9, atol=0)


def test_once_yes_never_rebuilds():
    s = lj_system((10, 10, 10))
    s.update(every=2, delay=0, check=True)
    o = make_oracle(s)
    o.neigh_modify(once=True)
    o.setup(1, 1)
    e = make_engine(s)
    e.neigh_modify(once=True)
    e.setup(1, 1)
    to, te = o.run(30, 0, 10), e.run(30, 10)
    assert e.stats()["nbuilds"] == o.ncalls == 0
    assert np.allclose(np.asarray(to)[:, 1:9], np.asarray(te)[:, 1:9], rtol=1e-9, atol=0)
    a = e.get_atoms(fields=("x", "tag"))
    (xe,) = by_tag(a["tag"], a["x"])
    (xo,) = by_tag(o.tag(), o.x())
    assert np.abs(xe - xo).max() < 1e-10
