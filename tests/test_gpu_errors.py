"""Error paths of the C ABI (include/b200_md.h status codes): each must surface as a negative
status with a message, never as a hang, a silent wrong answer or a CPU fallback."""
import numpy as np
import pytest

from common import lj_system, make_engine

pytestmark = pytest.mark.gpu


def _raises(fn, code, text):
    from lammps_b200.engine import B200Error
    with pytest.raises(B200Error) as ei:
        fn()
    assert f"b200 error {code}:" in str(ei.value) and text in str(ei.value), str(ei.value)


def test_capacity_neigh_modify_one():
    """B200_ECAPACITY (-3): npair_bin.cpp:248 `Neighbor list overflow, boost neigh_modify one`"""
    s = lj_system((8, 8, 8))
    e = make_engine(s)
    e.neighbor(s["skin"], every=20, delay=0, check=False, one=20)
    _raises(lambda: e.setup(1, 1), -3, "Neighbor list overflow, boost neigh_modify one")


def test_nonfinite_coordinates():
    """B200_ENONFINITE (-4): nbin.cpp:145 / domain.cpp:787 `Non-numeric atom coords`"""
    s = lj_system((8, 8, 8))
    s["x"] = s["x"].copy()
    s["x"][17, 1] = np.nan
    e = make_engine(s)
    _raises(lambda: e.setup(1, 1), -4, "Non-numeric atom coords")


def test_lost_atom_on_a_non_periodic_box():
    """B200_ELOST (-5): an atom that leaves a non-periodic box is outside the bin grid"""
    from lammps_b200.engine import Engine
    s = lj_system((8, 8, 8))
    e = Engine(0, "double", s["units"])
    lo, hi = np.asarray(s["lo"], float) - 2.0, np.asarray(s["hi"], float) + 2.0
    e.set_box(lo, hi, periodic=(0, 0, 0))
    v = s["v"].copy()
    v[5] = (400.0, 0.0, 0.0)   # 2 sigma per step: through the ghost shell within a few steps
    e.set_atoms(s["x"], v, s["type"], s["tag"], s["mass"])
    e.neighbor(s["skin"], every=1, delay=0, check=True)
    e.fix_nve(s["dt"])
    e.pair_lj_cut(s["tables"])
    e.setup(0, 0)
    _raises(lambda: e.run(20, 0), -5, "lost atom")


def test_call_order_and_arguments():
    """B200_EARG (-2): step before setup, pair style sized for another number of types"""
    from lammps_b200 import pair_lj
    s = lj_system((6, 6, 6))
    e = make_engine(s)
    _raises(lambda: e.run(1, 0), -2, "before b200_setup")
    e.pair_lj_cut(pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (1.0, 1.0, 2.5),
                                           (1, 2): (1.0, 1.0, 2.5)}, 2.5))
    _raises(lambda: e.setup(1, 1), -2, "pair style was set for 2 types but atoms have 1")


def test_group_rejects_a_grid_that_does_not_match():
    from lammps_b200.engine import B200Error, EngineGroup
    s = lj_system((8, 8, 8))
    g = EngineGroup([0, 0], "double", s["units"], grid=(2, 2, 1))
    with pytest.raises(B200Error) as ei:
        g.set_box(s["lo"], s["hi"])
    assert "does not match the number of sub-domains" in str(ei.value)
    g.close()
