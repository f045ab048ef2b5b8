"""CPU tests: the oracle (oracle/md_oracle.c) against fixtures dumped from the compiled reference
(tests/golden/make_golden.py) and against the numbers printed in the reference's bench logs.
These pin the oracle; the GPU tests then compare the CUDA path with the oracle."""
import json

import numpy as np

from common import GOLDEN, by_tag, eam_system, lj_system, make_oracle
from oracle.oracle import canonical_pairs_box


def _static_vs_fixture(s, d, per_atom_pe):
    s.update(x=d["x"], v=d["v"], image=d["image"])
    o = make_oracle(s)
    o.setup(1, 1)
    assert o.nghost == int(d["nghost"])
    pi, pj = o.pairs()
    keys = canonical_pairs_box(pi, pj, o.tag(True), o.x(True), s["lo"], s["hi"], nlocal=o.nlocal)
    assert np.array_equal(keys, d["pair_keys"].astype(np.int64))       # bit-exact pair set
    (f,) = by_tag(o.tag(), o.f())
    # the fixture's atom order differs from ours (the reference sorts atoms spatially), so the
    # summation order differs: 1e-13, not bit-equal
    assert np.abs(f - d["f"]).max() <= 1e-13 * np.abs(d["f"]).max()
    pe = o.eng_vdwl / (len(s["x"]) if per_atom_pe else 1)
    assert abs(pe - float(d["pe"])) <= 1e-13 * abs(float(d["pe"]))


def test_lj_melt_fixture():
    _static_vs_fixture(lj_system((10, 10, 10)), np.load(GOLDEN / "ref_lj_melt_4k.npz"), True)


def test_eam_melt_fixture():
    _static_vs_fixture(eam_system((8, 8, 8)), np.load(GOLDEN / "ref_eam_melt_2k.npz"), False)


def _thermo(o, s, raw, normalize):
    from lammps_b200 import units
    u = units.get(s["units"])
    n = len(s["x"])
    dof = 3.0 * n - 3.0
    temp = raw[1] * u.mvv2e / (dof * u.boltz)
    vol = float(np.prod(s["hi"] - s["lo"]))
    press = (dof * u.boltz * temp + raw[3:6].sum()) / 3.0 / vol * u.nktv2p
    norm = n if normalize else 1
    return temp, raw[2] / norm, (raw[2] + 0.5 * dof * u.boltz * temp) / norm, press


def test_lj_bench_log_32k():
    """bench/in.lj: thermo at step 0 and 100, neighbour statistics, 5 builds."""
    g = json.loads((GOLDEN / "ref_lj_32k.json").read_text())
    s = lj_system((20, 20, 20))
    o = make_oracle(s)
    o.setup(1, 1)
    raw0 = np.array([0, o.ke_sum(), o.eng_vdwl, *o.virial, 0])
    rows = [raw0] + list(o.run(100, 0, 100))
    for raw, ref, pub in zip(rows, g["thermo"], g["published_log"]["thermo"]):
        got = _thermo(o, s, raw, True)
        want = (ref["temp"], ref["e_pair"], ref["toteng"], ref["press"])
        for a, b, p in zip(got, want, pub[1:]):
            assert abs(a - b) <= 1e-12 * max(abs(b), 1e-3)
            assert f"{a:.8g}" == f"{p:.8g}"
    assert o.ncalls == g["published_log"]["builds"]
    assert o.nneigh == g["published_log"]["neighbors"]
    assert o.nghost == g["published_log"]["nghost"]


def test_eam_bench_log_32k():
    g = json.loads((GOLDEN / "ref_eam_32k.json").read_text())
    s = eam_system((20, 20, 20))
    o = make_oracle(s)
    o.setup(1, 1)
    raw0 = np.array([0, o.ke_sum(), o.eng_vdwl, *o.virial, 0])
    rows = [raw0] + list(o.run(100, 0, 50))
    for raw, ref, pub in zip(rows, g["thermo"], g["published_log"]["thermo"]):
        got = _thermo(o, s, raw, False)
        want = (ref["temp"], ref["e_pair"], ref["toteng"], ref["press"])
        for a, b, p in zip(got, want, pub[1:]):
            assert abs(a - b) <= 1e-12 * max(abs(b), 1e-3)
            assert f"{a:.8g}" == f"{p:.8g}"
    assert o.ncalls == g["published_log"]["builds"]
    assert o.ndanger == g["published_log"]["dangerous"]
    assert o.nneigh == g["published_log"]["neighbors"]
    assert o.nghost == g["published_log"]["nghost"]


def test_neighbor_class_counts():
    """unittest/cplusplus/test_neighbor_class.cpp:235-269: 4x4x4-cell sc lattice spacing 2.0,
    lj/cut 3.5, skin 0.3 default... pinned here in the form the oracle supports: numneigh of
    every atom of a perfect sc lattice must be the same half-shell count."""
    n = 6
    g = np.arange(n) * 2.0
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + 0.5
    from lammps_b200 import pair_lj
    from oracle.oracle import Oracle
    o = Oracle()
    o.set_box([0, 0, 0], [2.0 * n] * 3)
    o.set_atoms(x, np.zeros_like(x), np.ones(len(x), np.int32), np.arange(1, len(x) + 1), [0, 1.0])
    o.set_neighbor(0.3, 1, 0, True)
    o.fix_nve(0.005)
    o.pair_lj_cut(pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 3.5)}, 3.5))
    o.setup(1, 1)
    # neighbours within 3.8: 6 (r=2) + 12 (2.83) + 8 (3.46) = 26 full -> 13 per atom on average
    assert o.nneigh == 13 * len(x)
    full = np.bincount(np.concatenate(o.pairs()), minlength=o.nlocal + o.nghost)
    tags = o.tag(True)
    per_tag = np.bincount(tags, weights=full)[1:]
    assert np.all(per_tag == 26)
