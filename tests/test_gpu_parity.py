"""GPU parity tests: the CUDA engine (through the C ABI) against the oracle on the same seeded
inputs and against the committed reference fixtures.  Tolerances are BASELINE.json's:
FP64 mode forces <= 1e-12 relative (norm-wise: max|df| / max|f|), energy/virial <= 1e-12;
neighbour pair sets bit-exact (sorted multiset of canonical keys)."""
import json

import numpy as np
import pytest

from common import GOLDEN, by_tag, eam_system, lj_system, make_engine, make_oracle, melted
from oracle.oracle import canonical_pairs_box

pytestmark = pytest.mark.gpu

FTOL = 1e-12
ETOL = 1e-12


def _pair_keys_engine(e, s):
    a = e.get_atoms(ghosts=True, fields=("x", "tag"))
    nn, pi, pj = e.neighbor_list()
    nl, ng = e.counts()
    return canonical_pairs_box(pi, pj, a["tag"], a["x"], s["lo"], s["hi"], nlocal=nl), nn


def _pair_keys_oracle(o, s):
    pi, pj = o.pairs()
    return canonical_pairs_box(pi, pj, o.tag(True), o.x(True), s["lo"], s["hi"], nlocal=o.nlocal)


def _check_static(s, label, fscale=0.0):
    o = make_oracle(s)
    o.setup(1, 1)
    e = make_engine(s)
    e.setup(1, 1)
    nl, ng = e.counts()
    assert nl == o.nlocal
    assert ng == o.nghost, f"{label}: ghost count {ng} vs oracle {o.nghost}"
    ke, nn = _pair_keys_engine(e, s)
    ko = _pair_keys_oracle(o, s)
    assert ke.shape == ko.shape, f"{label}: {len(ke)} pairs vs oracle {len(ko)}"
    assert np.array_equal(ke, ko), f"{label}: neighbour pair sets differ"
    a = e.get_atoms(fields=("f", "tag", "x"))
    (fe,) = by_tag(a["tag"], a["f"])
    (fo,) = by_tag(o.tag(), o.f())
    fmax = max(np.abs(fo).max(), fscale)
    err = np.abs(fe - fo).max() / fmax
    assert err <= FTOL, f"{label}: force error {err:.3e} (max|f| {fmax:.3e})"
    eng, vir = e.tallies()
    assert abs(eng - o.eng_vdwl) <= ETOL * abs(o.eng_vdwl)
    vo = o.virial
    assert np.abs(vir - vo).max() <= ETOL * np.abs(vo).max()
    st = e.stats()
    assert st["npairs"] == o.nneigh
    return e, o


def test_lj_lattice_step0():
    # perfect lattice: net forces are rounding noise (~1e-13), so the error is measured against
    # the magnitude of a single pair force (~1 in LJ units) instead of max|f|
    _check_static(lj_system((8, 8, 8)), "lj lattice", fscale=1.0)


def test_lj_melted_forces_pairs_energy():
    s = melted(lj_system((10, 10, 10)), 80)
    _check_static(s, "lj melt")


def test_lj_noncubic_box():
    s = melted(lj_system((12, 7, 9)), 40)
    _check_static(s, "lj 12x7x9")


def test_eam_melted_forces_pairs_energy():
    s = melted(eam_system((8, 8, 8)), 60)
    e, o = _check_static(s, "eam melt")
    rho_e, fp_e = e.eam_rho_fp()
    a = e.get_atoms(fields=("tag",))
    rho_o, fp_o = o.rho_fp()
    re_, fe_ = by_tag(a["tag"], rho_e, fp_e)
    ro_, fo_ = by_tag(o.tag(), rho_o, fp_o)
    assert np.abs(re_ - ro_).max() <= 1e-12 * np.abs(ro_).max()
    assert np.abs(fe_ - fo_).max() <= 1e-12 * np.abs(fo_).max()


def test_lj_reference_fixture_4k():
    """Against the state dumped from the compiled reference (tests/golden/make_golden.py)."""
    d = np.load(GOLDEN / "ref_lj_melt_4k.npz")
    s = lj_system((10, 10, 10))
    s.update(x=d["x"], v=d["v"], image=d["image"])
    e = make_engine(s)
    e.setup(1, 1)
    assert e.counts()[1] == int(d["nghost"])
    ke, _ = _pair_keys_engine(e, s)
    assert np.array_equal(ke, d["pair_keys"].astype(np.int64))
    a = e.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(a["tag"], a["f"])
    assert np.abs(fe - d["f"]).max() / np.abs(d["f"]).max() <= FTOL
    eng, vir = e.tallies()
    assert abs(eng / len(s["x"]) - float(d["pe"])) <= ETOL * abs(float(d["pe"]))
    row = e.thermo_row([0, e.ke_sum(), eng, *vir, 0])
    assert abs(row["press"] - float(d["press"])) <= 1e-11 * abs(float(d["press"]))


def test_eam_reference_fixture_2k():
    d = np.load(GOLDEN / "ref_eam_melt_2k.npz")
    s = eam_system((8, 8, 8))
    s.update(x=d["x"], v=d["v"], image=d["image"])
    e = make_engine(s)
    e.setup(1, 1)
    assert e.counts()[1] == int(d["nghost"])
    ke, _ = _pair_keys_engine(e, s)
    assert np.array_equal(ke, d["pair_keys"].astype(np.int64))
    a = e.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(a["tag"], a["f"])
    assert np.abs(fe - d["f"]).max() / np.abs(d["f"]).max() <= FTOL
    eng, vir = e.tallies()
    assert abs(eng - float(d["pe"])) <= ETOL * abs(float(d["pe"]))


def _run_compare(s, nsteps, thermo_every):
    o = make_oracle(s)
    o.setup(1, 1)
    to = o.run(nsteps, 0, thermo_every)
    e = make_engine(s)
    e.setup(1, 1)
    te = e.run(nsteps, thermo_every)
    assert len(te) == len(to)
    st = e.stats()
    assert st["nbuilds"] == o.ncalls, f"builds {st['nbuilds']} vs oracle {o.ncalls}"
    assert st["ndanger"] == o.ndanger
    return e, o, te, to


def test_lj_100_steps_thermo_and_trajectory():
    """100-step run: rebuild schedule identical, thermo within 1e-9 relative of the oracle
    (chaotic divergence from summation-order differences stays far below that in 100 steps)."""
    s = lj_system((10, 10, 10))
    e, o, te, to = _run_compare(s, 100, 50)
    for a, b in zip(te, to):
        assert a[0] == b[0]
        assert abs(a[1] - b[1]) <= 1e-9 * abs(b[1])
        assert abs(a[2] - b[2]) <= 1e-9 * abs(b[2])
        assert np.abs(a[3:6] - b[3:6]).max() <= 1e-8 * np.abs(b[3:6]).max()
    a = e.get_atoms(fields=("x", "v", "tag", "image"))
    xe, ve, ie = by_tag(a["tag"], a["x"], a["v"], a["image"])
    xo, vo, io = by_tag(o.tag(), o.x(), o.v(), o.image())
    assert np.abs(xe - xo).max() < 1e-8
    assert np.abs(ve - vo).max() < 1e-8
    assert np.array_equal(ie, io)


def test_eam_100_steps_check_yes():
    s = eam_system((8, 8, 8))
    e, o, te, to = _run_compare(s, 100, 50)
    for a, b in zip(te, to):
        assert abs(a[1] - b[1]) <= 1e-9 * abs(b[1])
        assert abs(a[2] - b[2]) <= 1e-9 * abs(b[2])


def test_lj_bench_32k_golden_log():
    """bench/in.lj as-is: thermo must reproduce the reference's published log to all printed
    digits (FP64 mode) and its neighbour statistics exactly."""
    g = json.loads((GOLDEN / "ref_lj_32k.json").read_text())
    s = lj_system((20, 20, 20))
    e = make_engine(s)
    e.setup(1, 1)
    eng, vir = e.tallies()
    r0 = e.thermo_row([0, e.ke_sum(), eng, *vir, 0])
    rows = [r0] + [e.thermo_row(r) for r in e.run(100, 100)]
    for row, ref, pub in zip(rows, g["thermo"], g["published_log"]["thermo"]):
        assert row["step"] == ref["step"]
        for k, col in (("temp", 1), ("e_pair", 2), ("toteng", 3), ("press", 4)):
            assert abs(row[k] - ref[k]) <= 2e-9 * max(abs(ref[k]), 1e-3), (k, row[k], ref[k])
            assert f"{row[k]:.8g}" == f"{pub[col]:.8g}", (k, row[k], pub[col])
    st = e.stats()
    assert st["nbuilds"] == g["published_log"]["builds"]
    assert st["npairs"] == g["published_log"]["neighbors"]
    assert e.counts()[1] == g["published_log"]["nghost"]


def test_eam_bench_32k_golden_log():
    g = json.loads((GOLDEN / "ref_eam_32k.json").read_text())
    s = eam_system((20, 20, 20))
    e = make_engine(s)
    e.setup(1, 1)
    eng, vir = e.tallies()
    r0 = e.thermo_row([0, e.ke_sum(), eng, *vir, 0])
    rows = [r0] + [e.thermo_row(r) for r in e.run(100, 50)]
    for row, ref, pub in zip(rows, g["thermo"], g["published_log"]["thermo"]):
        assert row["step"] == ref["step"]
        for k, col in (("temp", 1), ("e_pair", 2), ("toteng", 3), ("press", 4)):
            assert abs(row[k] - ref[k]) <= 2e-9 * max(abs(ref[k]), 1e-3), (k, row[k], ref[k])
            assert f"{row[k]:.8g}" == f"{pub[col]:.8g}", (k, row[k], pub[col])
    st = e.stats()
    assert st["nbuilds"] == g["published_log"]["builds"]
    assert st["ndanger"] == g["published_log"]["dangerous"]
    assert st["npairs"] == g["published_log"]["neighbors"]
    assert e.counts()[1] == g["published_log"]["nghost"]


def test_step_granular_equals_run():
    """verlet/b200 drives the step-granular entry points; they must reproduce b200_run."""
    s = lj_system((8, 8, 8))
    s["every"], s["check"] = 2, True
    e1 = make_engine(s)
    e1.setup(0, 0)
    e1.run(30, 0)
    e2 = make_engine(s)
    e2.setup(0, 0)
    for _ in range(30):
        e2.initial_integrate()
        if e2.decide():
            e2.reneighbor()
        else:
            e2.forward_comm()
        e2.force_clear()
        e2.pair_compute(0, 0)
        e2.reverse_comm()
        e2.final_integrate()
    a1 = e1.get_atoms(fields=("x", "tag"))
    a2 = e2.get_atoms(fields=("x", "tag"))
    (x1,) = by_tag(a1["tag"], a1["x"])
    (x2,) = by_tag(a2["tag"], a2["x"])
    assert np.abs(x1 - x2).max() < 1e-10
    assert e1.stats()["nbuilds"] == e2.stats()["nbuilds"]


@pytest.mark.parametrize("fixture", ["ref_yaml_pair_eam.npz", "ref_yaml_pair_eam_alloy.npz",
                                     "ref_yaml_pair_eam_fs.npz"])
def test_two_element_eam_tables_from_reference_fixtures(fixture):
    """eam (two funcfl files), eam/alloy (setfl) and eam/fs tables of the reference's own
    known-answer tests, on the CUDA path: the 32-atom cell of data.metal replicated 3x3x3 (the
    engine keeps one layer of periodic images) against the oracle, which reproduces the yaml
    answers themselves on the original cell (tests/test_oracle_reference_yaml.py)."""
    from test_oracle_reference_yaml import _alloy_system, _system
    s, _ = _system() if fixture == "ref_yaml_pair_eam.npz" else _alloy_system(fixture)
    prd = np.asarray(s["hi"]) - np.asarray(s["lo"])
    x0 = s["lo"] + np.mod(s["x"] - s["lo"], prd)
    reps = [(i, j, k) for i in range(3) for j in range(3) for k in range(3)]
    s = dict(s)
    s["x"] = np.concatenate([x0 + prd * np.array(r) for r in reps])
    s["v"] = np.tile(s["v"], (27, 1))
    s["type"] = np.tile(s["type"], 27)
    s["tag"] = np.arange(1, len(s["x"]) + 1, dtype=np.int32)
    s["hi"] = s["lo"] + 3 * prd
    s.pop("image", None)
    e, o = _check_static(s, fixture)
    to = o.run(20, 0, 0)
    te = e.run(20, 0)
    a = e.get_atoms(fields=("x", "tag"))
    (xe,) = by_tag(a["tag"], a["x"])
    (xo,) = by_tag(o.tag(), o.x())
    assert np.abs(xe - xo).max() < 1e-10
    assert abs(te[-1][2] - to[-1][2]) <= 1e-10 * abs(to[-1][2])
