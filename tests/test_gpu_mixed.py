"""Mixed-precision mode (`package b200 prec mixed`): FP32 pair math, FP64 positions / cutoff
decisions / integration.  Tolerances are BASELINE.json's: forces <= 1e-5 relative (norm-wise),
energy and pressure <= 1e-6, neighbour pair sets still bit-exact; over a 100-step run the
thermo output must stay within the stated drift tolerance below."""
import numpy as np
import pytest

from common import by_tag, eam_system, lj_system, make_engine, make_oracle, melted
from test_gpu_parity import _pair_keys_engine, _pair_keys_oracle

pytestmark = pytest.mark.gpu

FTOL = 1e-5   # max|df| / max|f|
ETOL = 1e-6   # energy, virial trace (pressure)
# 100-step drift of thermo quantities against the FP64 oracle: relative, per quantity.  Measured
# (tools/mixed_drift_probe.py, FP32 pair math with FP64 accumulation): lj temp 4e-9, e_pair 9e-9,
# toteng 1e-8, press 4e-7; eam temp 1.5e-7, e_pair 5e-9, toteng 2e-10, press 3e-7 -- the 1e-7 level
# of the reference's own accelerator regression tolerance, with a margin of 5-7x.
DRIFT = {"temp": 1e-6, "e_pair": 1e-7, "toteng": 1e-7, "press": 2e-6}


def _static(s):
    o = make_oracle(s)
    o.setup(1, 1)
    e = make_engine(s, "mixed")
    e.setup(1, 1)
    assert e.counts() == (o.nlocal, o.nghost)
    ke, _ = _pair_keys_engine(e, s)
    ko = _pair_keys_oracle(o, s)
    assert np.array_equal(ke, ko), "mixed mode must not change the neighbour pair set"
    a = e.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(a["tag"], a["f"])
    (fo,) = by_tag(o.tag(), o.f())
    ferr = np.abs(fe - fo).max() / np.abs(fo).max()
    eng, vir = e.tallies()
    eerr = abs(eng - o.eng_vdwl) / abs(o.eng_vdwl)
    perr = abs(vir[:3].sum() - o.virial[:3].sum()) / abs(o.virial[:3].sum())
    return ferr, eerr, perr


def test_lj_mixed_forces_energy_pressure():
    ferr, eerr, perr = _static(melted(lj_system((12, 12, 12)), 60))
    assert 1e-9 < ferr <= FTOL, f"force error {ferr:.2e} (FP32 math must be visible and bounded)"
    assert eerr <= ETOL and perr <= ETOL, (eerr, perr)


def test_eam_mixed_forces_energy_pressure():
    ferr, eerr, perr = _static(melted(eam_system((8, 8, 8)), 40))
    assert 1e-9 < ferr <= FTOL, f"force error {ferr:.2e}"
    assert eerr <= ETOL and perr <= ETOL, (eerr, perr)


def test_lj_mixed_two_types_use_the_table_path():
    s = melted(lj_system((8, 8, 8)), 40)
    from lammps_b200 import pair_lj
    s["type"] = (1 + (np.arange(len(s["x"])) % 2)).astype(np.int32)
    s["mass"] = np.array([0.0, 1.0, 1.5])
    s["tables"] = pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.5),
                                            (1, 2): (0.9, 1.05, 2.5)}, 2.5)
    ferr, eerr, perr = _static(s)
    assert ferr <= FTOL and eerr <= ETOL and perr <= 5 * ETOL, (ferr, eerr, perr)


@pytest.mark.parametrize("kind", ["lj", "eam"])
def test_mixed_100_step_thermo_drift(kind):
    s = lj_system((12, 12, 12)) if kind == "lj" else eam_system((8, 8, 8))
    o = make_oracle(s)
    o.setup(1, 1)
    e = make_engine(s, "mixed")
    e.setup(1, 1)
    to = o.run(100, 0, 50)
    te = e.run(100, 50)
    assert len(to) == len(te) == 2
    assert e.stats()["nbuilds"] == o.ncalls
    for ro, re_ in zip(to, te):
        a, b = e.thermo_row(ro), e.thermo_row(re_)
        for k, tol in DRIFT.items():
            assert abs(a[k] - b[k]) <= tol * max(abs(a[k]), 1e-3), (kind, k, a[k], b[k])
