"""fix nve fused into the lj/cut pair kernel (k_tile_lj2<..., NVE>): inside b200_run, on steps
without tallies, the pair kernel applies final_integrate(n) + initial_integrate(n+1) itself and
writes the new positions to the other position buffer.  The engine turns this on for systems of
>= 65536 atoms per sub-domain; B200_FUSE_MIN=0 forces it for the test sizes, B200_FUSE=0 is the
unfused path.  Trajectories must follow the oracle exactly as the unfused path does."""
import numpy as np
import pytest

from common import by_tag, lj_system, make_engine, make_oracle, melted

pytestmark = pytest.mark.gpu


@pytest.fixture
def fused(monkeypatch):
    monkeypatch.setenv("B200_FUSE_MIN", "0")


def _trajectory(s, nsteps, thermo):
    o = make_oracle(s)
    o.setup(1, 1)
    e = make_engine(s)
    e.setup(1, 1)
    to = o.run(nsteps, 0, thermo)
    te = e.run(nsteps, thermo)
    assert len(to) == len(te)
    a = e.get_atoms(fields=("x", "v", "f", "tag", "image"))
    xe, ve, fe, ie = by_tag(a["tag"], a["x"], a["v"], a["f"], a["image"])
    xo, vo, fo, io = by_tag(o.tag(), o.x(), o.v(), o.f(), o.image())
    assert np.array_equal(ie, io)
    assert np.abs(xe - xo).max() < 1e-9 and np.abs(ve - vo).max() < 1e-9
    assert np.abs(fe - fo).max() / np.abs(fo).max() < 1e-8
    for ro, re_ in zip(to, te):
        assert np.allclose(ro[1:9], re_[1:9], rtol=1e-9, atol=0)
    assert e.stats()["nbuilds"] == o.ncalls
    return e


def test_fused_nve_follows_the_oracle(fused):
    e = _trajectory(lj_system((12, 12, 12)), 100, 50)
    # the fused steps launch no integrate kernel: per plain step one halo + one pair kernel
    assert e.stats()["launches"] < 100 * 3 + 400


def test_fused_nve_with_displacement_checks(fused):
    s = lj_system((10, 10, 10))
    s.update(every=1, delay=0, check=True)   # the vote for decide() comes from the pair kernel's epilogue
    _trajectory(s, 80, 0)


def test_fused_nve_two_types_and_sub_group(fused):
    from lammps_b200 import pair_lj
    s = melted(lj_system((10, 10, 10)), 30)
    n = len(s["x"])
    s["type"] = (1 + (np.arange(n) % 2)).astype(np.int32)
    s["mass"] = np.array([0.0, 1.0, 1.7])
    s["tables"] = pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.5),
                                            (1, 2): (0.9, 1.05, 2.5)}, 2.5)
    _trajectory(s, 60, 20)


def test_fused_equals_unfused(monkeypatch):
    """same operations in the same order: the two paths agree to rounding-level noise"""
    s = lj_system((10, 10, 10))
    out = []
    for env in ({"B200_FUSE": "0"}, {"B200_FUSE_MIN": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = make_engine(s)
        e.setup(1, 1)
        e.run(60, 0)
        a = e.get_atoms(fields=("x", "v", "tag"))
        out.append(by_tag(a["tag"], a["x"], a["v"]))
        e.close()
        for k in env:
            monkeypatch.delenv(k)
    dx, dv = np.abs(out[0][0] - out[1][0]).max(), np.abs(out[0][1] - out[1][1]).max()
    assert dx < 1e-11 and dv < 1e-11, (dx, dv)


def test_fused_nve_between_subdomains(fused):
    """8 sub-domains on one GPU, peer-memory halo, migration: the fused pair kernels write the
    positions the next step's halo packs"""
    from test_gpu_subdomains import _check
    _check(melted(lj_system((12, 12, 12)), 40), 8, 100)


def test_fused_nve_mixed_precision_drift(fused):
    """the mixed kernel with the sub-domain-wide fixed-point records AND the fused integrator (its
    epilogue writes the next step's staged record): 100-step thermo drift within the mixed-mode
    tolerances of test_gpu_mixed.py"""
    from test_gpu_mixed import DRIFT
    s = lj_system((12, 12, 12))
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
            assert abs(a[k] - b[k]) <= tol * max(abs(a[k]), 1e-3), (k, a[k], b[k])
    a = e.get_atoms(fields=("x", "tag"))
    (xe,) = by_tag(a["tag"], a["x"])
    (xo,) = by_tag(o.tag(), o.x())
    assert np.abs(xe - xo).max() < 5e-3   # chaotic divergence of an FP32-force trajectory, bounded
