"""eam on the second-generation tile kernels (kernels_eam2.cuh; the default for a single-element
potential in FP64): FULLGHOST rows split into NEAR and FAR entries, density + embedding in one
kernel, the fp forward halo as the only per-step exchange besides the positions, forces stored
(or fix nve applied in the force kernel's epilogue inside b200_run).
Against the oracle (PairEAM::compute, pair_eam.cpp:124-327): half-list pair set bit-exact, rho and
fp <= 1e-12, forces <= 1e-12 (relative to max|f|), energy and virial <= 1e-12; trajectories with
`check yes` rebuilds follow the oracle with and without the fused integrator; the flat half-list
kernels (B200_EAM2=0) and the first-generation tile kernels (B200_LIST=tile B200_EAM2=0) stay
covered; sub-domains sharing one GPU exercise the fp halo between sub-domains."""
import numpy as np
import pytest

from common import by_tag, eam_system, make_engine, make_oracle, melted
from oracle.oracle import canonical_pairs_box

pytestmark = pytest.mark.gpu


def _static(s, monkeypatch, env, list_kind, eam2):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    o = make_oracle(s)
    o.setup(1, 1)
    e = make_engine(s)
    e.setup(1, 1)
    st = e.stats()
    assert st["list_kind"] == list_kind
    assert st["npairs"] == o.nneigh
    # FULLGHOST rows hold every partner: about twice the half list
    assert (st["list_entries"] > 1.9 * o.nneigh) == eam2
    a = e.get_atoms(ghosts=True, fields=("x", "tag"))
    nn, pi, pj = e.neighbor_list()
    ke = canonical_pairs_box(pi, pj, a["tag"], a["x"], s["lo"], s["hi"], nlocal=e.counts()[0])
    opi, opj = o.pairs()
    ko = canonical_pairs_box(opi, opj, o.tag(True), o.x(True), s["lo"], s["hi"], nlocal=o.nlocal)
    assert np.array_equal(ke, ko), "half-list pair set differs from the oracle's"
    b = e.get_atoms(fields=("f", "tag"))
    (fe,) = by_tag(b["tag"], b["f"])
    (fo,) = by_tag(o.tag(), o.f())
    assert np.abs(fe - fo).max() <= 1e-12 * np.abs(fo).max()
    eng, vir = e.tallies()
    assert abs(eng - o.eng_vdwl) <= 1e-12 * abs(o.eng_vdwl)
    assert np.abs(np.asarray(vir) - o.virial).max() <= 1e-12 * np.abs(o.virial).max()
    rho_e, fp_e = e.eam_rho_fp()
    rho_o, fp_o = o.rho_fp()
    re_, fe_ = by_tag(b["tag"], rho_e, fp_e)
    ro_, fo_ = by_tag(o.tag(), rho_o, fp_o)
    assert np.abs(re_ - ro_).max() <= 1e-12 * np.abs(ro_).max()
    assert np.abs(fe_ - fo_).max() <= 1e-12 * np.abs(fo_).max()
    return e, o


@pytest.mark.parametrize("cells,tile", [((8, 8, 8), None), ((9, 6, 7), "4,2,2"), ((7, 7, 7), "3,2,1")])
def test_eam2_static_parity(monkeypatch, cells, tile):
    if tile:
        monkeypatch.setenv("B200_TILE", tile)
    _static(melted(eam_system(cells), 60), monkeypatch, {}, 1, True)


def test_eam2_every_partner_far_or_near(monkeypatch):
    """the NEAR/FAR split only orders a row: margins that put (almost) every partner into one of
    the two halves give the same physics"""
    s = melted(eam_system((7, 7, 7)), 60)
    for margin in ("-10.0", "10.0"):
        _static(s, monkeypatch, {"B200_EAM2_MARGIN": margin}, 1, True)


def test_flat_and_first_generation_tile_kernels_still_match(monkeypatch):
    s = melted(eam_system((7, 7, 7)), 60)
    _static(s, monkeypatch, {"B200_EAM2": "0"}, 0, False)
    _static(s, monkeypatch, {"B200_EAM2": "0", "B200_LIST": "tile"}, 1, False)


def _trajectory(s, nsteps, thermo):
    o = make_oracle(s)
    o.setup(1, 1)
    e = make_engine(s)
    e.setup(1, 1)
    assert e.stats()["list_kind"] == 1
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


def test_eam2_trajectory_check_yes():
    _trajectory(eam_system((8, 8, 8)), 100, 50)


def test_eam2_fused_nve_follows_the_oracle(monkeypatch):
    monkeypatch.setenv("B200_FUSE_MIN", "0")
    e = _trajectory(eam_system((8, 8, 8)), 100, 50)
    # fused steps launch no integrate kernel and no force clear: halo + 2 pair kernels + fp halo
    assert e.stats()["launches"] < 100 * 5 + 25 * 40


def test_eam2_fused_equals_unfused(monkeypatch):
    s = eam_system((7, 7, 7))
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
    assert dx < 1e-11 and dv < 1e-9, (dx, dv)


@pytest.mark.parametrize("nsub,fuse", [(2, False), (8, True)])
def test_eam2_between_subdomains(monkeypatch, nsub, fuse):
    """sub-domains sharing one GPU: boundary pairs evaluated on both sides, ghost fp by the forward
    halo, no density or force reverse halo; migration over 60 steps"""
    from test_gpu_subdomains import _check
    if fuse:
        monkeypatch.setenv("B200_FUSE_MIN", "0")
    _check(melted(eam_system((10, 10, 10)), 40), nsub, 60)
