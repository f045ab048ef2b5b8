"""Several brick sub-domains driven by ONE process on ONE GPU (b200_group_*, `package b200
gpus N`): the part of the multi-GPU path that needs no second device -- ownership split,
CommBrick::exchange (migration), borders with remote ghosts, the peer-memory forward halo
between sub-domains -- checked against the single-box oracle exactly like the torchrun
multi-rank check does (tests/multi_rank_check.py):
  * the union of the sub-domains' half lists == the oracle's pair multiset (bit-exact keys)
  * forces by tag <= 1e-12 (relative to max|f|), energy and virial <= 1e-12
  * after 100 (60) timesteps with rebuilds and migration: atoms conserved, positions by tag
    and the global tallies agree with the oracle to 1e-9, same number of list builds."""
import numpy as np
import pytest

from common import eam_system, lj_system, make_oracle, melted
from multi_rank_check import pair_keys, sort_keys

pytestmark = pytest.mark.gpu


def _make_group(s, nsub, precision="double", grid=None):
    from lammps_b200.engine import EngineGroup
    g = EngineGroup([0] * nsub, precision, s["units"], grid=grid)
    g.set_box(s["lo"], s["hi"])
    lo, hi = np.asarray(s["lo"], float), np.asarray(s["hi"], float)
    prd = hi - lo
    # Verlet::setup wraps atoms into the box before it exchanges them (domain->pbc)
    xw = s["x"] - np.floor((s["x"] - lo) / prd) * prd
    xw = np.where(xw >= hi, lo, xw)
    img = s.get("image")
    if img is not None:  # keep unwrapped coordinates: fold the wrap into the image flags
        sh = np.rint((s["x"] - xw) / prd).astype(np.int64)
        ix = (img & 1023) + sh[:, 0]
        iy = ((img >> 10) & 1023) + sh[:, 1]
        iz = ((img >> 20) & 1023) + sh[:, 2]
        img = ((ix & 1023) | ((iy & 1023) << 10) | ((iz & 1023) << 20)).astype(np.int32)
    g.set_atoms(xw, s["v"], s["type"], s["tag"], s["mass"], image=img)
    g.neighbor(s["skin"], every=s["every"], delay=s["delay"], check=s["check"])
    g.fix_nve(s["dt"])
    (g.pair_lj_cut if s["kind"] == "lj" else g.pair_eam)(s["tables"])
    return g


def _check(s, nsub, steps, grid=None):
    n = len(s["x"])
    prd = np.asarray(s["hi"], float) - np.asarray(s["lo"], float)
    o = make_oracle(s)
    o.setup(1, 1)
    g = _make_group(s, nsub, grid=grid)
    g.setup(1, 1)
    assert g.counts()[0] == n
    assert int(np.prod(g.grid)) == nsub
    # ---- static parity: pair multiset over all sub-domains
    xown = np.zeros((n + 1, 3))
    subs = []
    for e in g.sub:
        a = e.get_atoms(ghosts=True, fields=("x", "tag"))
        nl, ng = e.counts()
        xown[a["tag"][:nl]] = a["x"][:nl]
        subs.append((e, a, nl, ng))
    keys = []
    for e, a, nl, ng in subs:
        nn, pi, pj = e.neighbor_list()
        keys.append(pair_keys(pi, pj, a["tag"], a["x"], xown, prd))
    opi, opj = o.pairs()
    xo_own = np.zeros((n + 1, 3))
    xo_own[o.tag()] = o.x()
    ko = sort_keys(pair_keys(opi, opj, o.tag(True), o.x(True), xo_own, prd))
    ke = sort_keys(np.concatenate(keys))
    assert ke.shape == ko.shape and np.array_equal(ke, ko), "union of the sub-domain lists != oracle"
    assert g.stats()["npairs"] == o.nneigh
    fa = g.get_atoms(fields=("f", "tag"))
    assert np.array_equal(np.sort(fa["tag"]), np.arange(1, n + 1))
    f = np.zeros((n + 1, 3))
    f[fa["tag"]] = fa["f"]
    fo = np.zeros((n + 1, 3))
    fo[o.tag()] = o.f()
    assert np.abs(f - fo).max() / np.abs(fo).max() <= 1e-12
    eng, vir = g.tallies()
    assert abs(eng - o.eng_vdwl) <= 1e-12 * abs(o.eng_vdwl)
    assert np.abs(vir - o.virial).max() <= 1e-12 * np.abs(o.virial).max()
    # ---- dynamics: rebuilds + migration between the sub-domains
    own0 = [e.counts()[0] for e in g.sub]
    th = g.run(steps, 0)
    to = o.run(steps, 0, 0)
    b = g.get_atoms(fields=("x", "tag"))
    assert np.array_equal(np.sort(b["tag"]), np.arange(1, n + 1)), "atoms lost or duplicated"
    x = np.zeros((n + 1, 3))
    x[b["tag"]] = b["x"]
    xo = np.zeros((n + 1, 3))
    xo[o.tag()] = o.x()
    d = x[1:] - xo[1:]
    d -= np.rint(d / prd) * prd
    assert np.abs(d).max() < 1e-9
    terr = np.abs(th[-1][1:9] - to[-1][1:9]) / np.maximum(np.abs(to[-1][1:9]), 1e-300)
    assert terr.max() < 1e-9, terr
    assert g.stats()["nbuilds"] == o.ncalls
    own1 = [e.counts()[0] for e in g.sub]
    g.close()
    return own0, own1


def test_eight_subdomains_on_one_gpu_lj():
    s = melted(lj_system((12, 12, 12)), 40)
    own0, own1 = _check(s, 8, 100)
    assert own0 != own1, "no atom migrated between the sub-domains: the test would prove nothing"


def test_two_subdomains_on_one_gpu_eam():
    s = melted(eam_system((8, 8, 8)), 40)
    _check(s, 2, 60)


def test_four_subdomains_slab_grid_lj():
    s = melted(lj_system((16, 8, 8)), 40)
    own0, own1 = _check(s, 4, 60, grid=(4, 1, 1))
    assert own0 != own1
