"""CPU tests of the host-side setup code (lattice, velocities, coefficient tables, spline tables,
decomposition) against reference-derived fixtures."""
import numpy as np

from common import GOLDEN, eam_tables, lj_system
from lammps_b200 import decomp, lattice, pair_lj, units


def test_lattice_lj_matches_reference_box():
    x, lo, hi = lattice.fcc_block("lj", 0.8442, (20, 20, 20))
    assert len(x) == 32000
    assert abs(hi[0] - 33.591924) < 1e-6          # SURVEY appendix / bench log box edge
    assert np.all(x >= 0) and np.all(x < hi)
    d = np.load(GOLDEN / "ref_lj_melt_4k.npz")
    x4, lo4, hi4 = lattice.fcc_block("lj", 0.8442, (10, 10, 10))
    assert np.array_equal(hi4, d["hi"]) and np.array_equal(lo4, d["lo"])


def test_velocity_create_properties():
    s = lj_system((6, 6, 6))
    v, m = s["v"], 1.0
    n = len(v)
    assert np.abs(v.sum(axis=0)).max() < 1e-10            # momentum zeroed
    t = (v * v).sum() * m / (3 * n - 3)
    assert abs(t - 1.44) < 1e-12                          # scaled to the target temperature
    # loop geom: velocities depend on position only -> independent of how atoms are split
    half = lattice._geom_seeds(s["x"][: n // 2], 87287)
    full = lattice._geom_seeds(s["x"], 87287)
    assert np.array_equal(half, full[: n // 2])


def test_velocity_matches_oracle_rng():
    from oracle.oracle import velocity_loop_geom
    s = lj_system((4, 4, 4))
    raw = velocity_loop_geom(s["x"], 87287, np.ones(len(s["x"])))
    seeds = lattice._geom_seeds(s["x"], 87287)
    out = np.empty_like(raw)
    for c in range(3):
        seeds, u = lattice._park_uniform(seeds)
        out[:, c] = u - 0.5
    assert np.array_equal(out, raw)


def test_lj_tables_bench_values():
    t = pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5)
    assert t["lj1"][1, 1] == 48.0 and t["lj2"][1, 1] == 24.0
    assert t["lj3"][1, 1] == 4.0 and t["lj4"][1, 1] == 4.0
    assert t["cutsq"][1, 1] == 6.25 and t["offset"][1, 1] == 0.0


def test_lj_tables_mixing():
    t = pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0), (2, 2): (4.0, 2.0)}, 2.5, offset_flag=True)
    assert abs(t["lj4"][1, 2] - 4.0 * 2.0 * 2.0 ** 3) < 1e-12   # eps=2, sigma=sqrt(2)
    assert t["lj1"][1, 2] == t["lj1"][2, 1]
    assert t["offset"][1, 1] != 0.0


def test_eam_tables_shape_and_constants():
    T = eam_tables()
    assert (T.nr, T.nrho) == (499, 499)   # lround((n-1)*d/d), pair_eam.cpp:1026-1027
    assert abs(T.rdr - 100.0) < 1e-9 and abs(T.cutforcesq - 24.5025) < 1e-9
    assert T.rhor_spline.shape == (1, 500, 7) and T.frho_spline.shape == (2, 500, 7)
    assert T.mass[1] == 63.55
    # spline row = derivative quadratic | value cubic: d/dp of the cubic equals c5 at p=0
    s = T.z2r_spline[0]
    assert np.allclose(s[1:, 2] * T.dr, s[1:, 5])


def test_proc_grid_and_ownership_partition():
    assert decomp.proc_grid(2) == (1, 1, 2)
    assert decomp.proc_grid(4) == (1, 2, 2)
    assert decomp.proc_grid(8) == (2, 2, 2)
    s = lj_system((6, 6, 6))
    for n in (2, 4, 8):
        grid = decomp.proc_grid(n)
        owners = np.zeros(len(s["x"]), int)
        for r in range(n):
            loc = decomp.rank_to_loc(r, grid)
            assert decomp.loc_to_rank(loc, grid) == r
            owners += decomp.owned_mask(s["x"], s["lo"], s["hi"], grid, loc)
        assert np.all(owners == 1)          # every atom owned exactly once


def test_units():
    assert units.get("lj").ftm2v == 1.0
    assert abs(units.get("metal").ftm2v - 1.0 / 1.0364269e-4) < 1e-6
