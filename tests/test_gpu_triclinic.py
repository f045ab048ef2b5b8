"""Triclinic boxes: periodic wrap, migration and ghost slabs in lamda coordinates
(verlet.cpp:293-313, comm_brick.cpp:177-237, domain.cpp:2347-2390), bins over the bounding box
(nbin_standard.cpp:86-112), full stencil (nstencil_bin.cpp:36-62) and the tag rule of the
half/newton/tri list (npair_bin.cpp:133-155).

The checker is the compiled reference itself (oracle/_ref, built by oracle/build_ref.py from the
unmodified sources): md_oracle.c has no triclinic restatement.
* through the C ABI: the engine, given the reference's melted state, must store the reference's
  pair set exactly (unordered tag pair + separation vector), its forces to 1e-12, its energy and
  virial to 1e-12 -- one sub-domain and 2x2x2 sub-domains;
* through lmp_b200 -sf b200: thermo every step, rebuild counts and the final x, v, f against lmp_ref
  for lj/cut and eam, one and eight sub-domains, plus compute pe/atom."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "lammps_b200" / "lammps_pkg" / "lmp_b200"
REF = ROOT / "oracle" / "_ref" / "lmp_ref"
POT = ROOT / "oracle" / "_ref" / "potentials"

LJ_TRI = """
units lj
lattice fcc 0.8442
region box prism 0 CELLS 0 CELLS 0 CELLS TILT
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify NEIGH
fix 1 all nve
"""

EAM_TRI = """
units metal
lattice fcc 3.615
region box prism 0 CELLS 0 CELLS 0 CELLS TILT
create_box 1 box
create_atoms 1 box
mass 1 63.55
velocity all create 1600.0 376847 loop geom
pair_style eam
pair_coeff 1 1 POT/Cu_u3.eam
neighbor 1.0 bin
neigh_modify NEIGH
fix 1 all nve
"""


def _script(base, cells, tilt, neigh):
    return (base.replace("CELLS", str(cells)).replace("TILT", tilt).replace("NEIGH", neigh)
            .replace("POT", str(POT)))


def _pair_keys(pi, pj, tag, x):
    """unordered tag pair + separation vector (rounded to 1e-6): the identity of a stored pair
    whatever the atom order and whichever of its two mirror images an implementation holds"""
    ta, tb = tag[pi].astype(np.int64), tag[pj].astype(np.int64)
    d = x[pj] - x[pi]
    swap = ta > tb
    same = ta == tb
    a, b = np.where(swap, tb, ta), np.where(swap, ta, tb)
    d = np.where(swap[:, None], -d, d)
    q = np.rint(d * 1e6).astype(np.int64)
    if same.any():
        sgn = np.sign(q[:, 2] * 4 + q[:, 1] * 2 + q[:, 0])
        q = np.where((same & (sgn < 0))[:, None], -q, q)
    key = np.stack([a, b, q[:, 0], q[:, 1], q[:, 2]], axis=1)
    return key[np.lexsort(key.T[::-1])]


def _reference_state(script, nsteps):
    from oracle.ref_harness import RefLammps
    import ctypes as C
    L = RefLammps()
    L.commands(script + f"\nthermo 10\nrun {nsteps}\nrun 0\n")
    nl, ng = L.setting("nlocal"), L.setting("nghost")
    lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
    xy, yz, xz = C.c_double(), C.c_double(), C.c_double()
    L.lib.lammps_extract_box(L.h, lo, hi, C.byref(xy), C.byref(yz), C.byref(xz), None, None)
    st = dict(nlocal=nl, lo=np.array(lo), hi=np.array(hi), xy=xy.value, xz=xz.value, yz=yz.value,
              x=L.atom_vec3("x", nl + ng), v=L.atom_vec3("v", nl), f=L.atom_vec3("f", nl),
              tag=L.atom_int("id", nl + ng), type=L.atom_int("type", nl), image=L.atom_int("image", nl),
              pe=L.thermo("pe"), press=L.thermo("press"), natoms=L.natoms())
    st["pairs"] = L.neighbor_pairs("lj/cut")
    L.close()
    return st


@pytest.mark.parametrize("tilt,subdomains,mode", [
    ("2.0 -1.0 3.0", 1, "flat"), ("2.0 -1.0 3.0", 1, "tile"), ("-3.0 2.5 1.5", 1, "tile"),
    ("2.0 -1.0 3.0", 8, "flat"), ("2.0 -1.0 3.0", 8, "tile")],
    ids=["tilt-a-flat", "tilt-a-tile", "tilt-b-tile", "tilt-a-8-subdomains-flat", "tilt-a-8-subdomains-tile"])
def test_triclinic_list_forces_energy_equal_the_reference(tilt, subdomains, mode):
    from lammps_b200 import pair_lj
    from lammps_b200.engine import Engine, EngineGroup
    st = _reference_state(_script(LJ_TRI, 8, tilt, "every 1 delay 0 check yes"), 50)
    nl = st["nlocal"]
    assert abs(st["xy"]) > 0 and abs(st["xz"]) > 0 and abs(st["yz"]) > 0
    e = Engine(0, "double", "lj") if subdomains == 1 else EngineGroup([0] * subdomains, "double", "lj")
    e.set_box_triclinic(st["lo"], st["hi"], st["xy"], st["xz"], st["yz"])
    e.set_atoms(st["x"][:nl], st["v"], st["type"], st["tag"][:nl], np.array([0.0, 1.0]), image=st["image"])
    e.neighbor(0.3, every=1, delay=0, check=True)
    e.fix_nve(0.005)
    e.pair_lj_cut(pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5))
    for sub in ([e] if subdomains == 1 else e.sub):
        sub.set_option("list", mode)
    e.setup(1, 1)
    want = _pair_keys(*st["pairs"], st["tag"], st["x"])
    if subdomains == 1:
        assert e.stats()["list_kind"] == (1 if mode == "tile" else 0)
        a = e.get_atoms(ghosts=True, fields=("x", "tag"))
        _, pi, pj = e.neighbor_list()
        got = _pair_keys(pi, pj, a["tag"], a["x"])
        assert e.counts() == (nl, len(st["tag"]) - nl)
    else:
        keys = []
        for sub in e.sub:
            a = sub.get_atoms(ghosts=True, fields=("x", "tag"))
            _, pi, pj = sub.neighbor_list()
            keys.append(_pair_keys(pi, pj, a["tag"], a["x"]))
        got = np.concatenate(keys)
        got = got[np.lexsort(got.T[::-1])]
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)
    a = e.get_atoms(fields=("x", "f", "tag"))
    o = np.argsort(a["tag"])
    ro = np.argsort(st["tag"][:nl])
    assert np.abs(a["x"][o] - st["x"][:nl][ro]).max() < 1e-13
    ferr = np.abs(a["f"][o] - st["f"][ro]).max() / np.abs(st["f"]).max()
    assert ferr <= 1e-12, ferr
    eng, vir = e.tallies()
    assert abs(eng / st["natoms"] - st["pe"]) <= 1e-12 * abs(st["pe"])
    e.close()


def test_triclinic_fused_run_on_tiles_equals_flat_list_run():
    """100 steps through b200_run: bin tiles with the integrator inside the pair kernel against the
    flat half list stage by stage -- same rebuilds, positions to 1e-9"""
    from lammps_b200 import pair_lj
    from lammps_b200.engine import Engine
    st = _reference_state(_script(LJ_TRI, 8, "2.0 -1.0 3.0", "every 1 delay 0 check yes"), 20)
    nl = st["nlocal"]
    res = {}
    for mode in ("tile", "flat"):
        e = Engine(0, "double", "lj")
        e.set_box_triclinic(st["lo"], st["hi"], st["xy"], st["xz"], st["yz"])
        e.set_atoms(st["x"][:nl], st["v"], st["type"], st["tag"][:nl], np.array([0.0, 1.0]), image=st["image"])
        e.neighbor(0.3, every=1, delay=0, check=True)
        e.fix_nve(0.005)
        e.pair_lj_cut(pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5))
        e.set_option("list", mode)
        e.setup(1, 1)
        th = e.run(100, 50)
        a = e.get_atoms(fields=("x", "v", "tag"))
        o = np.argsort(a["tag"])
        res[mode] = (th, a["x"][o], a["v"][o], e.stats()["nbuilds"])
        e.close()
    assert res["tile"][3] == res["flat"][3] > 3
    assert np.abs(res["tile"][1] - res["flat"][1]).max() < 1e-9
    assert np.abs(res["tile"][2] - res["flat"][2]).max() < 1e-9
    assert np.abs(res["tile"][0] - res["flat"][0]).max() <= 1e-9 * np.abs(res["flat"][0]).max()


@pytest.mark.parametrize("variant", ["triclinic", "triclinic-newton-off", "newton-off-exclude-group", "eam-triclinic"])
def test_engine_equals_oracle_restatement(variant):
    """the CUDA path against md_oracle.c (itself pinned to the compiled reference in
    tests/test_oracle_tri_newton_live.py) on a melted state: pair sets exact, forces, energy and
    virial 1e-12, then 60 steps of both (positions 1e-9, same rebuild count)"""
    from common import eam_system, lj_system, make_engine, make_oracle
    eam = variant.startswith("eam")
    s = eam_system((6, 6, 6)) if eam else lj_system((8, 8, 8))
    n = len(s["x"])
    a = (s["hi"][0] - s["lo"][0]) / (6 if eam else 8)
    # tilt factors of whole lattice constants keep the crystal periodic in the sheared cell
    tilt = (2 * a, -a, 3 * a) if "triclinic" in variant else None
    mask = (1 | np.where(np.arange(n) % 2 == 0, 2, 4)).astype(np.int32)
    # melt it in that cell first (forces of a perfect lattice are ~1e-13: nothing to compare)
    o0 = make_oracle(s)
    if tilt:
        o0.set_box_triclinic(s["lo"], s["hi"], *tilt)
        o0.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"])
    o0.setup(0, 0)
    o0.run(40)
    order = np.argsort(o0.tag())
    s = dict(s, x=o0.x()[order], v=o0.v()[order], image=o0.image()[order])
    objs = []
    for make in (make_oracle, make_engine):
        o = make(s)
        if tilt:
            o.set_box_triclinic(s["lo"], s["hi"], *tilt)
        o.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"], mask=mask, image=s.get("image"))
        if "newton-off" in variant:
            o.set_newton(False)
        if "exclude-group" in variant:
            o.neigh_modify_groups([(2, 4)])
        o.setup(1, 1)
        objs.append(o)
    o, e = objs
    ae = e.get_atoms(ghosts=True, fields=("x", "tag"))
    _, pi, pj = e.neighbor_list()
    kor = _pair_keys(*o.pairs(), o.tag(True), o.x(True))
    keng = _pair_keys(pi, pj, ae["tag"], ae["x"])
    assert kor.shape == keng.shape and np.array_equal(kor, keng)
    assert e.counts() == (o.nlocal, o.nghost)

    def state(obj, is_engine):
        if is_engine:
            g = obj.get_atoms(fields=("x", "f", "tag"))
            order = np.argsort(g["tag"])
            eng, vir = obj.tallies()
            return g["x"][order], g["f"][order], eng, vir, obj.stats()["nbuilds"]
        order = np.argsort(obj.tag())
        return obj.x()[order], obj.f()[order], obj.eng_vdwl, obj.virial, obj.ncalls

    for nsteps, xtol, ftol, etol in ((0, 1e-13, 1e-12, 1e-12), (60, 1e-9, 1e-8, 1e-10)):
        if nsteps:
            o.run(nsteps, 0, nsteps)
            e.run(nsteps, nsteps)
        xo, fo, eo, vo, bo = state(o, False)
        xe, fe, ee, ve, be = state(e, True)
        assert np.abs(xo - xe).max() <= xtol
        assert np.abs(fo - fe).max() <= ftol * np.abs(fo).max()
        assert abs(eo - ee) <= etol * abs(eo)
        assert np.abs(vo - ve).max() <= max(etol, 1e-11) * np.abs(vo).max()
        assert bo == be
    e.close()


def test_triclinic_mixed_precision_meets_the_mixed_tolerances():
    """FP32 pair math (fixed-point staged positions, sub-domain-wide records) in a prism box:
    forces <= 1e-5, energy <= 1e-6 against the double-precision path (north_star's mixed bar)"""
    from lammps_b200 import pair_lj
    from lammps_b200.engine import Engine
    st = _reference_state(_script(LJ_TRI, 10, "-4.0 3.0 -2.0", "every 1 delay 0 check yes"), 40)
    nl = st["nlocal"]
    res = {}
    for prec in ("double", "mixed"):
        e = Engine(0, prec, "lj")
        e.set_box_triclinic(st["lo"], st["hi"], st["xy"], st["xz"], st["yz"])
        e.set_atoms(st["x"][:nl], st["v"], st["type"], st["tag"][:nl], np.array([0.0, 1.0]), image=st["image"])
        e.neighbor(0.3, every=1, delay=0, check=True)
        e.fix_nve(0.005)
        e.pair_lj_cut(pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5))
        e.setup(1, 1)
        assert e.stats()["list_kind"] == 1
        a = e.get_atoms(fields=("f", "tag"))
        eng, vir = e.tallies()
        th = e.run(40, 20)
        res[prec] = (a["f"][np.argsort(a["tag"])], eng, vir, th)
        e.close()
    fd, fm = res["double"][0], res["mixed"][0]
    assert np.abs(fd - fm).max() / np.abs(fd).max() <= 1e-5
    assert abs(res["double"][1] - res["mixed"][1]) <= 1e-6 * abs(res["double"][1])
    assert np.abs(res["double"][2] - res["mixed"][2]).max() <= 1e-6 * np.abs(res["double"][2]).max()
    td, tm = res["double"][3], res["mixed"][3]
    scale = np.maximum(np.abs(td).max(axis=0), 1e-3)
    assert (np.abs(td - tm).max(axis=0) <= 1e-5 * scale).all(), np.abs(td - tm).max(axis=0) / scale


def _run(exe, args, d, body, ncols):
    d.mkdir()
    (d / "in.t").write_text(body)
    r = subprocess.run([str(exe), *args, "-in", "in.t"], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows, on = [], False
    for ln in r.stdout.splitlines():
        if re.match(r"\s*Step\s+Temp\s+PotEng", ln):
            on = True
            continue
        if ln.startswith("Loop time"):
            on = False
        f = ln.split()
        if on and len(f) == ncols and re.fullmatch(r"\d+", f[0]):
            rows.append([float(t) for t in f])
    last = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[-1].splitlines()
    k = last.index([ln for ln in last if ln.startswith("ITEM: ATOMS")][0])
    dump = np.array([[float(t) for t in ln.split()] for ln in last[k + 1:] if ln.strip()])
    return np.array(rows), dump, r.stdout


@pytest.mark.parametrize("kind,cells,tilt,neigh,extra", [
    ("lj", 10, "2.0 -1.0 3.0", "every 20 delay 0 check no", []),
    ("lj", 10, "-4.0 3.0 -2.0", "every 1 delay 0 check yes", []),
    ("lj", 10, "2.0 -1.0 3.0", "every 2 delay 0 check yes", ["-pk", "b200", "subdomains", "8"]),
    ("eam", 8, "1.5 -2.0 1.0", "every 1 delay 5 check yes", []),
    ("eam", 8, "1.5 -2.0 1.0", "every 1 delay 5 check yes", ["-pk", "b200", "subdomains", "4"]),
    ("lj", 10, "-4.0 3.0 -2.0", "every 1 delay 0 check yes", ["-pk", "b200", "list", "flat"]),
    ("eam", 8, "1.5 -2.0 1.0", "every 1 delay 5 check yes", ["-pk", "b200", "list", "flat"]),
], ids=["lj-check-no", "lj-check-yes", "lj-8-subdomains", "eam", "eam-4-subdomains", "lj-flat-list",
        "eam-flat-list"])
def test_triclinic_run_matches_reference_executable(tmp_path, kind, cells, tilt, neigh, extra):
    body = _script(LJ_TRI if kind == "lj" else EAM_TRI, cells, tilt, neigh) + """
compute pea all pe/atom
thermo 1
thermo_style custom step temp pe etotal press pxy pxz pyz
thermo_modify format float %.12g
dump 1 all custom 60 f.dump id x y z vx fx fy fz c_pea
dump_modify 1 sort id format float %.10g
run 60
"""
    ta, da, oa = _run(REF, [], tmp_path / "ref", body, 8)
    tb, db, ob = _run(EXE, ["-sf", "b200", *extra], tmp_path / "b200", body, 8)
    assert "triclinic box" in oa and ta.shape == tb.shape == (61, 8)
    scale = np.maximum(np.abs(ta).max(axis=0), 1e-3)
    assert (np.abs(ta - tb).max(axis=0) <= 1e-9 * scale).all(), np.abs(ta - tb).max(axis=0) / scale
    assert np.array_equal(da[:, 0], db[:, 0])
    # positions: the dump wraps an atom sitting within rounding of a face either way
    assert np.abs(da[:, 1:4] - db[:, 1:4]).max() <= 1e-8
    assert np.abs(da[:, 4] - db[:, 4]).max() <= 1e-8 * max(1.0, np.abs(da[:, 4]).max())
    fs = np.abs(da[:, 5:8]).max()
    assert np.abs(da[:, 5:8] - db[:, 5:8]).max() <= 1e-8 * fs
    assert np.abs(da[:, 8] - db[:, 8]).max() <= 1e-8 * np.abs(da[:, 8]).max()
    for pat in (r"Neighbor list builds = (\d+)", r"Total # of neighbors = (\d+)"):
        m = re.search(pat, oa)
        assert m and m.group(0) in ob, (m.group(0), re.search(pat, ob).group(0))


def test_changing_triclinic_box_is_refused(tmp_path):
    d = tmp_path / "npt"
    d.mkdir()
    (d / "in.t").write_text(_script(LJ_TRI, 6, "1.0 0.5 -0.5", "every 1 delay 0 check yes")
                            .replace("fix 1 all nve", "fix 1 all npt temp 1.0 1.0 0.5 iso 1.0 1.0 5.0") + "run 5\n")
    r = subprocess.run([str(EXE), "-sf", "b200", "-in", "in.t"], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "orthogonal box" in (r.stdout + r.stderr)
