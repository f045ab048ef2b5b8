"""The drop-in boundary end to end: the reference's UNMODIFIED bench inputs (bench/in.lj,
bench/in.eam, copied untouched to lammps_pkg/bench_inputs by lammps_pkg/build_pkg.py) run by
lmp_b200 -- the reference LAMMPS with the B200 package compiled in -- with nothing but
`-sf b200` on the command line.  The thermo output must reproduce the reference's own golden
logs (bench/log.15Jul25.{lj,eam}.fixed.g++.1, digits committed in tests/golden/*.json) and
its neighbour statistics."""
import json
import re
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "lammps_b200" / "lammps_pkg" / "lmp_b200"
BENCH = ROOT / "lammps_b200" / "lammps_pkg" / "bench_inputs"
GOLDEN = ROOT / "tests" / "golden"


def run_lmp(args, cwd=BENCH):
    assert EXE.exists(), "lmp_b200 not built (python lammps_b200/lammps_pkg/build_pkg.py)"
    r = subprocess.run([str(EXE), *args], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def thermo_rows(out):
    """[step, temp, e_pair, toteng, press] per thermo line (columns picked by header name)"""
    rows, cols = [], None
    for ln in out.splitlines():
        if re.match(r"\s*Step\s+Temp\s+E_pair", ln):
            names = ln.split()
            cols = [names.index(k) for k in ("Step", "Temp", "E_pair", "TotEng", "Press")]
            continue
        if ln.startswith("Loop time"):
            cols = None
        if cols:
            f = ln.split()
            if len(f) > max(cols) and re.fullmatch(r"\d+", f[0]):
                rows.append([float(f[k]) for k in cols])
    return rows


def close_to_printed(val, printed):
    """equal to the golden log at its printed precision (8 significant digits), +-1 ulp of it"""
    if printed == 0:
        return abs(val) < 1e-7
    return abs(val - printed) <= 1.5e-8 * max(abs(printed), 1e-300) * 10


@pytest.mark.parametrize("name,pair,golden", [("in.lj", "lj/cut/b200", "ref_lj_32k.json"),
                                              ("in.eam", "eam/b200", "ref_eam_32k.json")])
def test_unmodified_bench_input_with_sf_b200(name, pair, golden):
    out = run_lmp(["-sf", "b200", "-in", name])
    g = json.loads((GOLDEN / golden).read_text())["published_log"]
    assert "B200 package: device" in out
    assert "Setting up Verlet/B200 run" in out
    rows = thermo_rows(out)
    assert [int(r[0]) for r in rows] == [int(r[0]) for r in g["thermo"]]
    for got, ref in zip(rows, g["thermo"]):
        for a, b in zip(got[1:], ref[1:]):
            assert close_to_printed(a, b), f"{name} step {int(ref[0])}: {got} vs golden {ref}"
    m = re.search(r"Neighbor list builds = (\d+)", out)
    assert m and int(m.group(1)) == g["builds"]
    m = re.search(r"Dangerous builds = (\d+)", out)
    assert (m and int(m.group(1)) == 0) or "Dangerous builds not checked" in out
    m = re.search(r"Total # of neighbors = (\d+)", out)
    assert m
    # (in.eam runs on the tile kernels of kernels_eam2.cuh: forces summed in a fixed order, so the
    # trajectory -- and with it every skin-shell pair -- is the reference's, not only its printed digits)
    assert int(m.group(1)) == g["neighbors"]
    assert re.search(r"Loop time of [0-9.e+-]+ on 1 procs for 100 steps with 32000 atoms", out)


def test_package_command_and_explicit_styles(tmp_path):
    """`package b200` + explicit /b200 style names instead of the -sf switch; two runs in a
    row (host <-> device round trip of the atoms between runs)."""
    script = tmp_path / "in.explicit"
    script.write_text("""
package b200 prec double profile yes
units lj
atom_style atomic
lattice fcc 0.8442
region box block 0 10 0 10 0 10
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
pair_style lj/cut/b200 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify delay 0 every 20 check no
fix 1 all nve/b200
run_style verlet/b200
thermo 50
run 50
run 50
""")
    out = run_lmp(["-in", str(script)], cwd=tmp_path)
    ref = tmp_path / "in.ref"
    ref.write_text(script.read_text().replace("package b200 prec double profile yes\n", "")
                   .replace("/b200", ""))
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    r = subprocess.run([str(refexe), "-in", str(ref)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0
    a, b = thermo_rows(out), thermo_rows(r.stdout)
    assert len(a) == len(b) == 4
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            assert abs(u - v) <= 2e-7 * max(1.0, abs(v)), (x, y)
    assert "B200 device time by phase" in out


def test_foreign_integrator_is_refused(tmp_path):
    """no silent CPU fallback: a /b200 pair style under plain run_style verlet must error out"""
    script = tmp_path / "in.bad"
    script.write_text("""
units lj
lattice fcc 0.8442
region box block 0 4 0 4 0 4
create_box 1 box
create_atoms 1 box
mass 1 1.0
pair_style lj/cut/b200 2.5
pair_coeff 1 1 1.0 1.0 2.5
fix 1 all nve
run 1
""")
    r = subprocess.run([str(EXE), "-in", str(script)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0
    assert "run_style verlet/b200" in (r.stdout + r.stderr)


def _both(tmp_path, body, nsteps_dump):
    """run the same script with lmp_ref (plain styles) and lmp_b200 (-sf b200); returns the two
    thermo tables and the two id-sorted force dumps of the last step"""
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    outs = {}
    for tag, exe, extra in (("ref", refexe, []), ("b200", EXE, ["-sf", "b200"])):
        d = tmp_path / tag
        d.mkdir()
        (d / "in.t").write_text(body.replace("POT", str(ROOT / "oracle" / "_ref" / "potentials")))
        r = subprocess.run([str(exe), *extra, "-in", "in.t"], cwd=d, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rows = []
        blocks = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[1:]
        last = blocks[-1].splitlines()
        assert int(last[1]) == nsteps_dump
        k = last.index([ln for ln in last if ln.startswith("ITEM: ATOMS")][0])
        for ln in last[k + 1:]:
            if ln.strip():
                rows.append([float(t) for t in ln.split()])
        outs[tag] = (thermo_rows(r.stdout), rows)
    return outs


def _compare(outs, ftol, ttol):
    import numpy as np
    (ta, fa), (tb, fb) = outs["ref"], outs["b200"]
    assert len(ta) == len(tb) > 1
    for x, y in zip(ta, tb):
        for u, v in zip(x, y):
            assert abs(u - v) <= ttol * max(1.0, abs(u)), (x, y)
    fa, fb = np.array(fa), np.array(fb)
    assert fa.shape == fb.shape and np.array_equal(fa[:, 0], fb[:, 0])
    err = np.abs(fa[:, 1:] - fb[:, 1:]).max() / np.abs(fa[:, 1:]).max()
    assert err <= ftol, f"force error vs the reference executable: {err:.2e}"


def test_two_element_eam_matches_reference_executable(tmp_path):
    """Al + Cu funcfl files (the reference's own unit-test pairing, atomic-pair-eam.yaml):
    exercises the per-type-pair rho / z2r table mixing against lmp_ref, forces by atom id
    (the dump prints 10 significant digits, so 1e-8 is the comparison's resolution)."""
    body = """
units metal
lattice fcc 3.8
region box block 0 6 0 6 0 6
create_box 2 box
create_atoms 1 box
set type 1 type/fraction 2 0.4 12345
mass 1 63.55
mass 2 26.98
velocity all create 1200.0 4928459 loop geom
pair_style eam
pair_coeff 1 1 POT/Cu_u3.eam
pair_coeff 2 2 POT/Al_jnp.eam
neighbor 1.0 bin
neigh_modify every 1 delay 5 check yes
fix 1 all nve
thermo 20
thermo_modify format float %.12g
dump 1 all custom 40 f.dump id type fx fy fz
dump_modify 1 sort id format float %.10g
run 40
"""
    _compare(_both(tmp_path, body, 40), ftol=1e-8, ttol=1e-9)


def test_nve_on_a_sub_group_matches_reference_executable(tmp_path):
    """`fix nve` applied to a group only (mask & groupbit, fix_nve.cpp:86) -- the frozen slab
    must not move, the rest follows the reference trajectory."""
    body = """
units lj
lattice fcc 0.8442
region box block 0 8 0 8 0 8
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
region slab block INF INF INF INF 0 2
group frozen region slab
group mobile subtract all frozen
velocity frozen set 0.0 0.0 0.0
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify every 10 delay 0 check no
fix 1 mobile nve
thermo 25
thermo_modify format float %.12g
dump 1 all custom 50 f.dump id x y z fx
dump_modify 1 sort id format float %.10g
run 50
"""
    _compare(_both(tmp_path, body, 50), ftol=1e-8, ttol=1e-9)


LJ_BODY = """
units lj
lattice fcc 0.8442
region box block 0 10 0 10 0 10
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify every 20 delay 0 check no
fix 1 all nve
"""


def _run_b200(tmp_path, body, extra=(), expect_fail=False):
    (tmp_path / "in.t").write_text(body)
    r = subprocess.run([str(EXE), "-sf", "b200", *extra, "-in", "in.t"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600)
    if expect_fail:
        assert r.returncode != 0, "expected an error:\n" + r.stdout[-1500:]
    else:
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout + r.stderr


def test_eight_subdomains_in_one_process_reproduce_the_golden_log():
    """`-pk b200 subdomains 8`: one LAMMPS process, the box split into 2x2x2 brick sub-domains
    with migration, borders and the peer-memory halo between them (here all on one GPU; with
    `gpus 8` one per GPU).  Thermo output and neighbour statistics must still be the reference's."""
    out = run_lmp(["-sf", "b200", "-pk", "b200", "subdomains", "8", "-in", "in.lj"])
    g = json.loads((GOLDEN / "ref_lj_32k.json").read_text())["published_log"]
    assert "8 sub-domains on 1 GPU(s) in this process" in out
    rows = thermo_rows(out)
    assert [int(r[0]) for r in rows] == [int(r[0]) for r in g["thermo"]]
    for got, ref in zip(rows, g["thermo"]):
        for a, b in zip(got[1:], ref[1:]):
            assert close_to_printed(a, b), f"step {int(ref[0])}: {got} vs golden {ref}"
    m = re.search(r"Neighbor list builds = (\d+)", out)
    assert m and int(m.group(1)) == g["builds"]
    m = re.search(r"Total # of neighbors = (\d+)", out)
    assert m and int(m.group(1)) == g["neighbors"]


def test_thermo_every_step_uses_device_sums_and_matches_reference(tmp_path):
    """thermo 1: temp/b200 (device sum of m v^2), pe and pressure from the device tallies -- no atom
    download on thermo steps -- against lmp_ref, every step; then a dump and a restart written
    from device-resident atoms are read back by the reference."""
    body = LJ_BODY + """
thermo 1
thermo_modify format float %.12g
dump 1 all custom 30 f.dump id x y z fx
dump_modify 1 sort id format float %.10g
run 30
write_restart r.restart
"""
    outs = _both(tmp_path, body, 30)
    _compare(outs, ftol=1e-8, ttol=1e-9)
    assert len(outs["b200"][0]) == 31
    # the suffix picked the device-aware computes for the thermo keywords (output.cpp:74-76)
    info = _run_b200(tmp_path / "b200", LJ_BODY + "\ninfo computes\n")
    for style in ("temp/b200", "pe/b200", "pressure/b200"):
        assert "style = " + style in info, info[-1500:]
    # the restart file written by lmp_b200 continues in the reference exactly like its own
    cont = """
read_restart RESTART
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify every 20 delay 0 check no
fix 1 all nve
thermo 10
thermo_modify format float %.12g
run 10
"""
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    rows = {}
    for tag in ("ref", "b200"):
        d = tmp_path / ("cont_" + tag)
        d.mkdir()
        (d / "in.c").write_text(cont.replace("RESTART", str(tmp_path / tag / "r.restart")))
        r = subprocess.run([str(refexe), "-in", "in.c"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        rows[tag] = thermo_rows(r.stdout)
    assert len(rows["ref"]) == len(rows["b200"]) == 2
    for x, y in zip(rows["ref"], rows["b200"]):
        for u, v in zip(x, y):
            assert abs(u - v) <= 1e-9 * max(1.0, abs(u)), (x, y)


@pytest.mark.parametrize("style", ["lj", "eam", "lj-subdomains"])
def test_per_atom_energy_and_stress_match_reference_executable(tmp_path, style):
    """compute pe/atom and compute stress/atom read Pair::eatom / Pair::vatom (pair.cpp:1087-1182):
    under -sf b200 the device fills them on the steps that ask (b200_pair_peratom).  The dump of
    both, sorted by id, must equal the reference executable's, at setup and during the run."""
    import numpy as np
    if style == "eam":
        head = """
units metal
lattice fcc 3.615
region box block 0 7 0 7 0 7
create_box 1 box
create_atoms 1 box
pair_style eam
pair_coeff 1 1 POT/Cu_u3.eam
velocity all create 1600.0 376847 loop geom
neighbor 1.0 bin
neigh_modify every 1 delay 5 check yes
fix 1 all nve
timestep 0.005
"""
    else:
        head = LJ_BODY
    body = head + """
compute pea all pe/atom
compute sa all stress/atom NULL virial
compute sk all stress/atom NULL
compute pes all reduce sum c_pea
thermo 10
thermo_style custom step temp pe c_pes press
thermo_modify format float %.12g
dump 1 all custom 10 f.dump id c_pea c_sa[1] c_sa[2] c_sa[3] c_sa[4] c_sa[5] c_sa[6] c_sk[1] c_sk[4]
dump_modify 1 sort id format float %.12g
run 20
"""
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    dumps = {}
    extra = ["-pk", "b200", "subdomains", "8"] if style == "lj-subdomains" else []
    for tag, exe, args in (("ref", refexe, []), ("b200", EXE, ["-sf", "b200", *extra])):
        d = tmp_path / tag
        d.mkdir()
        (d / "in.t").write_text(body.replace("POT", str(ROOT / "oracle" / "_ref" / "potentials")))
        r = subprocess.run([str(exe), *args, "-in", "in.t"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        blocks = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[1:]
        assert len(blocks) == 3
        per_step = []
        for blk in blocks:
            lines = blk.splitlines()
            k = [i for i, ln in enumerate(lines) if ln.startswith("ITEM: ATOMS")][0]
            per_step.append(np.array([[float(t) for t in ln.split()] for ln in lines[k + 1:] if ln.strip()]))
        dumps[tag] = per_step
    # scale of a column: its largest value over the three dumps, at least that of the diagonal
    # stress (on the perfect lattice of step 0 the per-atom virial is a sum of O(1) pair terms that
    # cancels to ~1e-7 -- eam -- or to rounding noise -- off-diagonal components)
    allref = np.concatenate(dumps["ref"])
    for a, b in zip(dumps["ref"], dumps["b200"]):
        assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])
        for col in range(1, a.shape[1]):
            scale = max(np.abs(allref[:, col]).max(), np.abs(allref[:, 2:5]).max() if col > 1 else 0.0)
            assert np.abs(a[:, col] - b[:, col]).max() <= 1e-8 * scale, (style, col)


def test_centroid_stress_is_refused_not_zero(tmp_path):
    out = _run_b200(tmp_path, LJ_BODY + """
compute cs all centroid/stress/atom NULL virial
dump 1 all custom 10 f.dump id c_cs[1]
run 10
""", expect_fail=True)
    assert "does not provide the per-atom centroid virial" in out


def test_minimize_is_refused_not_silently_wrong(tmp_path):
    out = _run_b200(tmp_path, LJ_BODY + "minimize 1.0e-4 1.0e-6 10 100\n", expect_fail=True)
    assert "computes forces only inside run_style verlet/b200" in out


@pytest.mark.parametrize("extra", [[], ["-pk", "b200", "list", "flat"], ["-pk", "b200", "subdomains", "8"]],
                         ids=["tiles", "flat-list", "8-subdomains"])
def test_neigh_modify_exclude_group_matches_reference_executable(tmp_path, extra):
    """neigh_modify exclude group (NPair::exclusion, npair.cpp:249-254): two interleaved groups that
    do not see each other plus a group excluded from itself; with 8 sub-domains the group masks of
    remote ghosts travel with the border exchange.  Thermo, forces and the neighbour count vs lmp_ref."""
    body = LJ_BODY.replace("fix 1 all nve", """group odd id 1:4000:2
group even id 2:4000:2
group slab id 1:600
neigh_modify exclude group odd even
neigh_modify exclude group slab slab
fix 1 all nve""") + """
thermo 10
thermo_modify format float %.12g
dump 1 all custom 40 f.dump id type fx fy fz
dump_modify 1 sort id format float %.10g
run 40
"""
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    outs, texts = {}, {}
    for tag, exe, args in (("ref", refexe, []), ("b200", EXE, ["-sf", "b200", *extra])):
        d = tmp_path / tag
        d.mkdir()
        (d / "in.t").write_text(body)
        r = subprocess.run([str(exe), *args, "-in", "in.t"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        last = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[-1].splitlines()
        k = last.index([ln for ln in last if ln.startswith("ITEM: ATOMS")][0])
        outs[tag] = (thermo_rows(r.stdout), [[float(t) for t in ln.split()] for ln in last[k + 1:] if ln.strip()])
        texts[tag] = r.stdout
    _compare(outs, ftol=1e-8, ttol=1e-9)
    m = re.search(r"Total # of neighbors = (\d+)", texts["ref"])
    assert m and m.group(0) in texts["b200"]


def test_neigh_modify_exclude_type_and_once_match_reference_executable(tmp_path):
    """two atom types that do not see each other (exclude type 1 2), list built once"""
    body = """
units lj
lattice fcc 0.8442
region box block 0 8 0 8 0 8
create_box 2 box
create_atoms 1 box
set type 1 type/ratio 2 0.4 4711
mass 1 1.0
mass 2 1.7
velocity all create 1.44 87287 loop geom
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
pair_coeff 2 2 0.8 1.1 2.2
pair_coeff 1 2 0.9 1.05 2.4
neighbor 0.3 bin
neigh_modify every 5 delay 0 check no exclude type 1 2
fix 1 all nve
thermo 10
thermo_modify format float %.12g
dump 1 all custom 20 f.dump id type fx fy fz
dump_modify 1 sort id format float %.10g
run 20
"""
    outs = _both(tmp_path, body, 20)
    _compare(outs, ftol=1e-8, ttol=1e-9)
    (tmp_path / "once").mkdir()
    out = _run_b200(tmp_path / "once", body.replace("check no exclude type 1 2", "check no once yes")
                    .replace("run 20", "run 10"))
    assert "Neighbor list builds = 0" in out


def test_neighbor_list_overflow_is_reported(tmp_path):
    """B200_ECAPACITY through the package: `neigh_modify one 20` cannot hold a 40-neighbour list"""
    out = _run_b200(tmp_path, LJ_BODY + "neigh_modify one 20 page 2000\nrun 5\n", expect_fail=True)
    assert "Neighbor list overflow, boost neigh_modify one" in out


def test_profile_fills_the_timer_breakdown(tmp_path):
    """`package b200 profile yes` steps stage by stage and stamps Timer::PAIR/NEIGH/COMM/MODIFY
    (verlet.cpp:257-355), so Finish's breakdown -- the reference's profiling surface -- is filled"""
    out = _run_b200(tmp_path, LJ_BODY + "thermo 50\nrun 100\n", extra=("-pk", "b200", "profile", "yes"))
    t = {}
    for sec in ("Pair", "Neigh", "Comm", "Modify"):
        m = re.search(rf"^{sec}\s*\|\s*([0-9.eE+-]+)\s*\|\s*([0-9.eE+-]+)", out, re.M)
        assert m, f"no {sec} line in the timing breakdown:\n" + out[-1500:]
        t[sec] = float(m.group(2))
    assert t["Pair"] > 0 and t["Neigh"] > 0 and t["Modify"] > 0
    assert t["Pair"] > t["Comm"], t
    assert "B200 device time by phase" in out


@pytest.mark.parametrize("neigh,extra", [("every 20 delay 0 check no", []), ("every 1 delay 0 check yes", []),
                                         ("every 5 delay 0 check yes", ["-pk", "b200", "subdomains", "8"]),
                                         ("every 1 delay 0 check yes", ["-pk", "b200", "lazy", "no"])],
                         ids=["check-no", "check-yes", "8-subdomains", "one-kernel-per-loop"])
def test_fix_nvt_matches_reference_executable(tmp_path, neigh, extra):
    """fix nvt under -sf b200 = fix nvt/b200: the reference's own Nose-Hoover chain (FixNH, inherited)
    on the host, its per-atom loops (nve_v, nve_x, nh_v_temp) and the temperature sum on the device.
    Thermo every step (temperature ramp 1.44 -> 0.8, chain of 3) and the final forces against
    lmp_ref; with `check yes` the displacement vote comes from the nve_x kernel."""
    body = LJ_BODY.replace("neigh_modify every 20 delay 0 check no", "neigh_modify " + neigh).replace(
        "fix 1 all nve", "fix 1 all nvt temp 1.44 0.8 0.5 tchain 3") + """
thermo 1
thermo_style custom step temp pe etotal press ecouple
thermo_modify format float %.12g
dump 1 all custom 60 f.dump id x y z fx
dump_modify 1 sort id format float %.10g
run 60
"""
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    import numpy as np
    tabs = {}
    for tag, exe, args in (("ref", refexe, []), ("b200", EXE, ["-sf", "b200", *extra])):
        d = tmp_path / tag
        d.mkdir()
        (d / "in.t").write_text(body)
        r = subprocess.run([str(exe), *args, "-in", "in.t"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rows, on = [], False
        for ln in r.stdout.splitlines():
            if re.match(r"\s*Step\s+Temp\s+PotEng", ln):
                on = True
                continue
            if ln.startswith("Loop time"):
                on = False
            f = ln.split()
            if on and len(f) == 6 and re.fullmatch(r"\d+", f[0]):
                rows.append([float(t) for t in f])
        blocks = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[1:]
        last = blocks[-1].splitlines()
        k = last.index([ln for ln in last if ln.startswith("ITEM: ATOMS")][0])
        dump = np.array([[float(t) for t in ln.split()] for ln in last[k + 1:] if ln.strip()])
        tabs[tag] = (np.array(rows), dump, r.stdout)
    ta, tb = tabs["ref"][0], tabs["b200"][0]
    assert ta.shape == tb.shape == (61, 6)
    assert np.abs(ta - tb).max() <= 1e-9 * np.abs(ta).max(), np.abs(ta - tb).max(axis=0)
    assert abs(ta[-1, 1] - 0.8) < 0.2                    # the thermostat did pull the temperature down
    da, db = tabs["ref"][1], tabs["b200"][1]
    assert np.array_equal(da[:, 0], db[:, 0])
    assert np.abs(da[:, 1:4] - db[:, 1:4]).max() <= 1e-8
    assert np.abs(da[:, 4] - db[:, 4]).max() <= 1e-8 * np.abs(da[:, 4]).max()
    m = re.search(r"Neighbor list builds = (\d+)", tabs["ref"][2])
    assert m and ("Neighbor list builds = " + m.group(1)) in tabs["b200"][2]


@pytest.mark.parametrize("extra", [[], ["-pk", "b200", "subdomains", "8"]], ids=["one-subdomain", "8-subdomains"])
def test_running_ahead_inside_lmp_b200_never_leaks_into_host_reads(tmp_path, monkeypatch, extra):
    """VerletB200::run tells the device when another step follows before any host read
    (b200_step_ahead): the pair kernel then also integrates (fix nve fused, B200_FUSE_MIN=0 turns it
    on at test sizes).  Dump-only steps (15, 30, 45: no thermo, so no tallies), thermo steps and the
    end of the run must all see x(n), v(n), f(n) exactly as the reference executable does."""
    monkeypatch.setenv("B200_FUSE_MIN", "0")
    body = LJ_BODY + """
thermo 20
thermo_modify format float %.12g
dump 1 all custom 15 f.dump id x y z vx fx
dump_modify 1 sort id format float %.10g
run 50
"""
    import numpy as np
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    res = {}
    for tag, exe, args in (("ref", refexe, []), ("b200", EXE, ["-sf", "b200", *extra])):
        d = tmp_path / tag
        d.mkdir()
        (d / "in.t").write_text(body)
        r = subprocess.run([str(exe), *args, "-in", "in.t"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        blocks = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[1:]
        assert [int(b.splitlines()[1]) for b in blocks] == [0, 15, 30, 45]
        per = []
        for blk in blocks:
            lines = blk.splitlines()
            k = [i for i, ln in enumerate(lines) if ln.startswith("ITEM: ATOMS")][0]
            per.append(np.array([[float(t) for t in ln.split()] for ln in lines[k + 1:] if ln.strip()]))
        res[tag] = (thermo_rows(r.stdout), per)
    for x, y in zip(res["ref"][0], res["b200"][0]):
        for u, v in zip(x, y):
            assert abs(u - v) <= 1e-9 * max(1.0, abs(u)), (x, y)
    assert len(res["b200"][0]) == 4
    for a, b in zip(res["ref"][1], res["b200"][1]):
        assert np.array_equal(a[:, 0], b[:, 0])
        assert np.abs(a[:, 1:4] - b[:, 1:4]).max() <= 1e-8       # x
        assert np.abs(a[:, 4] - b[:, 4]).max() <= 1e-8           # vx: full-step velocities, not half-kicked
        assert np.abs(a[:, 5] - b[:, 5]).max() <= 1e-8 * max(np.abs(a[:, 5]).max(), 1.0)


def test_run_continuation_pre_no_keeps_the_device_state(tmp_path):
    """`run N pre no` (run.cpp:169-172): no init, no setup -- Verlet::run continues from the state
    the previous run left, including the age of the neighbour list.  The device keeps atoms, list
    and `ago` across runs, so the rebuild schedule and the trajectory are the reference's; a third
    run with `pre yes` goes through setup again."""
    body = LJ_BODY + """
thermo 15
thermo_modify format float %.12g
run 30
run 30 pre no post no
dump 1 all custom 5 f.dump id x y z fx
dump_modify 1 sort id format float %.10g
run 25
"""
    outs = _both(tmp_path, body, 85)
    _compare(outs, ftol=1e-8, ttol=1e-9)
    assert len(outs["b200"][0]) == len(outs["ref"][0]) >= 8


@pytest.mark.parametrize("fix,neigh", [
    ("npt temp 1.44 1.0 0.5 iso 0.5 2.0 5.0", "every 20 delay 0 check no"),
    ("npt temp 1.2 1.2 0.5 aniso 1.0 1.0 4.0 tchain 2 pchain 2", "every 1 delay 0 check yes"),
    ("nph x 0.0 1.0 5.0 z 2.0 0.0 5.0", "every 2 delay 0 check yes"),
], ids=["npt-iso", "npt-aniso-check-yes", "nph-xz"])
def test_fix_npt_nph_match_reference_executable(tmp_path, fix, neigh):
    """fix npt / nph under -sf b200: FixNH's thermostat chain, barostat equations and box update run
    as the reference's own host code; its per-atom loops (nve_v, nve_x, nh_v_temp, nh_v_press, remap)
    on the device, which adopts the changing box (halo shifts at once, bins at the next rebuild, the
    displacement trigger shrinking with the box corners).  Thermo every step -- temperature, energy,
    pressure, volume, the conserved quantity's coupling term -- and the final state against lmp_ref."""
    body = LJ_BODY.replace("neigh_modify every 20 delay 0 check no", "neigh_modify " + neigh).replace(
        "fix 1 all nve", "fix 1 all " + fix) + """
thermo 1
thermo_style custom step temp pe press vol lx lz ecouple
thermo_modify format float %.12g
dump 1 all custom 50 f.dump id x y z vx fx
dump_modify 1 sort id format float %.10g
run 50
"""
    import numpy as np
    refexe = ROOT / "oracle" / "_ref" / "lmp_ref"
    tabs = {}
    for tag, exe, args in (("ref", refexe, []), ("b200", EXE, ["-sf", "b200"])):
        d = tmp_path / tag
        d.mkdir()
        (d / "in.t").write_text(body)
        r = subprocess.run([str(exe), *args, "-in", "in.t"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rows, on = [], False
        for ln in r.stdout.splitlines():
            if re.match(r"\s*Step\s+Temp\s+PotEng", ln):
                on = True
                continue
            if ln.startswith("Loop time"):
                on = False
            f = ln.split()
            if on and len(f) == 8 and re.fullmatch(r"\d+", f[0]):
                rows.append([float(t) for t in f])
        blocks = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[1:]
        last = blocks[-1].splitlines()
        k = last.index([ln for ln in last if ln.startswith("ITEM: ATOMS")][0])
        dump = np.array([[float(t) for t in ln.split()] for ln in last[k + 1:] if ln.strip()])
        tabs[tag] = (np.array(rows), dump, r.stdout)
    ta, tb = tabs["ref"][0], tabs["b200"][0]
    assert ta.shape == tb.shape == (51, 8)
    scale = np.maximum(np.abs(ta).max(axis=0), 1e-3)
    assert (np.abs(ta - tb).max(axis=0) <= 1e-9 * scale).all(), np.abs(ta - tb).max(axis=0) / scale
    assert abs(ta[-1, 4] - ta[0, 4]) > 1e-3 * ta[0, 4]    # the box did move
    da, db = tabs["ref"][1], tabs["b200"][1]
    assert np.array_equal(da[:, 0], db[:, 0])
    assert np.abs(da[:, 1:4] - db[:, 1:4]).max() <= 1e-8
    assert np.abs(da[:, 4] - db[:, 4]).max() <= 1e-8
    assert np.abs(da[:, 5] - db[:, 5]).max() <= 1e-8 * np.abs(da[:, 5]).max()
    m = re.search(r"Neighbor list builds = (\d+)", tabs["ref"][2])
    assert m and ("Neighbor list builds = " + m.group(1)) in tabs["b200"][2]
