"""fix langevin on the device (FixLangevin::post_force, fix_langevin.cpp:383-507).

* the kernel against the numpy restatement (oracle.langevin_post_force) with the same uniforms:
  bit for bit, both for host-supplied uniforms and for the device's own Philox stream (restated in
  oracle.philox4x32_10, itself pinned to the published Random123 known-answer vectors);
* lmp_b200 -sf b200 against the compiled reference: exact (1e-9) when the host draws the
  reference's own RanMars stream in tag order (`package b200 langevin_rng host`), including
  zero yes and a temperature ramp; statistical with the device stream;
* the device stream does not depend on the decomposition: 1 and 8 sub-domains give one trajectory."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from common import by_tag, lj_system, make_engine, melted

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "lammps_b200" / "lammps_pkg" / "lmp_b200"
REF = ROOT / "oracle" / "_ref" / "lmp_ref"


def test_philox_restatement_known_answers():
    from oracle.oracle import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(w[0]) for w in got) == want


@pytest.mark.parametrize("source", ["host-uniforms", "device-stream"])
def test_langevin_kernel_equals_restatement_bitwise(source):
    from oracle import oracle as O
    s = melted(lj_system((6, 6, 6)), 40)
    e = make_engine(s)
    e.setup(1, 1)
    a = e.get_atoms(fields=("v", "f", "tag", "type"))
    tag, v0, f0, typ = a["tag"], a["v"], a["f"], a["type"]
    n = len(tag)
    g1, g2 = O.langevin_prefactors(s["mass"], t_period=0.7, dt=s["dt"], boltz=1.0, ftm2v=1.0, mvv2e=1.0)
    tsqrt = np.sqrt(1.1)
    seed, step = 48279, 12345678901
    if source == "host-uniforms":
        rng = np.random.default_rng(5)
        ubytag = rng.random((n, 3))
        fs = e.langevin(g1, g2 * tsqrt, seed, step, uniforms_by_tag=ubytag, want_fsum=True)
        u = ubytag[tag - 1]
    else:
        fs = e.langevin(g1, g2 * tsqrt, seed, step, want_fsum=True)
        u = O.langevin_device_uniforms(tag, seed, step)
        assert 0.0 < u.min() and u.max() < 1.0
    f1 = e.get_atoms(fields=("f",))["f"]
    want = O.langevin_post_force(f0, v0, typ, g1, g2, tsqrt, u)
    assert np.array_equal(f1, want), np.abs(f1 - want).max()
    fran = (g2[typ] * tsqrt)[:, None] * (u - 0.5)
    assert np.abs(fs - fran.sum(0)).max() <= 1e-12 * np.abs(fran).sum()
    # zero yes: subtract the group mean of the random force
    e.add_force(-fs / n)
    f2 = e.get_atoms(fields=("f",))["f"]
    assert np.array_equal(f2, f1 + (-fs / n))
    assert np.abs((f2 - (f0 + g1[typ][:, None] * v0)).sum(0)).max() < 1e-9
    e.close()


BODY = """
units lj
atom_modify sort 0 0
lattice fcc 0.8442
region box block 0 10 0 10 0 10
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify NEIGH
fix 1 all nve
fix 2 all langevin LANGEVIN
thermo THERMO
thermo_style custom step temp pe etotal press
thermo_modify format float %.12g
dump 1 all custom NSTEPS f.dump id x y z vx fx
dump_modify 1 sort id format float %.10g
run NSTEPS
"""


def _run(exe, args, d, body):
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
        if on and len(f) == 5 and re.fullmatch(r"\d+", f[0]):
            rows.append([float(t) for t in f])
    last = (d / "f.dump").read_text().split("ITEM: TIMESTEP")[-1].splitlines()
    k = last.index([ln for ln in last if ln.startswith("ITEM: ATOMS")][0])
    dump = np.array([[float(t) for t in ln.split()] for ln in last[k + 1:] if ln.strip()])
    return np.array(rows), dump, r.stdout


def _body(langevin, neigh="every 20 delay 0 check no", nsteps=60, thermo=1):
    return (BODY.replace("LANGEVIN", langevin).replace("NEIGH", neigh).replace("NSTEPS", str(nsteps))
            .replace("THERMO", str(thermo)))


@pytest.mark.parametrize("langevin,neigh,extra", [
    ("1.44 0.6 0.5 48279", "every 20 delay 0 check no", []),
    ("1.0 1.0 0.2 9127 zero yes", "every 1 delay 0 check yes", []),
    ("0.0 0.0 1.0 77", "every 20 delay 0 check no", []),
    ("1.2 0.9 0.5 33 zero yes", "every 5 delay 0 check yes", ["subdomains", "8"]),
], ids=["ramp", "zero-yes-check-yes", "drag-only", "8-subdomains-zero-yes"])
def test_fix_langevin_reproduces_reference_with_its_own_random_stream(tmp_path, langevin, neigh, extra):
    """`langevin_rng host`: the uniforms are the reference's RanMars sequence in tag order (its host
    order under atom_modify sort 0 0), the arithmetic is the device kernel's: thermo every step and
    the final x, v, f equal lmp_ref's."""
    body = _body(langevin, neigh)
    ta, da, oa = _run(REF, [], tmp_path / "ref", body)
    tb, db, ob = _run(EXE, ["-sf", "b200", "-pk", "b200", "langevin_rng", "host", *extra], tmp_path / "b200", body)
    assert ta.shape == tb.shape == (61, 5)
    scale = np.maximum(np.abs(ta).max(axis=0), 1e-3)
    assert (np.abs(ta - tb).max(axis=0) <= 1e-9 * scale).all(), np.abs(ta - tb).max(axis=0) / scale
    assert np.array_equal(da[:, 0], db[:, 0])
    assert np.abs(da[:, 1:5] - db[:, 1:5]).max() <= 1e-8
    assert np.abs(da[:, 5] - db[:, 5]).max() <= 1e-8 * np.abs(da[:, 5]).max()
    m = re.search(r"Neighbor list builds = (\d+)", oa)
    assert m and ("Neighbor list builds = " + m.group(1)) in ob


def test_fix_langevin_device_stream_thermostats_like_the_reference(tmp_path):
    """the production stream (Philox on the device) is a different sequence with the same
    statistics: 4000 atoms quenched from T = 1.44 towards 0.7 follow the reference's temperature
    curve within the thermal noise, and two runs give the same trajectory to the last digit"""
    body = _body("0.7 0.7 0.5 48279", nsteps=600, thermo=20)
    ta, _, _ = _run(REF, [], tmp_path / "ref", body)
    tb, db, _ = _run(EXE, ["-sf", "b200"], tmp_path / "b200", body)
    tc, dc, _ = _run(EXE, ["-sf", "b200"], tmp_path / "b200again", body)
    assert ta.shape == tb.shape
    assert np.abs(ta[:, 1] - tb[:, 1]).max() < 0.05          # whole curve, sigma_T ~ 0.013 at N = 4000
    assert abs(ta[-10:, 1].mean() - tb[-10:, 1].mean()) < 0.02
    assert abs(tb[-10:, 1].mean() - 0.7) < 0.04
    assert np.array_equal(tb, tc) and np.array_equal(db, dc)


def test_fix_langevin_device_stream_is_decomposition_independent(tmp_path):
    body = _body("1.2 0.8 0.5 31337 zero yes", neigh="every 5 delay 0 check yes")
    ta, da, _ = _run(EXE, ["-sf", "b200"], tmp_path / "one", body)
    tb, db, _ = _run(EXE, ["-sf", "b200", "-pk", "b200", "subdomains", "8"], tmp_path / "eight", body)
    scale = np.maximum(np.abs(ta).max(axis=0), 1e-3)
    assert (np.abs(ta - tb).max(axis=0) <= 1e-9 * scale).all()
    assert np.abs(da[:, 1:5] - db[:, 1:5]).max() <= 1e-8


def test_fix_langevin_options_without_a_device_version_are_refused(tmp_path):
    for bad, msg in (("1.0 1.0 0.5 5 tally yes", "tally"),):
        d = tmp_path / msg
        d.mkdir()
        (d / "in.t").write_text(_body(bad))
        r = subprocess.run([str(EXE), "-sf", "b200", "-in", "in.t"], cwd=d, capture_output=True, text=True,
                           timeout=300)
        assert r.returncode != 0 and msg in (r.stdout + r.stderr)
