"""The oracle against the UNMODIFIED reference compiled here (oracle/_ref/liblammps_ref.so,
driven through its C library API), on cases the committed fixtures do not cover: two atom
types with mixed and explicit lj/cut coefficients, pair_modify shift, a non-cubic box, and
`check yes` rebuilds.  Skipped where oracle/_ref is not built.  CPU only."""
import numpy as np
import pytest

from common import by_tag, make_oracle
from lammps_b200 import pair_lj
from oracle import ref_harness as R
from oracle.oracle import canonical_pairs_box

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")

SETUP = """
units lj
atom_style atomic
lattice fcc 0.8442
region box block 0 {nx} 0 {ny} 0 {nz}
create_box 2 box
create_atoms 1 box
set type 1 type/ratio 2 0.4 4711
mass 1 1.0
mass 2 1.7
velocity all create 1.44 87287 loop geom
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
pair_coeff 2 2 0.8 1.1 2.2
{cross}
pair_modify shift {shift}
neighbor 0.3 bin
neigh_modify {neigh}
fix 1 all nve
timestep 0.005
thermo 10
run 0
"""


@pytest.mark.parametrize("cross,shift,neigh,every,delay,check", [
    ("pair_coeff 1 2 0.9 1.05 2.4", "no", "delay 0 every 20 check no", 20, 0, False),
    ("", "yes", "delay 0 every 1 check yes", 1, 0, True),     # geometric mixing, shifted energy
])
def test_two_type_lj_matches_the_compiled_reference(cross, shift, neigh, every, delay, check):
    nsteps = 40
    with R.RefLammps() as ref:
        ref.commands(SETUP.format(nx=6, ny=5, nz=7, cross=cross, shift=shift, neigh=neigh))
        n = ref.natoms()
        lo, hi = ref.box()
        x0, v0 = ref.atom_vec3("x", n), ref.atom_vec3("v", n)
        typ, tag = ref.atom_int("type", n), ref.atom_int("id", n)
        f0 = ref.atom_vec3("f", n)
        pe0 = ref.thermo("pe") * n
        nall = n + ref.setting("nghost")
        pi, pj = ref.neighbor_pairs("lj/cut")
        kref = canonical_pairs_box(pi, pj, ref.atom_int("id", nall), ref.atom_vec3("x", nall), lo, hi,
                                   nlocal=n)
        ref.command(f"run {nsteps}")
        x1, f1, tag1 = ref.atom_vec3("x", n), ref.atom_vec3("f", n), ref.atom_int("id", n)
        pe1, press1 = ref.thermo("pe") * n, ref.thermo("press")
    coeffs = {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2)}
    if cross:
        coeffs[(1, 2)] = (0.9, 1.05, 2.4)
    s = dict(kind="lj", units="lj", x=x0, v=v0, type=typ, tag=tag, mass=np.array([0.0, 1.0, 1.7]),
             lo=lo, hi=hi, skin=0.3, every=every, delay=delay, check=check, dt=0.005,
             tables=pair_lj.lj_cut_tables(2, coeffs, 2.5, offset_flag=shift == "yes"))
    o = make_oracle(s)
    o.setup(1, 1)
    pi, pj = o.pairs()
    kor = canonical_pairs_box(pi, pj, o.tag(True), o.x(True), lo, hi, nlocal=o.nlocal)
    assert np.array_equal(kor, kref), "half-list pair sets differ"
    (fo,) = by_tag(o.tag(), o.f())
    (fr,) = by_tag(tag, f0)
    assert np.abs(fo - fr).max() <= 1e-13 * np.abs(fr).max()
    assert abs(o.eng_vdwl - pe0) <= 1e-13 * abs(pe0)
    o.run(nsteps, 0, nsteps)
    xo, fo = by_tag(o.tag(), o.x(), o.f())
    xr, fr = by_tag(tag1, x1, f1)
    prd = hi - lo
    d = xo - xr
    d -= prd * np.rint(d / prd)
    assert np.abs(d).max() <= 1e-11
    assert np.abs(fo - fr).max() <= 1e-9 * np.abs(fr).max()
    assert abs(o.eng_vdwl - pe1) <= 1e-11 * abs(pe1)


def test_nve_on_a_sub_group_matches_the_compiled_reference():
    """`fix nve` applied to a group (FixNVE's `mask[i] & groupbit`, fix_nve.cpp:68-145): atoms
    outside the group feel forces but do not move -- the sub-group aspect of the reference's
    fix-timestep-nve.yaml, on an atomic system"""
    nsteps = 30
    text = SETUP.format(nx=5, ny=5, nz=5, cross="", shift="no", neigh="delay 0 every 5 check no")
    text = text.replace("fix 1 all nve", "group movers id <= 300\nfix 1 movers nve")
    with R.RefLammps() as ref:
        ref.commands(text)
        n = ref.natoms()
        lo, hi = ref.box()
        x0, v0 = ref.atom_vec3("x", n), ref.atom_vec3("v", n)
        typ, tag, mask = ref.atom_int("type", n), ref.atom_int("id", n), ref.atom_int("mask", n)
        ref.command(f"run {nsteps}")
        x1, v1, tag1 = ref.atom_vec3("x", n), ref.atom_vec3("v", n), ref.atom_int("id", n)
    groupbit = 2                       # group `all` is bit 0, the first user group bit 1
    assert set(np.unique(mask)) == {1, 3}
    from oracle.oracle import Oracle
    from lammps_b200 import units
    o = Oracle()
    o.set_box(lo, hi)
    o.set_atoms(x0, v0, typ, tag, np.array([0.0, 1.0, 1.7]), mask=mask)
    o.set_neighbor(0.3, every=5, delay=0, check=False)
    o.fix_nve(0.005, units.get("lj").ftm2v, groupbit)
    o.pair_lj_cut(pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2)}, 2.5))
    o.setup(1, 1)
    o.run(nsteps, 0, 0)
    xo, vo = by_tag(o.tag(), o.x(), o.v())
    xr, vr = by_tag(tag1, x1, v1)
    prd = hi - lo
    d = xo - xr
    d -= prd * np.rint(d / prd)
    assert np.abs(d).max() <= 1e-12 and np.abs(vo - vr).max() <= 1e-11
    frozen = np.sort(tag)[300:] - 1                   # ids > 300 are outside the group
    (xs,) = by_tag(tag, x0)
    assert np.array_equal(xr[frozen], xs[frozen])     # they never moved


PERATOM = """
compute pea all pe/atom
compute sa all stress/atom NULL virial
compute pes all reduce sum c_pea
compute sas all reduce sum c_sa[1] c_sa[2] c_sa[3] c_sa[4] c_sa[5] c_sa[6]
thermo_style custom step pe c_pes c_sas[1] c_sas[2] c_sas[3] c_sas[4] c_sas[5] c_sas[6]
run 0
"""


def _peratom_vs_reference(ref, s, nktv2p):
    """oracle per-atom energy / virial (orc_pair_peratom) against compute pe/atom and
    compute stress/atom NULL virial of the compiled reference (stress = -vatom * nktv2p)"""
    n = ref.natoms()
    tag = ref.atom_int("id", n)
    e_ref = ref.compute_peratom("pea", n)
    s_ref = ref.compute_peratom("sa", n, 6)
    o = make_oracle(s)
    o.setup(1, 1)
    eo, vo = o.pair_peratom()
    eo, vo = by_tag(o.tag(), eo, vo)
    er, sr = by_tag(tag, e_ref, s_ref)
    assert np.abs(eo - er).max() <= 1e-12 * np.abs(er).max()
    assert np.abs(-vo * nktv2p - sr).max() <= 1e-12 * np.abs(sr).max()
    # the per-atom values sum to the global tallies (pair.cpp:1087-1182)
    assert abs(eo.sum() - o.eng_vdwl) <= 1e-11 * abs(o.eng_vdwl)
    assert np.abs(vo.sum(axis=0) - o.virial).max() <= 1e-10 * np.abs(o.virial).max()


def test_per_atom_energy_and_virial_lj_match_the_compiled_reference():
    with R.RefLammps() as ref:
        ref.commands(SETUP.format(nx=5, ny=6, nz=5, cross="pair_coeff 1 2 0.9 1.05 2.4", shift="yes",
                                  neigh="delay 0 every 20 check no"))
        ref.command("run 25")
        ref.commands(PERATOM)
        n = ref.natoms()
        lo, hi = ref.box()
        s = dict(kind="lj", units="lj", x=ref.atom_vec3("x", n), v=ref.atom_vec3("v", n),
                 type=ref.atom_int("type", n), tag=ref.atom_int("id", n), mass=np.array([0.0, 1.0, 1.7]),
                 lo=lo, hi=hi, skin=0.3, every=20, delay=0, check=False, dt=0.005,
                 tables=pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2),
                                                  (1, 2): (0.9, 1.05, 2.4)}, 2.5, offset_flag=True))
        _peratom_vs_reference(ref, s, 1.0)


def test_per_atom_energy_and_virial_eam_match_the_compiled_reference():
    from common import eam_tables
    from lammps_b200 import units
    pot = R.LIB.parent / "potentials" / "Cu_u3.eam"
    if not pot.exists():
        pytest.skip("Cu_u3.eam not staged next to the compiled reference")
    with R.RefLammps() as ref:
        ref.commands(f"""
units metal
atom_style atomic
lattice fcc 3.615
region box block 0 6 0 5 0 5
create_box 1 box
create_atoms 1 box
pair_style eam
pair_coeff 1 1 {pot}
velocity all create 1600.0 376847 loop geom
neighbor 1.0 bin
neigh_modify every 1 delay 5 check yes
fix 1 all nve
timestep 0.005
run 30
""")
        ref.commands(PERATOM)
        n = ref.natoms()
        lo, hi = ref.box()
        T = eam_tables()
        s = dict(kind="eam", units="metal", x=ref.atom_vec3("x", n), v=ref.atom_vec3("v", n),
                 type=ref.atom_int("type", n), tag=ref.atom_int("id", n), mass=T.mass, lo=lo, hi=hi,
                 skin=1.0, every=1, delay=5, check=True, dt=0.005, tables=T.as_dict())
        _peratom_vs_reference(ref, s, units.get("metal").nktv2p)


def test_neigh_modify_exclude_type_and_once_match_the_compiled_reference():
    """neigh_modify exclude type 1 2 (NPair::exclusion, npair.cpp:244-248: the pair never enters
    the list, so unlike atoms do not interact) and once yes (neighbor.cpp:2420: the list of setup
    is kept for the whole run)"""
    nsteps = 12
    text = SETUP.format(nx=5, ny=5, nz=6, cross="pair_coeff 1 2 0.9 1.05 2.4", shift="no",
                        neigh="delay 0 every 2 check no exclude type 1 2 once yes")
    with R.RefLammps() as ref:
        ref.commands(text)
        n = ref.natoms()
        lo, hi = ref.box()
        x0, v0 = ref.atom_vec3("x", n), ref.atom_vec3("v", n)
        typ, tag = ref.atom_int("type", n), ref.atom_int("id", n)
        f0 = ref.atom_vec3("f", n)
        nall = n + ref.setting("nghost")
        pi, pj = ref.neighbor_pairs("lj/cut")
        kref = canonical_pairs_box(pi, pj, ref.atom_int("id", nall), ref.atom_vec3("x", nall), lo, hi, nlocal=n)
        ref.command(f"run {nsteps}")
        x1, tag1 = ref.atom_vec3("x", n), ref.atom_int("id", n)
    coeffs = {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2), (1, 2): (0.9, 1.05, 2.4)}
    s = dict(kind="lj", units="lj", x=x0, v=v0, type=typ, tag=tag, mass=np.array([0.0, 1.0, 1.7]),
             lo=lo, hi=hi, skin=0.3, every=2, delay=0, check=False, dt=0.005,
             tables=pair_lj.lj_cut_tables(2, coeffs, 2.5))
    o = make_oracle(s)
    o.neigh_modify(once=True, exclude_types=[(1, 2)], ntypes=2)
    o.setup(1, 1)
    pi, pj = o.pairs()
    assert not np.any(o.type(True)[pi] != o.type(True)[pj]), "an excluded 1-2 pair is in the list"
    kor = canonical_pairs_box(pi, pj, o.tag(True), o.x(True), lo, hi, nlocal=o.nlocal)
    assert np.array_equal(kor, kref), "half-list pair sets differ"
    (fo,) = by_tag(o.tag(), o.f())
    (fr,) = by_tag(tag, f0)
    assert np.abs(fo - fr).max() <= 1e-13 * np.abs(fr).max()
    o.run(nsteps, 0, 0)
    assert o.ncalls == 0, "once yes: no rebuild after setup"
    (xo,) = by_tag(o.tag(), o.x())
    (xr,) = by_tag(tag1, x1)
    prd = hi - lo
    d = xo - xr
    d -= prd * np.rint(d / prd)
    assert np.abs(d).max() <= 1e-12
