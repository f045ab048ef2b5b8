"""oracle.langevin_post_force / langevin_prefactors / RanMars (the numpy restatements the CUDA
kernel is compared with bit for bit in tests/test_gpu_langevin.py) pinned against the UNMODIFIED
reference compiled here: forces right after `run 0` of a system with fix langevin are
pair forces + FixLangevin::post_force (FixLangevin::setup, fix_langevin.cpp:295-305), drawn from
RanMars(seed + me) in host atom order = tag order (atom_modify sort 0 0).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_harness as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")

BODY = """
units lj
atom_modify sort 0 0
lattice fcc 0.8442
region box block 0 5 0 5 0 5
create_box 2 box
create_atoms 1 box
set type 1 type/ratio 2 0.3 991
mass 1 1.0
mass 2 2.5
velocity all create 1.44 87287 loop geom
pair_style lj/cut 2.5
pair_coeff * * 1.0 1.0 2.5
neighbor 0.3 bin
fix 1 all nve
{langevin}
thermo 10
run 0
"""


@pytest.mark.parametrize("args,zero", [("1.3 0.7 0.4 48279", False), ("0.9 0.9 2.0 1234 zero yes", True),
                                       ("1.0 1.0 0.5 77 scale 2 1.7", False)], ids=["ramp", "zero-yes", "scale"])
def test_langevin_restatement_matches_the_compiled_reference(args, zero):
    with R.RefLammps() as ref:
        ref.commands(BODY.format(langevin=""))
        n = ref.natoms()
        f_pair = ref.atom_vec3("f", n)
        tag = ref.atom_int("id", n)
        assert np.array_equal(tag, np.arange(1, n + 1))       # host order = tag order
    with R.RefLammps() as ref:
        ref.commands(BODY.format(langevin=f"fix 2 all langevin {args}"))
        f_all = ref.atom_vec3("f", n)
        v = ref.atom_vec3("v", n)
        typ = ref.atom_int("type", n)
    a = args.split()
    t_start, damp, seed = float(a[0]), float(a[2]), int(a[3])
    ratio = [1.0, 1.0, 1.0]
    if "scale" in a:
        ratio[int(a[a.index("scale") + 1])] = float(a[a.index("scale") + 2])
    g1, g2 = O.langevin_prefactors([0.0, 1.0, 2.5], t_period=damp, dt=0.005, boltz=1.0, ftm2v=1.0, mvv2e=1.0,
                                   ratio=ratio)
    # run 0: beginstep == endstep == ntimestep -> delta = 0 -> t_target = t_start (compute_target :519-527)
    u = O.RanMars(seed).uniforms(3 * n).reshape(n, 3)
    want = O.langevin_post_force(f_pair, v, typ, g1, g2, np.sqrt(t_start), u, zero=zero)
    assert np.abs(want - f_all).max() <= 1e-13 * np.abs(f_all).max()
    assert np.abs(f_all - f_pair).max() > 0.1          # the thermostat force is not a rounding effect
