"""The oracle against the reference's own known-answer fixture for this path:
unittest/force-styles/tests/atomic-pair-eam.yaml (test_pair_style.cpp:354-425): 32 atoms of two
elements in a 7 A box (shorter than the ghost cutoff: two layers of periodic images), funcfl
potentials Al_jnp.eam + Cu_u3.eam mixed by PairEAM::file2array, energy / virial / forces at
setup and after `fix nve` + `run 4` with `neigh_modify delay 2 every 2 check no`.
Same comparison as the reference's test: |a - b| <= epsilon * max(|a|, |b|), epsilon = 6e-12
(5x for the forces after the run, test_pair_style.cpp:411).  CPU only: pins oracle/md_oracle.c
and the host-side table construction lammps_b200/eam.py (file2array_funcfl + array2spline)."""
from pathlib import Path

import numpy as np

from common import by_tag, make_oracle
from lammps_b200 import eam as eam_mod
from lammps_b200 import units as units_mod

GOLDEN = Path(__file__).resolve().parent / "golden"


def _system(units="metal"):
    d = np.load(GOLDEN / "ref_yaml_pair_eam.npz")
    c = units_mod.METAL2REAL_ENERGY if units == "real" else 1.0
    files = [eam_mod.Funcfl(float(d[f"{k}_mass"]), int(d[f"{k}_nrho"]), float(d[f"{k}_drho"]),
                            int(d[f"{k}_nr"]), float(d[f"{k}_dr"]), float(d[f"{k}_cut"]),
                            d[f"{k}_frho"], d[f"{k}_zr"], d[f"{k}_rhor"]) for k in ("al", "cu")]
    T = eam_mod.funcfl_tables(files, [0, 1], c)   # pair_coeff 1 1 Al_jnp.eam / 2 2 Cu_u3.eam
    # in.metal: units metal (skin 2.0), neigh_modify delay 2 every 2 check no, timestep 0.0001;
    # PairEAM::coeff sets the masses from the funcfl files
    s = dict(kind="eam", units=units, x=d["x"], v=d["v"], type=d["type"], tag=d["tag"],
             image=d["image"], mass=T.mass, lo=d["lo"], hi=d["hi"], skin=2.0, every=2, delay=2,
             check=False, dt=0.0001, tables=T.as_dict())
    return s, d


def _close(a, b, eps):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.all(np.abs(a - b) <= eps * np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300))


def _check(s, d, pre):
    eps = float(d[pre + "epsilon"])
    o = make_oracle(s)
    o.setup(1, 1)
    (f,) = by_tag(o.tag(), o.f())
    assert _close(f, d[pre + "init_forces"], eps), np.abs(f - d[pre + "init_forces"]).max()
    assert _close(o.eng_vdwl, d[pre + "init_vdwl"], eps)
    assert _close(o.virial, d[pre + "init_stress"], eps)
    o.run(4, 0, 4)                                   # tallies on the last step, like the test
    (f,) = by_tag(o.tag(), o.f())
    assert _close(f, d[pre + "run_forces"], 5 * eps), np.abs(f - d[pre + "run_forces"]).max()
    assert _close(o.eng_vdwl, d[pre + "run_vdwl"], eps)
    assert _close(o.virial, d[pre + "run_stress"], eps)
    return o


def test_pair_eam_real_units_yaml():
    """atomic-pair-eam_real.yaml: the same files under `units real`; the reader converts F(rho)
    and Z(r) (pair_eam.cpp:701-706) and fix nve integrates with the real-units ftm2v"""
    s, d = _system("real")
    _check(s, d, "real_")


def test_pair_eam_yaml_init_and_run():
    s, d = _system()
    eps = float(d["epsilon"])
    o = make_oracle(s)
    o.setup(1, 1)
    assert o.nlocal == 32 and o.nghost > 32 * 7      # more than one layer of images
    (f,) = by_tag(o.tag(), o.f())
    assert _close(f, d["init_forces"], eps), np.abs(f - d["init_forces"]).max()
    assert _close(o.eng_vdwl, d["init_vdwl"], eps)
    assert _close(o.virial, d["init_stress"], eps)
    o.run(4, 0, 4)                                   # tallies on the last step, like the test
    (f,) = by_tag(o.tag(), o.f())
    assert _close(f, d["run_forces"], 5 * eps), np.abs(f - d["run_forces"]).max()
    assert _close(o.eng_vdwl, d["run_vdwl"], eps)
    assert _close(o.virial, d["run_stress"], eps)


def _alloy_system(name="ref_yaml_pair_eam_alloy.npz", units="metal"):
    d = np.load(GOLDEN / name)
    c = units_mod.METAL2REAL_ENERGY if units == "real" else 1.0
    nel = len(d["elements"])
    f = eam_mod.Setfl([str(e) for e in d["elements"]], d["mass"], int(d["nrho"]), float(d["drho"]),
                      int(d["nr"]), float(d["dr"]), float(d["cut"]), d["frho"], d["rhor"],
                      {(i, j): d[f"z2r_{i}_{j}"] for i in range(nel) for j in range(i + 1)})
    T = eam_mod.setfl_tables(f, [str(e) for e in d["type_elements"]], c)   # pair_coeff * * file Cu Ni
    s = dict(kind="eam", units=units, x=d["x"], v=d["v"], type=d["type"], tag=d["tag"],
             image=d["image"], mass=T.mass, lo=d["lo"], hi=d["hi"], skin=2.0, every=2, delay=2,
             check=False, dt=0.0001, tables=T.as_dict())
    return s, d


import pytest  # noqa: E402


@pytest.mark.parametrize("fixture", ["ref_yaml_pair_eam_alloy.npz", "ref_yaml_pair_eam_fs.npz"])
def test_pair_eam_alloy_and_fs_yaml_init_and_run(fixture):
    """atomic-pair-eam_alloy.yaml: setfl file (file2array_setfl, pair_eam.cpp:1211-1325) with the
    element order of the file (Ni, Cu) different from the type order (Cu, Ni), epsilon 5e-12;
    atomic-pair-eam_fs.yaml: Finnis-Sinclair file (file2array_fs, :1331-1456), densities per
    element pair, so rho_i and rho_j of a pair come from different tables"""
    s, d = _alloy_system(fixture)
    eps = float(d["epsilon"])
    o = make_oracle(s)
    o.setup(1, 1)
    (f,) = by_tag(o.tag(), o.f())
    assert _close(f, d["init_forces"], eps), np.abs(f - d["init_forces"]).max()
    assert _close(o.eng_vdwl, d["init_vdwl"], eps)
    assert _close(o.virial, d["init_stress"], eps)
    o.run(4, 0, 4)
    (f,) = by_tag(o.tag(), o.f())
    assert _close(f, d["run_forces"], 5 * eps), np.abs(f - d["run_forces"]).max()
    assert _close(o.eng_vdwl, d["run_vdwl"], eps)
    assert _close(o.virial, d["run_stress"], eps)


@pytest.mark.parametrize("fixture", ["ref_yaml_pair_eam_alloy.npz", "ref_yaml_pair_eam_fs.npz"])
def test_pair_eam_alloy_and_fs_real_units_yaml(fixture):
    """atomic-pair-eam_alloy_real.yaml / atomic-pair-eam_fs_real.yaml (units real)"""
    s, d = _alloy_system(fixture, "real")
    _check(s, d, "real_")
