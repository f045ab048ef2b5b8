"""Turns the reference's own known-answer fixtures for this path into small committed files
(run in the build container; needs /root/reference):

    python tests/golden/make_yaml_fixtures.py

  ref_yaml_pair_eam.npz   unittest/force-styles/tests/atomic-pair-eam.yaml (32 atoms, Al_jnp.eam +
                          Cu_u3.eam, epsilon 6e-12): the input state of in.metal/data.metal (box,
                          types, positions, velocities), the numeric content of the two funcfl
                          potential files, and the reference answers init_/run_ vdwl, stress,
                          forces (run_ = after `fix nve` + `run 4`, test_pair_style.cpp:158-160)
  ref_yaml_pair_eam_alloy.npz  atomic-pair-eam_alloy.yaml (`pair_coeff * * CuNi.eam.alloy Cu Ni`,
                          epsilon 5e-12): same input state, the numeric content of the setfl file,
                          and the reference answers
  ref_yaml_pair_eam_fs.npz     atomic-pair-eam_fs.yaml (`pair_coeff * * AlFe_mm.eam.fs Al Fe`,
                          Finnis-Sinclair file with per-element-pair densities)
"""
import sys
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from lammps_b200 import eam  # noqa: E402

REF = Path("/root/reference")
TESTS = REF / "unittest" / "force-styles" / "tests"
OUT = Path(__file__).resolve().parent


def read_data_file(path):
    lines = path.read_text().splitlines()
    box = {}
    for ln in lines:
        t = ln.split()
        if len(t) == 4 and t[2] in ("xlo", "ylo", "zlo"):
            box[t[2][0]] = (float(t[0]), float(t[1]))
    def section(name):
        i = next(k for k, ln in enumerate(lines) if ln.split("#")[0].strip() == name)
        rows = []
        for ln in lines[i + 2:]:
            if not ln.strip():
                break
            rows.append(ln.split())
        return rows
    atoms = section("Atoms")
    vels = section("Velocities")
    tag = np.array([int(r[0]) for r in atoms], np.int32)
    typ = np.array([int(r[1]) for r in atoms], np.int32)
    x = np.array([[float(v) for v in r[2:5]] for r in atoms])
    img = np.array([[int(v) for v in r[5:8]] for r in atoms], np.int32)
    vt = np.array([int(r[0]) for r in vels])
    v = np.zeros_like(x)
    v[np.searchsorted(tag, vt)] = np.array([[float(c) for c in r[1:4]] for r in vels])
    lo = np.array([box[d][0] for d in "xyz"])
    hi = np.array([box[d][1] for d in "xyz"])
    return dict(tag=tag, type=typ, x=x, v=v, image=img, lo=lo, hi=hi)


def block(text, ncol):
    rows = [[float(v) for v in ln.split()] for ln in text.strip().splitlines()]
    a = np.array(rows)
    assert a.shape[1] == ncol
    return a


def answers(out, yaml_name, prefix=""):
    y = yaml.safe_load((TESTS / yaml_name).read_text())
    for pre in ("init", "run"):
        out[f"{prefix}{pre}_vdwl"] = float(y[f"{pre}_vdwl"])
        out[f"{prefix}{pre}_stress"] = block(y[f"{pre}_stress"], 6)[0]
        f = block(y[f"{pre}_forces"], 4)
        assert np.array_equal(f[:, 0].astype(int), np.arange(1, 33))
        out[f"{prefix}{pre}_forces"] = f[:, 1:]
    out[f"{prefix}epsilon"] = float(y["epsilon"])
    return y


def pair_eam_fixture():
    y = yaml.safe_load((TESTS / "atomic-pair-eam.yaml").read_text())
    assert y["pair_style"] == "eam" and y["natoms"] == 32
    d = read_data_file(TESTS / "data.metal")
    order = np.argsort(d["tag"])
    out = {k: (v[order] if k in ("tag", "type", "x", "v", "image") else v) for k, v in d.items()}
    for key, fname in (("al", "Al_jnp.eam"), ("cu", "Cu_u3.eam")):
        f = eam.read_funcfl(str(REF / "potentials" / fname))
        for fld in ("mass", "nrho", "drho", "nr", "dr", "cut", "frho", "zr", "rhor"):
            out[f"{key}_{fld}"] = getattr(f, fld)
    for pre in ("init", "run"):
        out[f"{pre}_vdwl"] = float(y[f"{pre}_vdwl"])
        out[f"{pre}_stress"] = block(y[f"{pre}_stress"], 6)[0]
        f = block(y[f"{pre}_forces"], 4)
        assert np.array_equal(f[:, 0].astype(int), np.arange(1, 33))
        out[f"{pre}_forces"] = f[:, 1:]
    out["epsilon"] = float(y["epsilon"])
    # the same system under `units real` (the readers convert the metal-units files): answers only
    yr = answers(out, "atomic-pair-eam_real.yaml", "real_")
    assert yr["pair_coeff"] == y["pair_coeff"] and "units index real" in yr["pre_commands"]
    np.savez_compressed(OUT / "ref_yaml_pair_eam.npz", **out)


def pair_eam_alloy_fixture(yaml_name="atomic-pair-eam_alloy.yaml", style="eam/alloy",
                           potential="CuNi.eam.alloy", types=("Cu", "Ni"),
                           out_name="ref_yaml_pair_eam_alloy.npz"):
    y = yaml.safe_load((TESTS / yaml_name).read_text())
    assert y["pair_style"] == style and y["natoms"] == 32
    assert y["pair_coeff"].split() == ["*", "*", potential, *types]
    d = read_data_file(TESTS / "data.metal")
    order = np.argsort(d["tag"])
    out = {k: (v[order] if k in ("tag", "type", "x", "v", "image") else v) for k, v in d.items()}
    f = eam.read_setfl(str(REF / "potentials" / potential), fs=style == "eam/fs")
    out.update(elements=np.array(f.elements), type_elements=np.array(types), mass=f.mass,
               nrho=f.nrho, drho=f.drho, nr=f.nr, dr=f.dr, cut=f.cut, frho=f.frho, rhor=f.rhor)
    for (i, j), z in f.z2r.items():
        out[f"z2r_{i}_{j}"] = z
    for pre in ("init", "run"):
        out[f"{pre}_vdwl"] = float(y[f"{pre}_vdwl"])
        out[f"{pre}_stress"] = block(y[f"{pre}_stress"], 6)[0]
        fr = block(y[f"{pre}_forces"], 4)
        assert np.array_equal(fr[:, 0].astype(int), np.arange(1, 33))
        out[f"{pre}_forces"] = fr[:, 1:]
    out["epsilon"] = float(y["epsilon"])
    yr = answers(out, yaml_name.replace(".yaml", "_real.yaml"), "real_")
    assert yr["pair_coeff"] == y["pair_coeff"] and "units index real" in yr["pre_commands"]
    np.savez_compressed(OUT / out_name, **out)


if __name__ == "__main__":
    pair_eam_alloy_fixture()
    pair_eam_alloy_fixture("atomic-pair-eam_fs.yaml", "eam/fs", "AlFe_mm.eam.fs", ("Al", "Fe"),
                           "ref_yaml_pair_eam_fs.npz")
    pair_eam_fixture()
    print("written", OUT / "ref_yaml_pair_eam.npz")
