"""Generates the committed fixtures in tests/golden/ from the reference itself.
Run in the build container (needs /root/reference and oracle/_ref built):

    python tests/golden/make_golden.py

Fixtures:
  Cu_u3_funcfl.npz     the numeric content of potentials/Cu_u3.eam (funcfl) as arrays
  ref_lj_32k.json      thermo at steps 0/100 + neighbor statistics of bench/in.lj from lmp_ref,
                       plus the values printed in bench/log.15Jul25.lj.fixed.g++.1
  ref_eam_32k.json     same for bench/in.eam (steps 0/50/100)
  ref_lj_melt_4k.npz   a melted 4000-atom LJ state from the reference: x, v, f by tag, pe,
                       virial-derived pressure, and the canonical neighbour-pair keys
  ref_eam_melt_2k.npz  same for a 2048-atom Cu EAM state
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from lammps_b200 import eam  # noqa: E402
from oracle import ref_harness as R  # noqa: E402
from oracle.oracle import canonical_pairs_box  # noqa: E402

OUT = Path(__file__).resolve().parent


def funcfl_fixture():
    f = eam.read_funcfl(str(R.POTENTIALS / "Cu_u3.eam"))
    np.savez_compressed(OUT / "Cu_u3_funcfl.npz", mass=f.mass, nrho=f.nrho, drho=f.drho, nr=f.nr,
                        dr=f.dr, cut=f.cut, frho=f.frho, zr=f.zr, rhor=f.rhor)


def thermo_fixture(name, script_fn, style, steps):
    rows = []
    with R.RefLammps() as ref:
        ref.commands(script_fn(run=0))
        rows.append(dict(step=0, temp=ref.thermo("temp"), e_pair=ref.thermo("epair"),
                         toteng=ref.thermo("etotal"), press=ref.thermo("press")))
        nghost0 = ref.setting("nghost")
        pi, _ = ref.neighbor_pairs(style)
        npairs0 = int(len(pi))
        done = 0
        for s in steps:
            ref.command(f"run {s - done}")
            done = s
            rows.append(dict(step=s, temp=ref.thermo("temp"), e_pair=ref.thermo("epair"),
                             toteng=ref.thermo("etotal"), press=ref.thermo("press")))
        pi, _ = ref.neighbor_pairs(style)
        out = dict(natoms=ref.natoms(), thermo=rows, nghost_step0=nghost0, npairs_step0=npairs0,
                   nghost_end=ref.setting("nghost"), npairs_end=int(len(pi)))
    return out


def melt_fixture(fname, script, style, nsteps):
    with R.RefLammps() as ref:
        ref.commands(script)
        ref.command(f"run {nsteps}")
        nl, ng = ref.setting("nlocal"), ref.setting("nghost")
        x = ref.atom_vec3("x", nl + ng)
        v = ref.atom_vec3("v", nl)
        f = ref.atom_vec3("f", nl)
        tag = ref.atom_int("id", nl + ng)
        img = ref.atom_int("image", nl)
        lo, hi = ref.box()
        # the list in memory was built at the last reneighbor step with the positions of that
        # step; rebuild now so list and positions belong together: run 0 re-runs setup
        ref.command("run 0")
        nl, ng = ref.setting("nlocal"), ref.setting("nghost")
        x = ref.atom_vec3("x", nl + ng)
        v = ref.atom_vec3("v", nl)
        f = ref.atom_vec3("f", nl)
        tag = ref.atom_int("id", nl + ng)
        img = ref.atom_int("image", nl)
        pi, pj = ref.neighbor_pairs(style)
        keys = canonical_pairs_box(pi, pj, tag, x, lo, hi, nlocal=nl)
        order = np.argsort(tag[:nl])
        np.savez_compressed(OUT / fname, x=x[:nl][order], v=v[order], f=f[order], image=img[order],
                            lo=lo, hi=hi, pe=ref.thermo("pe"), press=ref.thermo("press"),
                            temp=ref.thermo("temp"), nghost=ng, pair_keys=keys.astype(np.int32))


if __name__ == "__main__":
    funcfl_fixture()
    lj = thermo_fixture("lj", R.lj_input, "lj/cut", [100])
    lj["published_log"] = {"file": "bench/log.15Jul25.lj.fixed.g++.1:60-61,79-86",
                           "thermo": [[0, 1.44, -6.7733681, -4.6134356, -5.0197073],
                                      [100, 0.7574531, -5.7585055, -4.6223613, 0.20726105]],
                           "neighbors": 1202833, "nghost": 19657, "builds": 5}
    (OUT / "ref_lj_32k.json").write_text(json.dumps(lj, indent=1))
    ea = thermo_fixture("eam", R.eam_input, "eam", [50, 100])
    ea["published_log"] = {"file": "bench/log.15Jul25.eam.fixed.g++.1:62-64,82-90",
                           "thermo": [[0, 1600, -113280, -106662.09, 18703.573],
                                      [50, 781.69049, -109873.35, -106640.13, 52273.088],
                                      [100, 801.832, -109957.3, -106640.77, 51322.821]],
                           "neighbors": 1207784, "nghost": 19909, "builds": 13, "dangerous": 0}
    (OUT / "ref_eam_32k.json").write_text(json.dumps(ea, indent=1))
    melt_fixture("ref_lj_melt_4k.npz", R.lj_input(run=0, cells=10), "lj/cut", 100)
    melt_fixture("ref_eam_melt_2k.npz", R.eam_input(run=0, cells=8), "eam", 100)
    print("fixtures written to", OUT)
