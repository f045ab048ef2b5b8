"""Shared builders for the parity tests: the two reference workloads (LJ melt, EAM Cu) at
test sizes, configured identically on the oracle (oracle/md_oracle.c) and on the CUDA engine."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from lammps_b200 import eam as eam_mod  # noqa: E402
from lammps_b200 import lattice, pair_lj, units  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


def lj_system(ncell=(10, 10, 10), seed=87287, temp=1.44):
    x, lo, hi = lattice.fcc_block("lj", 0.8442, ncell)
    n = len(x)
    mass = np.array([0.0, 1.0])
    typ = np.ones(n, np.int32)
    v = lattice.velocity_create(x, typ, mass, temp, seed, "lj")
    return dict(kind="lj", units="lj", x=x, v=v, type=typ, tag=np.arange(1, n + 1, dtype=np.int32),
                mass=mass, lo=lo, hi=hi, skin=0.3, every=20, delay=0, check=False, dt=0.005,
                tables=pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5))


def cu_funcfl_path():
    p = GOLDEN / "Cu_u3_funcfl.npz"
    return p


def eam_tables():
    """Cu_u3 funcfl data from the committed fixture (tests/golden/make_golden.py)."""
    d = np.load(cu_funcfl_path())
    f = eam_mod.Funcfl(float(d["mass"]), int(d["nrho"]), float(d["drho"]), int(d["nr"]),
                       float(d["dr"]), float(d["cut"]), d["frho"], d["zr"], d["rhor"])
    return eam_mod.funcfl_tables([f], [0])


def eam_system(ncell=(8, 8, 8), seed=376847, temp=1600.0):
    T = eam_tables()
    x, lo, hi = lattice.fcc_block("metal", 3.615, ncell)
    n = len(x)
    typ = np.ones(n, np.int32)
    v = lattice.velocity_create(x, typ, T.mass, temp, seed, "metal")
    return dict(kind="eam", units="metal", x=x, v=v, type=typ,
                tag=np.arange(1, n + 1, dtype=np.int32), mass=T.mass, lo=lo, hi=hi, skin=1.0,
                every=1, delay=5, check=True, dt=0.005, tables=T.as_dict())


def configure(obj, s, is_engine):
    """Same calls on Oracle and Engine (their method names mirror each other)."""
    obj.set_box(s["lo"], s["hi"])
    obj.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"])
    if is_engine:
        obj.neighbor(s["skin"], every=s["every"], delay=s["delay"], check=s["check"])
        obj.fix_nve(s["dt"])
    else:
        obj.set_neighbor(s["skin"], every=s["every"], delay=s["delay"], check=s["check"])
        obj.fix_nve(s["dt"], units.get(s["units"]).ftm2v)
    if s["kind"] == "lj":
        obj.pair_lj_cut(s["tables"])
    else:
        obj.pair_eam(s["tables"])


def make_oracle(s):
    from oracle.oracle import Oracle
    o = Oracle()
    configure(o, s, False)
    return o


def make_engine(s, precision="double"):
    from lammps_b200.engine import Engine
    e = Engine(0, precision, s["units"])
    configure(e, s, True)
    return e


def melted(s, nsteps=60):
    """Advance the system on the oracle so forces are O(1) (a perfect lattice has f ~ 1e-13)."""
    o = make_oracle(s)
    o.setup(0, 0)
    o.run(nsteps)
    order = np.argsort(o.tag())
    s2 = dict(s)
    s2["x"] = o.x()[order]
    s2["v"] = o.v()[order]
    s2["image"] = o.image()[order]
    return s2


def by_tag(tag, *arrays):
    order = np.argsort(tag)
    return [a[order] for a in arrays]
