#!/bin/bash
# compute-sanitizer passes over the tile kernels (lj/cut FP64 + mixed, list build) and the
# peer-memory halo between sub-domains sharing one GPU: memcheck and racecheck, small systems
tag=${1:-r02b}
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from common import lj_system, eam_system, make_engine
import numpy as np
which = sys.argv[1]
import ctypes as C
if which == "lj":
    for prec in ("double", "mixed"):
        e = make_engine(lj_system((8, 8, 8)), prec); e.setup(1, 1); e.run(25, 10)
        if prec == "double":
            e.pair_peratom()
            # the staged integrator pieces of fix nvt/b200
            L, h = e.L, e.h
            L.b200_scale_v(h, C.c_double(0.999), C.c_int(1)); L.b200_nve_v(h, C.c_double(0.0025), C.c_int(1))
            L.b200_nve_x(h, C.c_double(0.005), C.c_int(1))
        print(prec, e.stats()["npairs"]); e.close()
    # two atom types, exclusions, a download before fused steps, the warp-per-bin build
    import os
    from lammps_b200 import pair_lj
    s = lj_system((8, 8, 8)); n = len(s["x"])
    s["type"] = (1 + (np.arange(n) % 3 == 0)).astype(np.int32); s["mass"] = np.array([0.0, 1.0, 1.5])
    s["tables"] = pair_lj.lj_cut_tables(2, {(1, 1): (1.0, 1.0, 2.5), (2, 2): (0.8, 1.1, 2.2), (1, 2): (0.9, 1.05, 2.4)}, 2.5)
    for b2 in ("0", "1"):
        os.environ["B200_BUILD2"] = b2
        e = make_engine(s); e.neigh_modify(exclude_types=[(1, 2)], ntypes=2); e.setup(1, 1)
        e.get_atoms(ghosts=True); e.run(25, 10); print("two types build2", b2, e.stats()["npairs"]); e.close()
elif which == "eam":
    import os
    for env in ("auto", "0"):   # tile kernels (kernels_eam2.cuh), then the flat half-list kernels
        os.environ["B200_EAM2"] = env
        e = make_engine(eam_system((6, 6, 6))); e.setup(1, 1); e.run(12, 6); e.pair_peratom()
        print(env, e.stats()["list_kind"], e.stats()["npairs"]); e.close()
elif which == "new":
    # round-2 additions: triclinic box (bin tiles, flat list, 4 lamda bricks), newton off,
    # neigh_modify exclude group, fix langevin (device stream, host uniforms, zero yes)
    from lammps_b200.engine import EngineGroup
    s = lj_system((8, 8, 8)); n = len(s["x"])
    a = (s["hi"][0] - s["lo"][0]) / 8
    mask = (1 | np.where(np.arange(n) % 2 == 0, 2, 4)).astype(np.int32)
    for mode in ("tile", "flat"):
        e = make_engine(s); e.set_box_triclinic(s["lo"], s["hi"], 2 * a, -a, 3 * a); e.set_option("list", mode)
        e.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"], mask=mask)
        e.neigh_modify_groups([(2, 4)])
        if mode == "tile": e.set_newton(False)
        e.setup(1, 1); e.run(25, 10)
        g1 = np.array([0.0, -1.0]); g2 = np.array([0.0, 3.0])
        e.langevin(g1, g2, 77, 5, want_fsum=True); e.langevin(g1, g2, 77, 6, uniforms_by_tag=np.random.rand(n, 3))
        e.add_force(np.array([1e-3, 0, 0])); e.nve_v(0.0025)
        print("tri", mode, e.stats()["npairs"], e.stats()["list_kind"]); e.close()
    g = EngineGroup([0] * 4, "double", s["units"]); g.set_box_triclinic(s["lo"], s["hi"], 2 * a, -a, 3 * a)
    g.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"], mask=mask)
    g.neighbor(s["skin"], every=5, delay=0, check=True); g.fix_nve(s["dt"]); g.pair_lj_cut(s["tables"])
    for sub in g.sub: sub.neigh_modify_groups([(2, 4)])
    g.setup(1, 1); g.run(25, 0); print("tri group", g.stats()["npairs"], g.counts()); g.close()
else:
    from lammps_b200.engine import EngineGroup
    s = lj_system((10, 10, 10)); n = len(s["x"])
    g = EngineGroup([0] * 4, "double", s["units"]); g.set_box(s["lo"], s["hi"])
    g.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"])
    g.neighbor(s["skin"], every=10, delay=0, check=False); g.fix_nve(s["dt"]); g.pair_lj_cut(s["tables"])
    g.setup(1, 1); g.run(25, 0); g.pair_peratom(); print(g.stats()["npairs"], g.counts()); g.close()
PY
for tool in memcheck racecheck; do
  for c in ${2:-lj eam group new}; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py $c > gpurun_out/${tag}_sanitizer_${tool}_${c}.txt 2>&1
    echo "== $tool $c: $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${tag}_sanitizer_${tool}_${c}.txt | tail -1)"
  done
done
