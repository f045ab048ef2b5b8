#!/bin/bash
# compute-sanitizer passes over the tile kernels (lj/cut FP64 + mixed, list build) and the
# peer-memory halo between sub-domains sharing one GPU: memcheck and racecheck, small systems
tag=${1:-r02}
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from common import lj_system, eam_system, make_engine
import numpy as np
which = sys.argv[1]
if which == "lj":
    for prec in ("double", "mixed"):
        e = make_engine(lj_system((8, 8, 8)), prec); e.setup(1, 1); e.run(25, 10); print(prec, e.stats()["npairs"]); e.close()
elif which == "eam":
    e = make_engine(eam_system((6, 6, 6))); e.setup(1, 1); e.run(12, 6); print(e.stats()["npairs"]); e.close()
else:
    from lammps_b200.engine import EngineGroup
    s = lj_system((10, 10, 10)); n = len(s["x"])
    g = EngineGroup([0] * 4, "double", s["units"]); g.set_box(s["lo"], s["hi"])
    g.set_atoms(s["x"], s["v"], s["type"], s["tag"], s["mass"])
    g.neighbor(s["skin"], every=10, delay=0, check=False); g.fix_nve(s["dt"]); g.pair_lj_cut(s["tables"])
    g.setup(1, 1); g.run(25, 0); print(g.stats()["npairs"], g.counts()); g.close()
PY
for tool in memcheck racecheck; do
  for c in lj eam group; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py $c > gpurun_out/${tag}_sanitizer_${tool}_${c}.txt 2>&1
    echo "== $tool $c: $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${tag}_sanitizer_${tool}_${c}.txt | tail -1)"
  done
done
