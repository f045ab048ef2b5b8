#!/usr/bin/env python3
"""4 M-atom LJ melt in a prism box (tilt 2, -1, 3 lattice constants) vs the orthogonal box:
atom-steps/s and per-phase device times over 100 steps (one GPU)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from lammps_b200.engine import Engine  # noqa: E402

for tilt in (None, (2.0, -1.0, 3.0)):
    s = bench.build_system("lj", (100, 100, 100))
    n = len(s["x"])
    e = Engine(0, "double", s["units"])
    bench.configure(e, s, s["x"], s["v"], s["type"], np.arange(1, n + 1, dtype=np.int32), n)
    if tilt:
        a = (s["hi"][0] - s["lo"][0]) / 100
        e.set_box_triclinic(s["lo"], s["hi"], tilt[0] * a, tilt[1] * a, tilt[2] * a)
    e.setup(1, 1)
    e.run(20, 0)
    e.run(100, 0)
    ms = e.last_run_ms()
    e.profiling(True)
    e.run(100, 0)
    ph = e.phase_times()
    st = e.stats()
    print("tilt", tilt, f"{n * 100 / (ms * 1e-3):.4g} atom-steps/s", {k: (round(t * 1e3 / max(c, 1), 1), c) for k, (t, c) in ph.items() if c},
          "tiles", st["tiles_interior"] + st["tiles_boundary"], "stage_max", st["tile_stage_max"])
    e.close()
