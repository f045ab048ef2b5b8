"""Development probe: run a workload on one GPU with per-phase CUDA-event timing.
usage: python tools/perf_probe.py [lj|eam] [cells] [steps] [precision]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from common import eam_system, lj_system, make_engine  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "lj"
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 100
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
prec = sys.argv[4] if len(sys.argv) > 4 else "double"

t0 = time.time()
s = (lj_system if kind == "lj" else eam_system)((cells, cells, cells))
n = len(s["x"])
print(f"{kind} natoms={n} host setup {time.time() - t0:.1f}s", flush=True)
e = make_engine(s, prec)
t0 = time.time()
e.setup(1, 1)
print(f"device setup {time.time() - t0:.2f}s  counts={e.counts()} stats={e.stats()}", flush=True)
e.run(20, 0)  # warm-up
e.profiling(True)
t0 = time.time()
th = e.run(steps, 0)
wall = time.time() - t0
ms = e.last_run_ms()
ph = e.phase_times()
print(f"{steps} steps: wall {wall:.3f}s device {ms:.1f} ms -> {n * steps / (ms * 1e-3) / 1e6:.1f} Matom-step/s")
for k, (t, c) in ph.items():
    if c:
        print(f"  {k:18s} {t:10.2f} ms total {c:5d} calls  {t / c * 1e3:10.1f} us/call  {100 * t / ms:5.1f}%")
st = e.stats()
print("stats", st)
row = e.thermo_row(th[-1])
print("thermo", row)
