"""Development probe: time the pair kernel alone (no integration), e.g. for timing experiments
whose arithmetic is deliberately wrong.  usage: pair_only_probe.py [cells] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from common import lj_system, make_engine  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
s = lj_system((cells,) * 3)
e = make_engine(s, "double")
e.setup(0, 0)
for _ in range(5):
    e.pair_compute(0, 0)
e.profiling(True)
for _ in range(reps):
    e.pair_compute(0, 0)
t, c = e.phase_times()["pair"]
print(f"pair-only: {t / c * 1e3:.1f} us/call over {c} calls, natoms={len(s['x'])}")
