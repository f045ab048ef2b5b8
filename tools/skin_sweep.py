"""Skin / rebuild-policy sweep on one GPU (BASELINE.json config 3: LJ melt 4M atoms).
usage: python tools/skin_sweep.py [cells] [steps]   -> one line per (skin, policy)"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from common import lj_system, make_engine  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
base = lj_system((cells, cells, cells))
n = len(base["x"])
rows = []
for skin in (0.1, 0.2, 0.3, 0.4, 0.5):
    for policy, every, delay, check in (("every 20 check no", 20, 0, False),
                                        ("every 1 delay 0 check yes", 1, 0, True)):
        s = dict(base, skin=skin, every=every, delay=delay, check=check)
        e = make_engine(s)
        e.setup(1, 1)
        e.run(40, 0)           # melt a little so displacements are representative
        b0 = e.stats()["nbuilds"]
        e.run(steps, 0)
        st = e.stats()
        ms = e.last_run_ms()
        row = {"skin": skin, "policy": policy, "natoms": n, "steps": steps,
               "builds_per_100_steps": round((st["nbuilds"] - b0) * 100.0 / steps, 2),
               "dangerous": st["ndanger"], "pairs_per_atom": round(st["npairs"] / n, 2),
               "ms_per_step": round(ms / steps, 4), "matom_steps_per_s": round(n * steps / ms / 1e3, 1)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        e.close()
