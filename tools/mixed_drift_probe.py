import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from common import lj_system, eam_system, make_engine, make_oracle
for kind in ("lj", "eam"):
    s = lj_system((12, 12, 12)) if kind == "lj" else eam_system((8, 8, 8))
    o = make_oracle(s); o.setup(1, 1)
    e = make_engine(s, "mixed"); e.setup(1, 1)
    to = o.run(100, 0, 50); te = e.run(100, 50)
    for ro, re_ in zip(to, te):
        a, b = e.thermo_row(ro), e.thermo_row(re_)
        print(kind, a["step"], {k: abs(a[k] - b[k]) / max(abs(a[k]), 1e-3) for k in ("temp", "e_pair", "toteng", "press")})
