"""Fixtures of round 2, generated from the UNMODIFIED reference compiled here (oracle/_ref):

    python tests/golden/make_golden_round2.py

  ref_r2_<case>.npz  a melted state handed over after `run 0` (list rebuilt at that state):
      box (lo, hi, xy, xz, yz), x/v/type/tag/mask/image of the owned atoms, the reference's forces,
      potential energy, pxy, and the identity of every stored pair of its neighbour list
      (unordered tag pair + separation vector rounded to 1e-6, sorted: count + sha256 of the int64
      array; the array itself for one case); then x, f, pe after 40 more steps.  Cases: triclinic lj (tag rule), newton off (orthogonal and triclinic), exclude group,
      triclinic eam, newton-off eam.
The live comparison (tests/test_oracle_tri_newton_live.py) needs oracle/_ref; these files do not.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import ref_harness as R  # noqa: E402
import test_oracle_tri_newton_live as T  # noqa: E402

OUT = Path(__file__).resolve().parent
NAMES = ["lj_tri_a", "lj_tri_b", "lj_newtoff", "lj_tri_newtoff", "lj_exclude_group", "eam_tri", "eam_newtoff"]


def main():
    for name, (kind, region, newton, groups, neigh) in zip(NAMES, T.CASES):
        every, delay, check = neigh
        ntext = f"every {every} delay {delay} check {'yes' if check else 'no'}"
        style = "lj/cut" if kind == "lj" else "eam"
        with R.RefLammps() as ref:
            if kind == "lj":
                ref.commands(T.LJ.format(newton=newton, region=region, groups=groups, neigh=ntext))
            else:
                ref.commands(T.EAM.format(newton=newton, region=region, neigh=ntext, pot=R.POTENTIALS))
            ref.command("run 30")
            ref.command("run 0")
            s0 = T.ref_state(ref, kind == "lj")
            pi, pj = ref.neighbor_pairs(style)
            keys = T.pair_keys(pi, pj, s0["tag"], s0["x"])
            ref.command("run 40")
            s1 = T.ref_state(ref, kind == "lj")
        n = s0["n"]
        o1 = np.argsort(s1["tag"][:n])
        np.savez_compressed(
            OUT / f"ref_r2_{name}.npz", kind=kind, triclinic="prism" in region, newton=0 if newton else 1,
            exclude=bool(groups), every=every, delay=delay, check=check, lo=s0["lo"], hi=s0["hi"],
            tilt=np.array([s0["xy"], s0["xz"], s0["yz"]]), x=s0["x"][:n], v=s0["v"], type=s0["type"],
            tag=s0["tag"][:n], mask=s0["mask"], image=s0["image"], f=s0["f"], pe=s0["pe"], pxy=s0["pxy"],
            press=s0["press"], vol=s0["vol"], nghost=len(s0["tag"]) - n, npairs=len(keys),
            pair_hash=hashlib.sha256(np.ascontiguousarray(keys, dtype=np.int64).tobytes()).hexdigest(),
            pair_keys=keys.astype(np.int32) if name == "lj_tri_a" else np.zeros((0, 5), np.int32),
            x40=s1["x"][:n][o1], f40=s1["f"][o1], tag40=s1["tag"][:n][o1], pe40=s1["pe"])
        print(name, n, len(keys))


if __name__ == "__main__":
    main()
