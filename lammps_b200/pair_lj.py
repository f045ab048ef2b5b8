"""pair_style lj/cut coefficient tables, as PairLJCut::init_one builds them
(src/pair_lj_cut.cpp:503-524; mixing src/pair.cpp:774-802, default geometric).
Tables are (ntypes+1) x (ntypes+1), row/col 0 unused, like the reference's arrays."""
from __future__ import annotations

import math

import numpy as np


def lj_cut_tables(ntypes: int, coeffs: dict, cut_global: float, offset_flag: bool = False,
                  mix: str = "geometric") -> dict:
    """coeffs: {(i, j): (epsilon, sigma[, cut])} with 1-based types, i <= j."""
    n1 = ntypes + 1
    eps = np.zeros((n1, n1))
    sig = np.zeros((n1, n1))
    cut = np.zeros((n1, n1))
    setflag = np.zeros((n1, n1), dtype=bool)
    for (i, j), c in coeffs.items():
        i, j = min(i, j), max(i, j)
        eps[i, j], sig[i, j] = c[0], c[1]
        cut[i, j] = c[2] if len(c) > 2 else cut_global
        setflag[i, j] = True
    t = {k: np.zeros((n1, n1)) for k in ("cutsq", "lj1", "lj2", "lj3", "lj4", "offset")}
    for i in range(1, n1):
        for j in range(i, n1):
            if not setflag[i, j]:
                if not (setflag[i, i] and setflag[j, j]):
                    raise ValueError(f"All pair coeffs are not set ({i},{j})")
                eps[i, j] = math.sqrt(eps[i, i] * eps[j, j])
                if mix == "geometric":
                    sig[i, j] = math.sqrt(sig[i, i] * sig[j, j])
                    cut[i, j] = math.sqrt(cut[i, i] * cut[j, j])
                elif mix == "arithmetic":
                    sig[i, j] = 0.5 * (sig[i, i] + sig[j, j])
                    cut[i, j] = 0.5 * (cut[i, i] + cut[j, j])
                else:
                    raise ValueError("mix must be geometric or arithmetic")
            e, s, c = eps[i, j], sig[i, j], cut[i, j]
            t["lj1"][i, j] = 48.0 * e * math.pow(s, 12.0)
            t["lj2"][i, j] = 24.0 * e * math.pow(s, 6.0)
            t["lj3"][i, j] = 4.0 * e * math.pow(s, 12.0)
            t["lj4"][i, j] = 4.0 * e * math.pow(s, 6.0)
            if offset_flag and c > 0.0:
                ratio = s / c
                t["offset"][i, j] = 4.0 * e * (math.pow(ratio, 12.0) - math.pow(ratio, 6.0))
            t["cutsq"][i, j] = c * c  # pair.cpp Pair::init: cutsq = init_one()^2
            for k in t:
                t[k][j, i] = t[k][i, j]
    out = {k: np.ascontiguousarray(v) for k, v in t.items()}
    out["ntypes"] = ntypes
    out["special_lj"] = np.ones(4)
    return out
