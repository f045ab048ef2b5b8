"""Unit systems: the constants `units lj` / `units metal` set in the reference
(src/update.cpp:146-200 Update::set_units)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Units:
    name: str
    boltz: float
    mvv2e: float
    ftm2v: float
    nktv2p: float
    dt: float      # default timestep
    skin: float    # default neighbor skin
    normalize: bool  # thermo_modify norm default (thermo.cpp: lj -> yes)


LJ = Units("lj", boltz=1.0, mvv2e=1.0, ftm2v=1.0, nktv2p=1.0, dt=0.005, skin=0.3, normalize=True)
METAL = Units("metal", boltz=8.617343e-5, mvv2e=1.0364269e-4, ftm2v=1.0 / 1.0364269e-4,
              nktv2p=1.6021765e6, dt=0.001, skin=2.0, normalize=False)


def get(name: str) -> Units:
    try:
        return {"lj": LJ, "metal": METAL}[name]
    except KeyError:
        raise ValueError(f"unsupported unit style '{name}' (lj, metal)") from None
