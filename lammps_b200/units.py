"""Unit systems: the constants `units lj` / `units metal` / `units real` set in the reference
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


REAL = Units("real", boltz=0.0019872067, mvv2e=48.88821291 * 48.88821291,
             ftm2v=1.0 / 48.88821291 / 48.88821291, nktv2p=68568.415, dt=1.0, skin=2.0,
             normalize=False)

# utils::get_conversion_factor(ENERGY, METAL2REAL) (utils.cpp): eV -> kcal/mol, applied by the
# potential file readers when a metal-units file is used under `units real`
METAL2REAL_ENERGY = 23.060549


def get(name: str) -> Units:
    try:
        return {"lj": LJ, "metal": METAL, "real": REAL}[name]
    except KeyError:
        raise ValueError(f"unsupported unit style '{name}' (lj, metal, real)") from None
