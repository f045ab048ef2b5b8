"""Synthetic inputs of the reference benchmarks, built on the host exactly as the reference's
one-time setup commands build them, so that a run can be compared digit-for-digit with the
reference's golden logs:

  lattice fcc / region block / create_box / create_atoms
      src/lattice.cpp:296-305 (lj scale), :551-568 (lattice2box),
      src/create_atoms.cpp:1379-1460 (loop order k,j,i,basis; tags in creation order)
  velocity all create T SEED loop geom   (dist uniform, mom yes, rot no)
      src/velocity.cpp:158-400, src/random_park.cpp:41-48,96-130,
      src/group.cpp:873-896 (mass), :1174-1210 (vcm), src/compute_temp.cpp:57-97

These are host-side, one-time O(N) numpy passes (the reference runs them on the host too);
nothing here is on the per-timestep path.
"""
from __future__ import annotations

import numpy as np

from . import units as _units

FCC_BASIS = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]])


def lattice_scale(unit_style: str, value: float, nbasis: int = 4) -> float:
    """lattice.cpp:296-305: in lj units the argument is the reduced density rho*."""
    if unit_style == "lj":
        volume = 1.0  # cubic primitive cell a1.(a2 x a3)
        return float(pow(nbasis / volume / value, 1.0 / 3.0))
    return float(value)


def fcc_block(unit_style: str, value: float, ncell):
    """Atoms of `create_atoms 1 box` in `region block 0 nx 0 ny 0 nz` (lattice units).

    Returns (x[N,3], boxlo[3], boxhi[3]); atom k has tag k+1 (serial creation order).
    """
    nx, ny, nz = (int(c) for c in ncell)
    scale = lattice_scale(unit_style, value)
    boxlo = np.zeros(3)
    boxhi = np.array([scale * nx, scale * ny, scale * nz])  # region_block.cpp: xscale*hi
    # loop order k (z) outermost, then j, i, then basis m  (create_atoms.cpp:1389-1392)
    k, j, i, m = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), np.arange(4),
                             indexing="ij")
    lat = np.stack([i.ravel() + FCC_BASIS[m.ravel(), 0],
                    j.ravel() + FCC_BASIS[m.ravel(), 1],
                    k.ravel() + FCC_BASIS[m.ravel(), 2]], axis=1)
    x = lat * scale  # lattice2box: primitive = rotaterow = identity, origin 0
    return np.ascontiguousarray(x), boxlo, boxhi


_IA, _IM, _IQ, _IR = 16807, 2147483647, 127773, 2836
_AM = 1.0 / _IM


def _park_uniform(seed: np.ndarray):
    k = seed // _IQ
    seed = _IA * (seed - k * _IQ) - _IR * k
    seed = np.where(seed < 0, seed + _IM, seed)
    return seed, _AM * seed


def _geom_seeds(x: np.ndarray, seed: int) -> np.ndarray:
    """RanPark::reset(ibase, coord): Jenkins one-at-a-time hash over the bytes of the seed
    and of the three coordinates; bytes are *signed* chars on the reference platform."""
    n = x.shape[0]
    h = np.zeros(n, dtype=np.uint32)

    def mix(h, byte_u32):
        h = h + byte_u32
        h = h + (h << np.uint32(10))
        h = h ^ (h >> np.uint32(6))
        return h

    with np.errstate(over="ignore"):
        for b in np.array([seed], dtype=np.int32).view(np.int8):
            h = mix(h, np.uint32(np.int64(b) & 0xFFFFFFFF))
        xb = np.ascontiguousarray(x, dtype=np.float64).view(np.int8).reshape(n, 24)
        for c in range(24):
            h = mix(h, xb[:, c].astype(np.int32).astype(np.uint32))
        h = h + (h << np.uint32(3))
        h = h ^ (h >> np.uint32(11))
        h = h + (h << np.uint32(15))
    s = (h & np.uint32(0x7FFFFFF)).astype(np.int64)
    s[s == 0] = 1
    for _ in range(5):  # warm-up draws
        s, _u = _park_uniform(s)
    return s


def velocity_create(x, type, mass, t_target: float, seed: int, unit_style: str):
    """`velocity all create T seed loop geom` with the defaults dist=uniform mom=yes rot=no."""
    u = _units.get(unit_style)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    m_atom = np.asarray(mass, dtype=np.float64)[np.asarray(type)]
    s = _geom_seeds(x, seed)
    v = np.empty((n, 3))
    for c in range(3):
        s, uni = _park_uniform(s)
        v[:, c] = (uni - 0.5) * (1.0 / np.sqrt(m_atom))
    # zero_momentum / temperature: the reference accumulates sequentially in atom order;
    # np.cumsum does the same, which keeps the velocities bit-identical to the reference's
    def seqsum(a):
        return float(np.cumsum(a)[-1])
    masstotal = seqsum(m_atom)
    vcm = np.array([seqsum(v[:, c] * m_atom) for c in range(3)]) / masstotal
    v -= vcm
    # scale to the target temperature: compute temp with dof = 3N - 3
    dof = 3.0 * n - 3.0
    tfactor = u.mvv2e / (dof * u.boltz)
    t = seqsum((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]) * m_atom) * tfactor
    v *= np.sqrt(t_target / t)
    return v
