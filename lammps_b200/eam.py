"""pair_style eam host-side setup: DYNAMO funcfl file -> common-grid arrays -> 7-coefficient
splines, following the reference operation by operation so the tables are bit-identical:

  read_funcfl         src/MANYBODY/pair_eam.cpp:663-732
  file2array_funcfl   src/MANYBODY/pair_eam.cpp:999-1205   (4-point Lagrange resampling)
  array2spline        src/MANYBODY/pair_eam.cpp:1492-1512
  interpolate         src/MANYBODY/pair_eam.cpp:1516-1545  (row m is 1-based, m in [1,n])
  init_one            src/MANYBODY/pair_eam.cpp:624-643    (cutforcesq = cutmax^2, scale = 1)

The finished tables are uploaded once through b200_pair_eam(); nothing here runs per step.
(In the LAMMPS-hosted build the C++ style inherits these tables from PairEAM instead.)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Funcfl:
    mass: float
    nrho: int
    drho: float
    nr: int
    dr: float
    cut: float
    frho: np.ndarray  # 1-based: index 0 unused
    zr: np.ndarray
    rhor: np.ndarray


def read_funcfl(path: str) -> Funcfl:
    with open(path) as fh:
        fh.readline()  # comment line
        w = fh.readline().split()
        mass = float(w[1])
        w = fh.readline().split()
        nrho, drho, nr, dr, cut = int(w[0]), float(w[1]), int(w[2]), float(w[3]), float(w[4])
        vals = []
        need = nrho + 2 * nr
        for line in fh:
            vals.extend(float(t) for t in line.split())
            if len(vals) >= need:
                break
    if nrho <= 0 or nr <= 0 or dr <= 0.0 or len(vals) < need:
        raise ValueError("Invalid EAM potential file")
    vals = np.array(vals[:need])
    one = lambda a: np.concatenate([[0.0], a])  # noqa: E731  (shift to 1-based)
    return Funcfl(mass, nrho, drho, nr, dr, cut, one(vals[:nrho]), one(vals[nrho:nrho + nr]),
                  one(vals[nrho + nr:]))


def _lagrange(table: np.ndarray, ntab: int, dtab: float, r: np.ndarray) -> np.ndarray:
    """The resampling kernel shared by frho / rhor / zr in file2array_funcfl."""
    sixth = 1.0 / 6.0
    p = r / dtab + 1.0
    k = p.astype(np.int64)  # static_cast<int>, p >= 1 so truncation == floor
    k = np.minimum(k, ntab - 2)
    k = np.maximum(k, 2)
    p = p - k
    p = np.minimum(p, 2.0)
    cof1 = -sixth * p * (p - 1.0) * (p - 2.0)
    cof2 = 0.5 * (p * p - 1.0) * (p - 2.0)
    cof3 = -0.5 * p * (p + 1.0) * (p - 2.0)
    cof4 = sixth * p * (p * p - 1.0)
    return cof1 * table[k - 1] + cof2 * table[k] + cof3 * table[k + 1] + cof4 * table[k + 2]


def _interpolate(n: int, delta: float, f: np.ndarray) -> np.ndarray:
    """PairEAM::interpolate: f is 1-based [n+1]; returns spline[n+1][7]."""
    s = np.zeros((n + 1, 7))
    s[1:, 6] = f[1:n + 1]
    s[1, 5] = s[2, 6] - s[1, 6]
    s[2, 5] = 0.5 * (s[3, 6] - s[1, 6])
    s[n - 1, 5] = 0.5 * (s[n, 6] - s[n - 2, 6])
    s[n, 5] = s[n, 6] - s[n - 1, 6]
    m = np.arange(3, n - 1)
    s[m, 5] = ((s[m - 2, 6] - s[m + 2, 6]) + 8.0 * (s[m + 1, 6] - s[m - 1, 6])) / 12.0
    m = np.arange(1, n)
    s[m, 4] = 3.0 * (s[m + 1, 6] - s[m, 6]) - 2.0 * s[m, 5] - s[m + 1, 5]
    s[m, 3] = s[m, 5] + s[m + 1, 5] - 2.0 * (s[m + 1, 6] - s[m, 6])
    s[n, 4] = 0.0
    s[n, 3] = 0.0
    m = np.arange(1, n + 1)
    s[m, 2] = s[m, 5] / delta
    s[m, 1] = 2.0 * s[m, 4] / delta
    s[m, 0] = 3.0 * s[m, 3] / delta
    return s


@dataclass
class EAMTables:
    ntypes: int
    nr: int
    nrho: int
    dr: float
    drho: float
    rdr: float
    rdrho: float
    rhomax: float
    cutmax: float
    cutforcesq: float
    type2frho: np.ndarray   # [ntypes+1]
    type2rhor: np.ndarray   # [ntypes+1, ntypes+1]
    type2z2r: np.ndarray    # [ntypes+1, ntypes+1]
    scale: np.ndarray       # [ntypes+1, ntypes+1]
    frho_spline: np.ndarray  # [nfrho, nrho+1, 7]
    rhor_spline: np.ndarray  # [nrhor, nr+1, 7]
    z2r_spline: np.ndarray   # [nz2r, nr+1, 7]
    mass: np.ndarray         # [ntypes+1] per-type mass from the files (pair_eam.cpp coeff)

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k in (
            "ntypes", "nr", "nrho", "rdr", "rdrho", "rhomax", "cutforcesq", "type2frho",
            "type2rhor", "type2z2r", "scale", "frho_spline", "rhor_spline", "z2r_spline")}
        d["nfrho"] = self.frho_spline.shape[0]
        d["nrhor"] = self.rhor_spline.shape[0]
        d["nz2r"] = self.z2r_spline.shape[0]
        return d


def funcfl_tables(files: list[Funcfl], type_map: list[int], energy_scale: float = 1.0) -> EAMTables:
    """`pair_coeff i i file` for each type: type_map[t] (t = 1..ntypes) = index into files.

    file2array_funcfl + array2spline + init_one.  energy_scale != 1: the reader's transparent
    unit conversion (read_funcfl, pair_eam.cpp:701-706): F(rho) * c, Z(r) * sqrt(c).
    """
    if energy_scale != 1.0:
        files = [Funcfl(f.mass, f.nrho, f.drho, f.nr, f.dr, f.cut, f.frho * energy_scale,
                        f.zr * np.sqrt(energy_scale), f.rhor) for f in files]
    ntypes = len(type_map)
    map_ = [-1] + list(type_map)
    nfuncfl = len(files)
    active = [f for i, f in enumerate(files) if i in type_map]
    dr = max(f.dr for f in active)
    drho = max(f.drho for f in active)
    rmax = max((f.nr - 1) * f.dr for f in active)
    rhomax = max((f.nrho - 1) * f.drho for f in active)
    nr = int(np.floor(rmax / dr + 0.5))      # std::lround for positive values
    nrho = int(np.floor(rhomax / drho + 0.5))

    # frho: one array per file + one of zeros
    nfrho = nfuncfl + 1
    frho = np.zeros((nfrho, nrho + 1))
    r = (np.arange(1, nrho + 1) - 1) * drho
    for n, f in enumerate(files):
        frho[n, 1:] = _lagrange(f.frho, f.nrho, f.drho, r)
    type2frho = np.array([0] + [map_[i] if map_[i] >= 0 else nfrho - 1
                                for i in range(1, ntypes + 1)], dtype=np.int32)

    # rhor: one array per file
    rhor = np.zeros((nfuncfl, nr + 1))
    r = (np.arange(1, nr + 1) - 1) * dr
    for n, f in enumerate(files):
        rhor[n, 1:] = _lagrange(f.rhor, f.nr, f.dr, r)
    type2rhor = np.zeros((ntypes + 1, ntypes + 1), dtype=np.int32)
    for i in range(1, ntypes + 1):
        type2rhor[i, 1:] = map_[i]

    # z2r: lower-triangular file pairs, with the eV/Hartree * Ang/Bohr unit conversion
    nz2r = nfuncfl * (nfuncfl + 1) // 2
    z2r = np.zeros((nz2r, nr + 1))
    n = 0
    for i, fi in enumerate(files):
        for j in range(i + 1):
            fj = files[j]
            zri = _lagrange(fi.zr, fi.nr, fi.dr, r)
            zrj = _lagrange(fj.zr, fj.nr, fj.dr, r)
            z2r[n, 1:] = 27.2 * 0.529 * zri * zrj
            n += 1
    type2z2r = np.zeros((ntypes + 1, ntypes + 1), dtype=np.int32)
    for i in range(1, ntypes + 1):
        for j in range(1, ntypes + 1):
            irow, icol = map_[i], map_[j]
            if irow == -1 or icol == -1:
                continue
            if irow < icol:
                irow, icol = icol, irow
            type2z2r[i, j] = sum(m + 1 for m in range(irow)) + icol

    frho_s = np.stack([_interpolate(nrho, drho, frho[i]) for i in range(nfrho)])
    rhor_s = np.stack([_interpolate(nr, dr, rhor[i]) for i in range(nfuncfl)])
    z2r_s = np.stack([_interpolate(nr, dr, z2r[i]) for i in range(nz2r)])
    cutmax = max(f.cut for f in files)
    mass = np.array([0.0] + [files[map_[i]].mass for i in range(1, ntypes + 1)])
    return EAMTables(ntypes=ntypes, nr=nr, nrho=nrho, dr=dr, drho=drho, rdr=1.0 / dr,
                     rdrho=1.0 / drho, rhomax=rhomax, cutmax=cutmax, cutforcesq=cutmax * cutmax,
                     type2frho=type2frho, type2rhor=type2rhor, type2z2r=type2z2r,
                     scale=np.ones((ntypes + 1, ntypes + 1)),
                     frho_spline=np.ascontiguousarray(frho_s),
                     rhor_spline=np.ascontiguousarray(rhor_s),
                     z2r_spline=np.ascontiguousarray(z2r_s), mass=mass)


# ----------------------------------------------------------------------------- setfl (eam/alloy)
@dataclass
class Setfl:
    """Numeric content of a DYNAMO setfl file (pair_eam.cpp:738-805 PairEAM::read_setfl).
    Arrays are 0-based here: frho[e][k], rhor[e][k], z2r[i][j][k] for i >= j."""
    elements: list
    mass: np.ndarray
    nrho: int
    drho: float
    nr: int
    dr: float
    cut: float
    frho: np.ndarray   # [nelements, nrho]
    rhor: np.ndarray   # [nelements, nr]
    z2r: dict          # (i, j) with i >= j -> [nr]


def read_setfl(path: str, fs: bool = False) -> Setfl:
    """3 comment lines; `nelements names...`; `nrho drho nr dr cut`; per element a header
    (Z mass ...) + nrho F values + nr rho values; then r*phi for every pair i >= j.
    fs=True: the Finnis-Sinclair variant (PairEAM::read_fs, pair_eam.cpp:856-960): every element
    carries one density function per partner element, rhor[i][j] (nelements x nr values)."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    t = lines[3].split()
    nel = int(t[0])
    names = t[1:1 + nel]
    if len(names) != nel:
        raise ValueError("Incorrect element names in EAM potential file")
    t = lines[4].split()
    nrho, drho, nr, dr, cut = int(t[0]), float(t[1]), int(t[2]), float(t[3]), float(t[4])
    if nrho <= 0 or nr <= 0 or dr <= 0.0:
        raise ValueError("Invalid EAM potential file")
    pos = 5
    toks: list = []

    def header():
        # a header line holds 4 tokens, the last one a word: it cannot be taken for data
        nonlocal pos
        h = lines[pos].split()
        pos += 1
        return h

    def take(n):
        nonlocal pos, toks
        while len(toks) < n:
            toks.extend(lines[pos].replace("D", "E").replace("d", "e").split())
            pos += 1
        out, toks = toks[:n], toks[n:]
        return np.array([float(v) for v in out])

    mass = np.zeros(nel)
    frho = np.zeros((nel, nrho))
    rhor = np.zeros((nel, nel, nr)) if fs else np.zeros((nel, nr))
    for e in range(nel):
        assert not toks, "setfl data blocks must end at a line break"
        mass[e] = float(header()[1])
        frho[e] = take(nrho)
        if fs:
            for j in range(nel):
                rhor[e, j] = take(nr)
        else:
            rhor[e] = take(nr)
    z2r = {}
    for i in range(nel):
        for j in range(i + 1):
            z2r[(i, j)] = take(nr)
    return Setfl(names, mass, nrho, drho, nr, dr, cut, frho, rhor, z2r)


def setfl_tables(f: Setfl, type_elements: list, energy_scale: float = 1.0) -> EAMTables:
    """`pair_coeff * * file E1 E2 ...`: type_elements[t-1] = element name of atom type t.
    PairEAM::coeff (map), file2array_setfl (pair_eam.cpp:1211-1325), array2spline.
    energy_scale != 1: the reader's unit conversion (read_setfl/read_fs: F(rho) * c, r*phi * c)."""
    if energy_scale != 1.0:
        f = Setfl(f.elements, f.mass, f.nrho, f.drho, f.nr, f.dr, f.cut, f.frho * energy_scale,
                  f.rhor, {k: v * energy_scale for k, v in f.z2r.items()})
    ntypes = len(type_elements)
    map_ = [-1] + [f.elements.index(e) if e != "NULL" else -1 for e in type_elements]
    nel = len(f.elements)
    nr, nrho, dr, drho = f.nr, f.nrho, f.dr, f.drho
    nfrho = nel + 1                                     # + one array of zeros
    frho = np.zeros((nfrho, nrho + 1))
    frho[:nel, 1:] = f.frho
    type2frho = np.array([0] + [map_[i] if map_[i] >= 0 else nfrho - 1
                                for i in range(1, ntypes + 1)], dtype=np.int32)
    type2rhor = np.zeros((ntypes + 1, ntypes + 1), dtype=np.int32)
    if f.rhor.ndim == 3:                                # eam/fs: file2array_fs, pair_eam.cpp:1331-1456
        rhor = np.zeros((nel * nel, nr + 1))
        rhor[:, 1:] = f.rhor.reshape(nel * nel, nr)
        for i in range(1, ntypes + 1):
            for j in range(1, ntypes + 1):
                type2rhor[i, j] = map_[i] * nel + map_[j]
    else:
        rhor = np.zeros((nel, nr + 1))
        rhor[:, 1:] = f.rhor
        for i in range(1, ntypes + 1):
            type2rhor[i, 1:] = map_[i]                  # the density depends on the source element
    nz2r = nel * (nel + 1) // 2
    z2r = np.zeros((nz2r, nr + 1))
    n = 0
    for i in range(nel):
        for j in range(i + 1):
            z2r[n, 1:] = f.z2r[(i, j)]
            n += 1
    type2z2r = np.zeros((ntypes + 1, ntypes + 1), dtype=np.int32)
    for i in range(1, ntypes + 1):
        for j in range(1, ntypes + 1):
            irow, icol = map_[i], map_[j]
            if irow == -1 or icol == -1:
                continue
            if irow < icol:
                irow, icol = icol, irow
            type2z2r[i, j] = sum(m + 1 for m in range(irow)) + icol
    frho_s = np.stack([_interpolate(nrho, drho, frho[i]) for i in range(nfrho)])
    rhor_s = np.stack([_interpolate(nr, dr, rhor[i]) for i in range(rhor.shape[0])])
    z2r_s = np.stack([_interpolate(nr, dr, z2r[i]) for i in range(nz2r)])
    mass = np.array([0.0] + [f.mass[map_[i]] if map_[i] >= 0 else 0.0 for i in range(1, ntypes + 1)])
    return EAMTables(ntypes=ntypes, nr=nr, nrho=nrho, dr=dr, drho=drho, rdr=1.0 / dr,
                     rdrho=1.0 / drho, rhomax=(nrho - 1) * drho, cutmax=f.cut,
                     cutforcesq=f.cut * f.cut, type2frho=type2frho, type2rhor=type2rhor,
                     type2z2r=type2z2r, scale=np.ones((ntypes + 1, ntypes + 1)),
                     frho_spline=np.ascontiguousarray(frho_s),
                     rhor_spline=np.ascontiguousarray(rhor_s),
                     z2r_spline=np.ascontiguousarray(z2r_s), mass=mass)
