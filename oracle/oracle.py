"""ctypes wrapper around oracle/md_oracle.c (the CPU restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg -- never by the product package lammps_b200.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "md_oracle.c"
LIB = HERE / "libmd_oracle.so"


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o",
                               str(LIB), str(SRC), "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.orc_create.restype = C.c_void_p
        for name in ("orc_nlocal", "orc_nghost", "orc_ncalls", "orc_ndanger", "orc_ago",
                     "orc_decide", "orc_step", "orc_run"):
            getattr(L, name).restype = C.c_int
        L.orc_nneigh.restype = C.c_int64
        L.orc_eng_vdwl.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One periodic box (orthogonal or triclinic) on one process, driven like the reference's Verlet."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.ntypes = 1

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- configuration (mirrors the C-ABI of the product, include/b200_md.h)
    def set_box(self, lo, hi, periodic=(1, 1, 1)):
        lo, hi, per = _d(lo), _d(hi), _i(periodic)
        self.L.orc_set_box(self.h, _p(lo), _p(hi), _p(per))

    def set_box_triclinic(self, lo, hi, xy, xz, yz, periodic=(1, 1, 1), angstrom=1.0):
        lo, hi, per = _d(lo), _d(hi), _i(periodic)
        self.L.orc_set_box_triclinic(self.h, _p(lo), _p(hi), C.c_double(xy), C.c_double(xz), C.c_double(yz),
                                     _p(per), C.c_double(angstrom))

    def set_newton(self, newton_pair):
        self.L.orc_set_newton(self.h, C.c_int(1 if newton_pair else 0))

    def neigh_modify_groups(self, pairs=()):
        b1, b2 = _i([p[0] for p in pairs]), _i([p[1] for p in pairs])
        self.L.orc_neigh_modify_groups(self.h, C.c_int(len(pairs)), _p(b1), _p(b2))

    def set_atoms(self, x, v, type, tag, mass, mask=None, image=None):
        x, v, type, tag, mass = _d(x), _d(v), _i(type), _i(tag), _d(mass)
        n = x.shape[0]
        self.ntypes = mass.shape[0] - 1
        mk = _i(mask) if mask is not None else None
        im = _i(image) if image is not None else None
        self.L.orc_set_atoms(self.h, C.c_int(n), C.c_int(self.ntypes), _p(mass), _p(x), _p(v),
                             _p(type), _p(tag), _p(mk) if mk is not None else None,
                             _p(im) if im is not None else None)

    def set_neighbor(self, skin, every=1, delay=0, check=True):
        self.L.orc_set_neighbor(self.h, C.c_double(skin), C.c_int(every), C.c_int(delay),
                                C.c_int(1 if check else 0))

    def neigh_modify(self, once=False, exclude_types=(), ntypes=1):
        """neigh_modify once yes|no, exclude type i j (pairs of types, symmetric)"""
        ex = None
        if exclude_types:
            n1 = ntypes + 1
            ex = np.zeros(n1 * n1, np.int32)
            for i, j in exclude_types:
                ex[i * n1 + j] = ex[j * n1 + i] = 1
        self.L.orc_neigh_modify(self.h, C.c_int(1 if once else 0), C.c_int(ntypes),
                                _p(ex) if ex is not None else None)

    def fix_nve(self, dt, ftm2v=1.0, groupbit=1):
        self.L.orc_fix_nve(self.h, C.c_double(dt), C.c_double(ftm2v), C.c_int(groupbit))

    def pair_lj_cut(self, p):
        """p: dict of (ntypes+1)^2 tables cutsq, lj1..lj4, offset (see lammps_b200.pair_lj)."""
        t = {k: _d(p[k]) for k in ("cutsq", "lj1", "lj2", "lj3", "lj4", "offset")}
        sp = _d(p.get("special_lj", [1.0, 1.0, 1.0, 1.0]))
        self.L.orc_pair_lj_cut(self.h, C.c_int(p["ntypes"]), _p(t["cutsq"]), _p(t["lj1"]),
                               _p(t["lj2"]), _p(t["lj3"]), _p(t["lj4"]), _p(t["offset"]), _p(sp))

    def pair_eam(self, t):
        """t: dict from lammps_b200.eam.EAMTables.as_dict()."""
        a = {k: _i(t[k]) for k in ("type2frho", "type2rhor", "type2z2r")}
        b = {k: _d(t[k]) for k in ("scale", "frho_spline", "rhor_spline", "z2r_spline")}
        self.L.orc_pair_eam(self.h, C.c_int(t["ntypes"]), C.c_int(t["nr"]), C.c_int(t["nrho"]),
                            C.c_double(t["rdr"]), C.c_double(t["rdrho"]), C.c_double(t["rhomax"]),
                            C.c_double(t["cutforcesq"]), _p(a["type2frho"]), _p(a["type2rhor"]),
                            _p(a["type2z2r"]), _p(b["scale"]), C.c_int(t["nfrho"]),
                            _p(b["frho_spline"]), C.c_int(t["nrhor"]), _p(b["rhor_spline"]),
                            C.c_int(t["nz2r"]), _p(b["z2r_spline"]))

    # ---- driver
    def setup(self, eflag=1, vflag=1):
        self.L.orc_setup(self.h, C.c_int(eflag), C.c_int(vflag))

    def step(self, eflag=0, vflag=0) -> int:
        return self.L.orc_step(self.h, C.c_int(eflag), C.c_int(vflag))

    def run(self, nsteps, step0=0, thermo_every=0):
        out = np.zeros((nsteps + 1, 10))
        n = self.L.orc_run(self.h, C.c_int(nsteps), C.c_int(step0), C.c_int(thermo_every),
                           _p(out), C.c_int(out.shape[0]))
        return out[:n]

    def pair_compute(self, eflag=1, vflag=1):
        self.L.orc_force_clear(self.h)
        self.L.orc_pair_compute(self.h, C.c_int(eflag), C.c_int(vflag))

    def reverse_comm(self):
        self.L.orc_reverse_comm(self.h)

    # ---- results
    @property
    def nlocal(self):
        return self.L.orc_nlocal(self.h)

    @property
    def nghost(self):
        return self.L.orc_nghost(self.h)

    @property
    def nneigh(self):
        return self.L.orc_nneigh(self.h)

    @property
    def ncalls(self):
        return self.L.orc_ncalls(self.h)

    @property
    def ndanger(self):
        return self.L.orc_ndanger(self.h)

    @property
    def eng_vdwl(self):
        return self.L.orc_eng_vdwl(self.h)

    @property
    def virial(self):
        v = np.zeros(6)
        self.L.orc_get_virial(self.h, _p(v))
        return v

    def ke_sum(self):
        out = C.c_double()
        self.L.orc_ke_sum(self.h, C.byref(out))
        return out.value

    def _vec(self, which, ghosts=False):
        n = self.nlocal + (self.nghost if ghosts else 0)
        out = np.zeros((n, 3))
        self.L.orc_get_vec(self.h, C.c_int(which), C.c_int(n), _p(out))
        return out

    def _ivec(self, which, ghosts=False):
        n = self.nlocal + (self.nghost if ghosts else 0)
        out = np.zeros(n, np.int32)
        self.L.orc_get_ivec(self.h, C.c_int(which), C.c_int(n), _p(out))
        return out

    def x(self, ghosts=False):
        return self._vec(0, ghosts)

    def v(self):
        return self._vec(1)

    def f(self, ghosts=False):
        return self._vec(2, ghosts)

    def type(self, ghosts=False):
        return self._ivec(0, ghosts)

    def tag(self, ghosts=False):
        return self._ivec(1, ghosts)

    def image(self):
        return self._ivec(3)

    def numneigh(self):
        return self._ivec(4)

    def bins(self):
        out = np.zeros(10, np.int32)
        self.L.orc_get_bins(self.h, _p(out))
        keys = ("nbinx", "nbiny", "nbinz", "mbinx", "mbiny", "mbinz", "mbinxlo", "mbinylo",
                "mbinzlo", "nstencil")
        return dict(zip(keys, out.tolist()))

    def stencil(self):
        out = np.zeros(self.bins()["nstencil"], np.int32)
        self.L.orc_get_stencil(self.h, _p(out))
        return out

    def rho_fp(self, ghosts=False):
        n = self.nlocal + (self.nghost if ghosts else 0)
        rho, fp = np.zeros(n), np.zeros(n)
        self.L.orc_get_rho_fp(self.h, C.c_int(n), _p(rho), _p(fp))
        return rho, fp

    def pair_peratom(self):
        """Pair::ev_tally's eatom[nlocal], vatom[nlocal][6] after the reverse communication of the
        per-atom computes (orc_pair_peratom); call after a compute that tallied."""
        e, v = np.zeros(self.nlocal), np.zeros((self.nlocal, 6))
        self.L.orc_pair_peratom.restype = None
        self.L.orc_pair_peratom(self.h, _p(e), _p(v))
        return e, v

    def pairs(self):
        n = self.nneigh
        pi, pj = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.orc_get_pairs(self.h, _p(pi), _p(pj))
        return pi, pj


def velocity_loop_geom(x, seed, mass_per_atom):
    """velocity.cpp:327-352 (`loop geom`, uniform distribution) -- raw velocities before
    momentum zeroing and temperature scaling."""
    x = _d(x)
    m = _d(mass_per_atom)
    v = np.zeros_like(x)
    lib().orc_velocity_loop_geom(C.c_int(x.shape[0]), C.c_int(seed), _p(x), _p(m), _p(v))
    return v


def canonical_pairs_box(pi, pj, tag, x, boxlo, boxhi, nlocal=None):
    """Order-independent identity of a neighbour list: for every stored pair (i, j) the key
    (tag_a, tag_b, sx, sy, sz) with tag_a <= tag_b and (sx,sy,sz) = the lattice vector (in
    box lengths) separating the stored image of b from the owned image of a.  Sorted
    lexicographically -> a multiset that two implementations must reproduce exactly,
    whatever their atom order and whichever of the two mirror (owned, ghost) images each
    happened to store.  `tag`, `x` cover owned+ghost atoms (owned first, `nlocal` of them).
    """
    pi = np.asarray(pi, np.int64)
    pj = np.asarray(pj, np.int64)
    tag = np.asarray(tag, np.int64)
    boxlo = np.asarray(boxlo, float)
    prd = np.asarray(boxhi, float) - boxlo
    if nlocal is None:
        nlocal = int(pi.max()) + 1 if pi.size else 0
    owner = np.full(int(tag.max()) + 1, -1, np.int64)
    owner[tag[:nlocal]] = np.arange(nlocal)
    # image shift of every atom relative to its owned copy (0 for owned atoms)
    shift = np.rint((x - x[owner[tag]]) / prd).astype(np.int64)
    s = shift[pj] - shift[pi]
    ta, tb = tag[pi], tag[pj]
    swap = (ta > tb)
    a = np.where(swap, tb, ta)
    b = np.where(swap, ta, tb)
    s = np.where(swap[:, None], -s, s)
    # self-image pairs (ta == tb): canonical sign = first non-zero component positive
    same = (ta == tb)
    if same.any():
        sgn = np.sign(s[:, 0] * 9 + s[:, 1] * 3 + s[:, 2])
        flip = same & (sgn < 0)
        s = np.where(flip[:, None], -s, s)
    key = np.stack([a, b, s[:, 0], s[:, 1], s[:, 2]], axis=1)
    order = np.lexsort(key.T[::-1])
    return key[order]


# ---- fix langevin (test infrastructure like everything in this file) --------------------------
def langevin_prefactors(mass, t_period, dt, boltz, ftm2v, mvv2e, ratio=None):
    """FixLangevin::init, fix_langevin.cpp:268-280: gfactor1[t], gfactor2[t] (index 0 unused)."""
    mass = np.asarray(mass, np.float64)
    g1 = np.zeros_like(mass)
    g2 = np.zeros_like(mass)
    for t in range(1, len(mass)):
        r = 1.0 if ratio is None else ratio[t]
        g1[t] = -mass[t] / t_period / ftm2v
        g2[t] = np.sqrt(mass[t]) / ftm2v
        g2[t] *= np.sqrt(24.0 * boltz / t_period / dt / mvv2e)
        g1[t] *= 1.0 / r
        g2[t] *= 1.0 / np.sqrt(r)
    return g1, g2


def langevin_post_force(f, v, type, g1, g2, tsqrt, uniforms, zero=False):
    """FixLangevin::post_force_templated<0,0,0,0,ZERO>, fix_langevin.cpp:424-497, group all:
    f += gamma1 v + gamma2 (u - 0.5), each operation rounded separately (numpy float64 does);
    `uniforms` [n,3] are the three draws of each atom.  Returns the new forces."""
    gamma1 = g1[type][:, None]
    gamma2 = (g2[type] * tsqrt)[:, None]
    fran = gamma2 * (uniforms - 0.5)
    fdrag = gamma1 * v
    out = f + (fdrag + fran)
    if zero:
        fsum = np.zeros(3)
        for i in range(len(f)):           # the reference's sequential sum
            fsum += fran[i]
        out = out - fsum / len(f)
    return out


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 (Salmon et al., SC'11) on numpy uint32 arrays: the counter-based stream of
    the product's fix langevin kernel (kernels_step.cuh), restated to check it word for word."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c0, c1, c2, c3 = (np.asarray(a, np.uint32).copy() for a in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
        k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def langevin_device_uniforms(tag, seed, step):
    """u[n,3] of the device stream: key = seed (lo, hi), counter = (tag, step lo, step hi, 0)."""
    tag = np.asarray(tag, np.uint32)
    z = np.zeros_like(tag)
    r = philox4x32_10(tag, z + np.uint32(step & 0xFFFFFFFF), z + np.uint32((step >> 32) & 0xFFFFFFFF), z,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([(w.astype(np.float64) + 0.5) * 2.0 ** -32 for w in r[:3]], axis=1)


class RanMars:
    """RanMars (random_mars.cpp:32-100): the Marsaglia generator fix langevin draws from, restated
    so that a test can rebuild the reference's uniform stream (seed + comm->me, one draw consumed
    by the constructor)."""

    def __init__(self, seed):
        assert 0 < seed <= 900000000
        ij = (seed - 1) // 30082
        kl = (seed - 1) - 30082 * ij
        i = (ij // 177) % 177 + 2
        j = ij % 177 + 2
        k = (kl // 169) % 178 + 1
        l = kl % 169
        self.u = [0.0] * 98
        for ii in range(1, 98):
            s, t = 0.0, 0.5
            for _ in range(24):
                m = ((i * j) % 179) * k % 179
                i, j, k = j, k, m
                l = (53 * l + 1) % 169
                if (l * m) % 64 >= 32:
                    s = s + t
                t = 0.5 * t
            self.u[ii] = s
        self.c = 362436.0 / 16777216.0
        self.cd = 7654321.0 / 16777216.0
        self.cm = 16777213.0 / 16777216.0
        self.i97, self.j97 = 97, 33
        self.uniform()

    def uniform(self):
        uni = self.u[self.i97] - self.u[self.j97]
        if uni < 0.0:
            uni += 1.0
        self.u[self.i97] = uni
        self.i97 -= 1
        if self.i97 == 0:
            self.i97 = 97
        self.j97 -= 1
        if self.j97 == 0:
            self.j97 = 97
        self.c -= self.cd
        if self.c < 0.0:
            self.c += self.cm
        uni -= self.c
        if uni < 0.0:
            uni += 1.0
        return uni

    def uniforms(self, n):
        return np.array([self.uniform() for _ in range(n)])
