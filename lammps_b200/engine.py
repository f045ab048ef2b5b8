"""Host mirror of the C ABI in include/b200_md.h (ctypes; no torch types cross the boundary).

`Engine` exposes the same vocabulary a LAMMPS input uses (units, pair_style, neighbor,
neigh_modify, fix nve, run, thermo) so tests read like the reference's own inputs, and maps
one-to-one onto b200_* entry points.  There is NO CPU fallback: if libb200md.so is missing or
no CUDA device is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import units as _units

HERE = Path(__file__).resolve().parent
LIBPATH = HERE / "libb200md.so"

NPHASE = 9
PHASES = ("initial_integrate", "final_integrate", "forward_comm", "reverse_comm", "pair",
          "neigh", "neigh_build", "force_clear", "thermo")

EXPORTS = [
    "b200_create", "b200_destroy", "b200_last_error", "b200_device_count", "b200_set_box",
    "b200_set_decomposition", "b200_set_rank_grid", "b200_set_neighbor", "b200_neigh_modify", "b200_set_atoms",
    "b200_get_atoms",
    "b200_get_counts", "b200_pair_lj_cut", "b200_pair_eam", "b200_fix_nve", "b200_setup",
    "b200_run", "b200_step", "b200_step_ahead", "b200_group_step_ahead", "b200_group_decide", "b200_group_reneighbor", "b200_group_forward_comm",
    "b200_group_force_clear", "b200_group_pair_compute", "b200_group_reverse_comm", "b200_group_nve_v",
    "b200_group_nve_x", "b200_group_scale_v", "b200_last_run_ms", "b200_initial_integrate", "b200_final_integrate", "b200_decide",
    "b200_forward_comm", "b200_reverse_comm", "b200_reneighbor", "b200_force_clear",
    "b200_pair_compute", "b200_nve_v", "b200_nve_x", "b200_scale_v", "b200_scale_v3", "b200_remap", "b200_group_scale_v3",
    "b200_group_remap", "b200_get_tallies", "b200_ke_sum", "b200_get_stats",
    "b200_get_neighbor_list", "b200_get_eam_rho_fp", "b200_pair_peratom", "b200_group_pair_peratom",
    "b200_set_profiling",
    "b200_get_phase_times", "b200_comm_unique_id", "b200_comm_init", "b200_neighbor_ranks",
    "b200_group_create", "b200_group_destroy", "b200_group_last_error", "b200_group_size",
    "b200_group_context", "b200_group_auto_grid", "b200_group_set_grid", "b200_group_set_atoms",
    "b200_group_count", "b200_group_get_atoms", "b200_group_setup", "b200_group_step",
    "b200_group_run", "b200_group_get_tallies", "b200_group_ke_sum", "b200_group_last_run_ms",
    "b200_group_get_stats", "b200_group_ke_group", "b200_ke_group", "b200_sync", "b200_set_option",
    "b200_langevin", "b200_add_force", "b200_group_langevin", "b200_group_add_force",
    "b200_set_box_triclinic", "b200_set_newton", "b200_neigh_modify_groups",
]


class B200Error(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("nbuilds", C.c_int64), ("ndanger", C.c_int64), ("ago", C.c_int64),
                ("npairs", C.c_int64), ("maxneigh", C.c_int64), ("max_numneigh", C.c_int64),
                ("nbins", C.c_int64 * 3), ("mbins", C.c_int64), ("nstencil", C.c_int64),
                ("launches", C.c_int64), ("device_bytes", C.c_double),
                ("halo_transport", C.c_int64), ("lanes_per_atom", C.c_int64),
                ("list_kind", C.c_int64), ("tile", C.c_int64 * 3), ("list_entries", C.c_int64),
                ("tile_stage_max", C.c_int64), ("tiles_interior", C.c_int64),
                ("tiles_boundary", C.c_int64), ("halo_overlap", C.c_int64)]


_lib = None


def load_library():
    """dlopen libb200md.so; raises if it has not been built (never falls back)."""
    global _lib
    if _lib is None:
        global LIBPATH
        if os.environ.get("B200_LIBPATH"):  # development: an alternative build of the same ABI
            LIBPATH = Path(os.environ["B200_LIBPATH"])
        if not LIBPATH.exists():
            raise B200Error(f"{LIBPATH} not found: run `python -m lammps_b200.build` "
                            "(there is no CPU fallback)")
        L = C.CDLL(str(LIBPATH))
        L.b200_last_error.restype = C.c_char_p
        L.b200_last_error.argtypes = [C.c_void_p]
        L.b200_destroy.restype = None
        L.b200_destroy.argtypes = [C.c_void_p]
        L.b200_group_last_error.restype = C.c_char_p
        L.b200_group_last_error.argtypes = [C.c_void_p]
        L.b200_group_destroy.restype = None
        L.b200_group_destroy.argtypes = [C.c_void_p]
        L.b200_group_context.restype = C.c_void_p
        L.b200_group_context.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, device: int = 0, precision: str = "double", units: str = "lj"):
        self.L = load_library()
        if self.L.b200_device_count() <= 0:
            raise B200Error("no CUDA device visible (the B200 engine has no CPU fallback)")
        self.h = C.c_void_p()
        prec = {"double": 0, "mixed": 1}[precision]
        rc = self.L.b200_create(C.byref(self.h), C.c_int(device), C.c_int(prec))
        self._chk(rc)
        self.precision = precision
        self.units = _units.get(units)
        self.boxlo = self.boxhi = None
        self.natoms_total = 0
        self.step = 0
        self.thermo_every = 0
        self.dt = self.units.dt

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.L.b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            msg = self.L.b200_last_error(self.h) if self.h else b""
            raise B200Error(f"b200 error {rc}: {(msg or b'').decode()}")

    # ------------------------------------------------------------ configuration
    def set_box(self, lo, hi, periodic=(1, 1, 1)):
        lo, hi, per = _d(lo), _d(hi), _i(periodic)
        self.boxlo, self.boxhi = lo.copy(), hi.copy()
        self._chk(self.L.b200_set_box(self.h, _p(lo), _p(hi), _p(per)))

    def set_box_triclinic(self, lo, hi, xy, xz, yz, periodic=(1, 1, 1), angstrom=1.0):
        lo, hi, per = _d(lo), _d(hi), _i(periodic)
        self.boxlo, self.boxhi = lo.copy(), hi.copy()
        self._chk(self.L.b200_set_box_triclinic(self.h, _p(lo), _p(hi), C.c_double(xy), C.c_double(xz),
                                                C.c_double(yz), _p(per), C.c_double(angstrom)))

    def neigh_modify_groups(self, pairs=()):
        """neigh_modify exclude group: pairs of group BITS (1 << igroup)"""
        b1 = _i([p[0] for p in pairs])
        b2 = _i([p[1] for p in pairs])
        self._chk(self.L.b200_neigh_modify_groups(self.h, C.c_int(len(pairs)), _p(b1) if len(pairs) else None,
                                                  _p(b2) if len(pairs) else None))

    def set_newton(self, newton_pair: bool):
        self._chk(self.L.b200_set_newton(self.h, C.c_int(1 if newton_pair else 0)))

    def set_decomposition(self, procgrid, myloc):
        pg, ml = _i(procgrid), _i(myloc)
        self._chk(self.L.b200_set_decomposition(self.h, _p(pg), _p(ml)))

    def neighbor(self, skin, every=1, delay=0, check=True, one=2000):
        """`neighbor SKIN bin` + `neigh_modify every E delay D check yes|no one N`."""
        self._chk(self.L.b200_set_neighbor(self.h, C.c_double(skin), C.c_int(every),
                                           C.c_int(delay), C.c_int(1 if check else 0),
                                           C.c_int(one)))

    def neigh_modify(self, once=False, exclude_types=(), ntypes=1):
        """neigh_modify once yes|no, exclude type i j (pairs of types, made symmetric)"""
        ex = None
        if exclude_types:
            n1 = ntypes + 1
            ex = np.zeros(n1 * n1, np.int32)
            for a, b in exclude_types:
                ex[a * n1 + b] = ex[b * n1 + a] = 1
        self._chk(self.L.b200_neigh_modify(self.h, C.c_int(1 if once else 0), C.c_int(ntypes), _p(ex)))

    def set_atoms(self, x, v, type, tag, mass, mask=None, image=None, natoms_total=None):
        x, v, type, tag, mass = _d(x), _d(v), _i(type), _i(tag), _d(mass)
        n = x.shape[0]
        mk = _i(mask) if mask is not None else None
        im = _i(image) if image is not None else None
        self.mass = mass
        self.natoms_total = natoms_total if natoms_total is not None else n
        self._chk(self.L.b200_set_atoms(self.h, C.c_int(n), C.c_int(mass.shape[0] - 1), _p(mass),
                                        _p(x), _p(v), _p(type), _p(tag), _p(mk), _p(im)))

    def pair_lj_cut(self, tables: dict):
        t = {k: _d(tables[k]) for k in ("cutsq", "lj1", "lj2", "lj3", "lj4", "offset")}
        sp = _d(tables.get("special_lj", np.ones(4)))
        self._chk(self.L.b200_pair_lj_cut(self.h, C.c_int(tables["ntypes"]), _p(t["cutsq"]),
                                          _p(t["lj1"]), _p(t["lj2"]), _p(t["lj3"]), _p(t["lj4"]),
                                          _p(t["offset"]), _p(sp)))

    def pair_eam(self, t: dict):
        a = {k: _i(t[k]) for k in ("type2frho", "type2rhor", "type2z2r")}
        b = {k: _d(t[k]) for k in ("scale", "frho_spline", "rhor_spline", "z2r_spline")}
        self._chk(self.L.b200_pair_eam(
            self.h, C.c_int(t["ntypes"]), C.c_int(t["nr"]), C.c_int(t["nrho"]),
            C.c_double(t["rdr"]), C.c_double(t["rdrho"]), C.c_double(t["rhomax"]),
            C.c_double(t["cutforcesq"]), _p(a["type2frho"]), _p(a["type2rhor"]),
            _p(a["type2z2r"]), _p(b["scale"]), C.c_int(t["nfrho"]), _p(b["frho_spline"]),
            C.c_int(t["nrhor"]), _p(b["rhor_spline"]), C.c_int(t["nz2r"]), _p(b["z2r_spline"])))

    def fix_nve(self, dt=None, groupbit=1):
        if dt is not None:
            self.dt = dt
        dtf = 0.5 * self.dt * self.units.ftm2v  # fix_nve.cpp:58
        self._chk(self.L.b200_fix_nve(self.h, C.c_double(self.dt), C.c_double(dtf),
                                      C.c_int(groupbit)))

    # ------------------------------------------------------------ driver
    def setup(self, eflag=1, vflag=1):
        self._chk(self.L.b200_setup(self.h, C.c_int(eflag), C.c_int(vflag)))

    def run(self, nsteps, thermo_every=None):
        """Verlet::run; returns raw tallies [[step, sum m v^2, eng_vdwl, v0..v5, 0], ...]."""
        te = self.thermo_every if thermo_every is None else thermo_every
        cap = (nsteps // te + 2) if te > 0 else 2
        out = np.zeros((cap, 10))
        n = C.c_int(0)
        rc = self.L.b200_run(self.h, C.c_int(nsteps), C.c_int64(self.step), C.c_int(te), _p(out),
                             C.c_int(cap), C.byref(n))
        self._chk(rc)
        self.step += nsteps
        return out[:n.value]

    def last_run_ms(self) -> float:
        ms = C.c_double(0)
        self._chk(self.L.b200_last_run_ms(self.h, C.byref(ms)))
        return ms.value

    def initial_integrate(self):
        self._chk(self.L.b200_initial_integrate(self.h))

    def final_integrate(self):
        self._chk(self.L.b200_final_integrate(self.h))

    def decide(self) -> int:
        r = C.c_int(0)
        self._chk(self.L.b200_decide(self.h, C.byref(r)))
        return r.value

    def forward_comm(self):
        self._chk(self.L.b200_forward_comm(self.h))

    def reverse_comm(self):
        self._chk(self.L.b200_reverse_comm(self.h))

    def reneighbor(self):
        self._chk(self.L.b200_reneighbor(self.h))

    def force_clear(self):
        self._chk(self.L.b200_force_clear(self.h))

    def pair_compute(self, eflag=1, vflag=1):
        self._chk(self.L.b200_pair_compute(self.h, C.c_int(eflag), C.c_int(vflag)))

    def nve_v(self, dtf: float, groupbit: int = 1):
        self._chk(self.L.b200_nve_v(self.h, C.c_double(dtf), C.c_int(groupbit)))

    def nve_x(self, dtv: float, groupbit: int = 1):
        self._chk(self.L.b200_nve_x(self.h, C.c_double(dtv), C.c_int(groupbit)))

    def langevin(self, gfactor1, gfactor2_tsqrt, seed: int, step: int, groupbit: int = 1,
                 uniforms_by_tag=None, want_fsum=False):
        """FixLangevin::post_force (fix_langevin.cpp:383-507) on the stored forces; per-type
        prefactors are indexed by type (entry 0 unused) like the reference's gfactor arrays."""
        g1, g2 = _d(gfactor1), _d(gfactor2_tsqrt)
        assert g1.shape == g2.shape
        u = None if uniforms_by_tag is None else _d(uniforms_by_tag).ravel()
        fs = np.zeros(3) if want_fsum else None
        self._chk(self.L.b200_langevin(self.h, C.c_int(g1.shape[0] - 1), _p(g1), _p(g2), C.c_int(groupbit),
                                       C.c_uint64(seed), C.c_int64(step), _p(u),
                                       C.c_int64(0 if u is None else u.shape[0]), _p(fs)))
        return fs

    def add_force(self, df, groupbit: int = 1):
        self._chk(self.L.b200_add_force(self.h, _p(_d(df)), C.c_int(groupbit)))

    # ------------------------------------------------------------ results
    def counts(self):
        a, b = C.c_int(0), C.c_int(0)
        self._chk(self.L.b200_get_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_atoms(self, ghosts=False, fields=("x", "v", "f", "type", "tag", "image"), into=None):
        """Download per-atom arrays in current device order.  `into` may hold preallocated
        (e.g. pinned) C-contiguous host buffers of sufficient length per field; the returned
        arrays are then views of them."""
        nl, ng = self.counts()
        n = nl + (ng if ghosts else 0)
        shapes = {"x": ((n, 3), np.float64), "v": ((nl, 3), np.float64), "f": ((n, 3), np.float64),
                  "type": ((n,), np.int32), "tag": ((n,), np.int32), "mask": ((n,), np.int32),
                  "image": ((nl,), np.int32)}
        out = {}
        for k in fields:
            shape, dt = shapes[k]
            if into is not None and k in into:
                buf = into[k]
                assert buf.dtype == dt and buf.flags.c_contiguous and buf.shape[0] >= shape[0]
                out[k] = buf[:shape[0]]
            else:
                out[k] = np.zeros(shape, dt)
        g = lambda k: _p(out[k]) if k in out else None  # noqa: E731
        self._chk(self.L.b200_get_atoms(self.h, C.c_int(1 if ghosts else 0), g("x"), g("v"),
                                        g("f"), g("type"), g("tag"), g("mask"), g("image")))
        return out

    def tallies(self):
        e = C.c_double(0)
        v = np.zeros(6)
        self._chk(self.L.b200_get_tallies(self.h, C.byref(e), _p(v)))
        return e.value, v

    def ke_sum(self):
        e = C.c_double(0)
        self._chk(self.L.b200_ke_sum(self.h, C.byref(e)))
        return e.value

    def ke_group(self, groupbit=1):
        """(sum m v^2, kinetic tensor sums[6]) of the atoms in a group bit (compute temp/b200)"""
        e = C.c_double(0)
        t = np.zeros(6)
        self._chk(self.L.b200_ke_group(self.h, C.c_int(groupbit), C.byref(e), _p(t)))
        return e.value, t

    def set_option(self, key, value):
        """a `package b200` keyword (list, tile, overlap, graph, tpa, mixed_fx, tallies)"""
        self._chk(self.L.b200_set_option(self.h, str(key).encode(), str(value).encode()))

    def sync(self):
        self._chk(self.L.b200_sync(self.h))

    def stats(self) -> dict:
        s = Stats()
        self._chk(self.L.b200_get_stats(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in Stats._fields_ if k not in ("nbins", "tile")}
        d["nbins"] = list(s.nbins)
        d["tile"] = list(s.tile)
        return d

    def neighbor_list(self):
        """(numneigh[nlocal], pair_i, pair_j) with local indices in current device order."""
        nl, _ = self.counts()
        npairs = C.c_int64(0)
        nn = np.zeros(nl, np.int32)
        self._chk(self.L.b200_get_neighbor_list(self.h, _p(nn), None, C.c_int64(0),
                                                C.byref(npairs)))
        flat = np.zeros(npairs.value, np.int32)
        self._chk(self.L.b200_get_neighbor_list(self.h, _p(nn), _p(flat),
                                                C.c_int64(npairs.value), C.byref(npairs)))
        pi = np.repeat(np.arange(nl, dtype=np.int32), nn)
        return nn, pi, flat

    def eam_rho_fp(self, ghosts=False):
        nl, ng = self.counts()
        n = nl + (ng if ghosts else 0)
        rho, fp = np.zeros(n), np.zeros(n)
        self._chk(self.L.b200_get_eam_rho_fp(self.h, C.c_int(1 if ghosts else 0), _p(rho), _p(fp)))
        return rho, fp

    def pair_peratom(self):
        """Pair::ev_tally's eatom[nlocal], vatom[nlocal][6] (xx,yy,zz,xy,xz,yz) in device order;
        call right after a setup/step that tallied"""
        nl, _ = self.counts()
        e, v = np.zeros(nl), np.zeros((nl, 6))
        self._chk(self.L.b200_pair_peratom(self.h, _p(e), _p(v)))
        return e, v

    def profiling(self, on=True):
        self._chk(self.L.b200_set_profiling(self.h, C.c_int(1 if on else 0)))

    def phase_times(self) -> dict:
        ms = np.zeros(NPHASE)
        calls = np.zeros(NPHASE, np.int64)
        self._chk(self.L.b200_get_phase_times(self.h, _p(ms), _p(calls)))
        return {PHASES[k]: (float(ms[k]), int(calls[k])) for k in range(NPHASE)}

    # ------------------------------------------------------------ thermo (host arithmetic)
    def thermo_row(self, raw) -> dict:
        """Temp / E_pair / TotEng / Press as thermo.cpp prints them, from the raw tallies.
        compute_temp.cpp:57-97 (dof = 3N-3), compute_pressure.cpp:254-259, thermo norm."""
        u = self.units
        n = self.natoms_total
        dof = 3.0 * n - 3.0
        tfactor = u.mvv2e / (dof * u.boltz)
        temp = raw[1] * tfactor
        vol = float(np.prod(self.boxhi - self.boxlo))
        press = (dof * u.boltz * temp + raw[3] + raw[4] + raw[5]) / 3.0 / vol * u.nktv2p
        ke = 0.5 * dof * u.boltz * temp
        pe = raw[2]
        norm = n if u.normalize else 1
        return {"step": int(raw[0]), "temp": temp, "e_pair": pe / norm,
                "toteng": (pe + ke) / norm, "press": press}


class EngineGroup:
    """Several brick sub-domains driven by this one process (`package b200 gpus N`): the mirror of
    the b200_group_* entry points.  `devices` lists the CUDA device of every sub-domain; devices
    may repeat (several sub-domains sharing one GPU: how a 1-GPU box exercises migration, borders
    and the peer-memory halo).  Same vocabulary as `Engine`; per-context settings are applied to
    every sub-domain."""

    def __init__(self, devices, precision="double", units="lj", grid=None):
        self.L = load_library()
        if self.L.b200_device_count() <= 0:
            raise B200Error("no CUDA device visible (the B200 engine has no CPU fallback)")
        dev = _i(devices)
        self.n = len(dev)
        self.g = C.c_void_p()
        prec = {"double": 0, "mixed": 1}[precision]
        self._chk(self.L.b200_group_create(C.byref(self.g), C.c_int(self.n), _p(dev), C.c_int(prec)))
        # sub-domain views: Engine objects around the group's contexts (not owning them)
        self.sub = []
        for i in range(self.n):
            e = Engine.__new__(Engine)
            e.L = self.L
            e.h = C.c_void_p(self.L.b200_group_context(self.g, C.c_int(i)))
            e.precision, e.units = precision, _units.get(units)
            e.boxlo = e.boxhi = None
            e.natoms_total, e.step, e.thermo_every, e.dt = 0, 0, 0, e.units.dt
            e.close = lambda: None
            self.sub.append(e)
        self.units = _units.get(units)
        self.grid = grid
        self.step = 0
        self.natoms_total = 0
        self.boxlo = self.boxhi = None
        self.dt = self.units.dt

    def close(self):
        if getattr(self, "g", None) and self.g:
            for e in self.sub:
                e.h = C.c_void_p()
            self.L.b200_group_destroy(self.g)
            self.g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            msg = self.L.b200_group_last_error(self.g) if self.g else b""
            raise B200Error(f"b200 group error {rc}: {(msg or b'').decode()}")

    def set_box(self, lo, hi, periodic=(1, 1, 1)):
        self.boxlo, self.boxhi = _d(lo).copy(), _d(hi).copy()
        for e in self.sub:
            e.set_box(lo, hi, periodic)
        grid = self.grid
        if grid is None:
            g3 = np.zeros(3, np.int32)
            prd = _d(hi) - _d(lo)
            self._chk(self.L.b200_group_auto_grid(C.c_int(self.n), _p(prd), _p(g3)))
            grid = tuple(int(v) for v in g3)
        self.grid = tuple(grid)
        self._chk(self.L.b200_group_set_grid(self.g, _p(_i(self.grid))))

    def set_box_triclinic(self, lo, hi, xy, xz, yz, periodic=(1, 1, 1), angstrom=1.0):
        self.set_box(lo, hi, periodic)
        for e in self.sub:
            e.set_box_triclinic(lo, hi, xy, xz, yz, periodic, angstrom)

    def neighbor(self, *a, **k):
        for e in self.sub:
            e.neighbor(*a, **k)

    def fix_nve(self, dt=None, groupbit=1):
        for e in self.sub:
            e.fix_nve(dt, groupbit)
        self.dt = self.sub[0].dt

    def pair_lj_cut(self, tables):
        for e in self.sub:
            e.pair_lj_cut(tables)

    def pair_eam(self, tables):
        for e in self.sub:
            e.pair_eam(tables)

    def set_atoms(self, x, v, type, tag, mass, mask=None, image=None, natoms_total=None):
        x, v, type, tag, mass = _d(x), _d(v), _i(type), _i(tag), _d(mass)
        n = x.shape[0]
        mk = _i(mask) if mask is not None else None
        im = _i(image) if image is not None else None
        self.natoms_total = natoms_total if natoms_total is not None else n
        for e in self.sub:
            e.natoms_total, e.mass, e.boxlo, e.boxhi = self.natoms_total, mass, self.boxlo, self.boxhi
        self._chk(self.L.b200_group_set_atoms(self.g, C.c_int(n), C.c_int(mass.shape[0] - 1), _p(mass),
                                              _p(x), _p(v), _p(type), _p(tag), _p(mk), _p(im)))

    def setup(self, eflag=1, vflag=1):
        self._chk(self.L.b200_group_setup(self.g, C.c_int(eflag), C.c_int(vflag)))

    def run(self, nsteps, thermo_every=0):
        cap = (nsteps // thermo_every + 2) if thermo_every > 0 else 2
        out = np.zeros((cap, 10))
        n = C.c_int(0)
        self._chk(self.L.b200_group_run(self.g, C.c_int(nsteps), C.c_int64(self.step),
                                        C.c_int(thermo_every), _p(out), C.c_int(cap), C.byref(n)))
        self.step += nsteps
        return out[:n.value]

    def last_run_ms(self):
        ms = C.c_double(0)
        self._chk(self.L.b200_group_last_run_ms(self.g, C.byref(ms)))
        return ms.value

    def counts(self):
        a, b = C.c_int(0), C.c_int(0)
        self._chk(self.L.b200_group_count(self.g, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_atoms(self, fields=("x", "v", "f", "type", "tag", "image")):
        n, _ = self.counts()
        shapes = {"x": ((n, 3), np.float64), "v": ((n, 3), np.float64), "f": ((n, 3), np.float64),
                  "type": ((n,), np.int32), "tag": ((n,), np.int32), "mask": ((n,), np.int32),
                  "image": ((n,), np.int32)}
        out = {k: np.zeros(*shapes[k]) for k in fields}
        g = lambda k: _p(out[k]) if k in out else None  # noqa: E731
        self._chk(self.L.b200_group_get_atoms(self.g, g("x"), g("v"), g("f"), g("type"), g("tag"),
                                              g("mask"), g("image")))
        return out

    def pair_peratom(self):
        """eatom[n], vatom[n][6] of all sub-domains, in the order get_atoms returns the atoms"""
        n, _ = self.counts()
        e, v = np.zeros(n), np.zeros((n, 6))
        self._chk(self.L.b200_group_pair_peratom(self.g, _p(e), _p(v)))
        return e, v

    def tallies(self):
        e = C.c_double(0)
        v = np.zeros(6)
        self._chk(self.L.b200_group_get_tallies(self.g, C.byref(e), _p(v)))
        return e.value, v

    def ke_sum(self):
        e = C.c_double(0)
        self._chk(self.L.b200_group_ke_sum(self.g, C.byref(e)))
        return e.value

    def stats(self):
        s = Stats()
        self._chk(self.L.b200_group_get_stats(self.g, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in Stats._fields_ if k not in ("nbins", "tile")}
        d["nbins"], d["tile"] = list(s.nbins), list(s.tile)
        return d

    def thermo_row(self, raw):
        e = self.sub[0]
        e.natoms_total, e.boxlo, e.boxhi = self.natoms_total, self.boxlo, self.boxhi
        return e.thermo_row(raw)
