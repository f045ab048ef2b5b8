"""ctypes harness for oracle/_ref/liblammps_ref.so -- the UNMODIFIED reference, compiled by
oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY (used to pin md_oracle.c and to generate the
fixtures in tests/golden/).  Binds the reference's C library API, src/library.h:138-330.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "liblammps_ref.so"
EXE = HERE / "_ref" / "lmp_ref"
POTENTIALS = HERE / "_ref" / "potentials"


def available() -> bool:
    return LIB.exists()


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(LIB), mode=C.RTLD_GLOBAL)
        lib.lammps_open_no_mpi.restype = C.c_void_p
        lib.lammps_open_no_mpi.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_void_p]
        lib.lammps_close.argtypes = [C.c_void_p]
        lib.lammps_command.restype = C.c_void_p
        lib.lammps_command.argtypes = [C.c_void_p, C.c_char_p]
        lib.lammps_commands_string.argtypes = [C.c_void_p, C.c_char_p]
        lib.lammps_get_natoms.restype = C.c_double
        lib.lammps_get_natoms.argtypes = [C.c_void_p]
        lib.lammps_get_thermo.restype = C.c_double
        lib.lammps_get_thermo.argtypes = [C.c_void_p, C.c_char_p]
        lib.lammps_extract_setting.restype = C.c_int
        lib.lammps_extract_setting.argtypes = [C.c_void_p, C.c_char_p]
        lib.lammps_extract_atom.restype = C.c_void_p
        lib.lammps_extract_atom.argtypes = [C.c_void_p, C.c_char_p]
        lib.lammps_extract_global.restype = C.c_void_p
        lib.lammps_extract_global.argtypes = [C.c_void_p, C.c_char_p]
        lib.lammps_extract_pair.restype = C.c_void_p
        lib.lammps_extract_pair.argtypes = [C.c_void_p, C.c_char_p]
        lib.lammps_find_pair_neighlist.restype = C.c_int
        lib.lammps_find_pair_neighlist.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        lib.lammps_neighlist_num_elements.restype = C.c_int
        lib.lammps_neighlist_num_elements.argtypes = [C.c_void_p, C.c_int]
        lib.lammps_neighlist_element_neighbors.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
            C.POINTER(C.POINTER(C.c_int))]
        lib.lammps_has_error.restype = C.c_int
        lib.lammps_has_error.argtypes = [C.c_void_p]
        lib.lammps_get_last_error_message.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.lammps_extract_box.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        lib.lammps_extract_compute.restype = C.c_void_p
        lib.lammps_extract_compute.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        _lib = lib
    return _lib


class RefLammps:
    """A live instance of the reference, serial (MPI STUBS)."""

    def __init__(self, args=("-log", "none", "-screen", "none")):
        lib = _load()
        argv = [b"lmp_ref"] + [a.encode() for a in args]
        arr = (C.c_char_p * len(argv))(*argv)
        self.lib = lib
        self.h = lib.lammps_open_no_mpi(len(argv), arr, None)
        if not self.h:
            raise RuntimeError("lammps_open_no_mpi failed")

    def close(self):
        if self.h:
            self.lib.lammps_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self):
        if self.lib.lammps_has_error(self.h):
            buf = C.create_string_buffer(2048)
            self.lib.lammps_get_last_error_message(self.h, buf, 2048)
            raise RuntimeError("reference LAMMPS error: " + buf.value.decode())

    def command(self, cmd: str):
        self.lib.lammps_command(self.h, cmd.encode())
        self._check()

    def commands(self, text: str):
        self.lib.lammps_commands_string(self.h, text.encode())
        self._check()

    def setting(self, name: str) -> int:
        return self.lib.lammps_extract_setting(self.h, name.encode())

    def thermo(self, key: str) -> float:
        return self.lib.lammps_get_thermo(self.h, key.encode())

    def natoms(self) -> int:
        return int(self.lib.lammps_get_natoms(self.h))

    def box(self):
        lo = (C.c_double * 3)()
        hi = (C.c_double * 3)()
        self.lib.lammps_extract_box(self.h, lo, hi, None, None, None, None, None)
        return np.array(lo), np.array(hi)

    def atom_vec3(self, name: str, n: int) -> np.ndarray:
        """x / v / f: double** whose rows are one contiguous block (memory.h:144-160)."""
        p = self.lib.lammps_extract_atom(self.h, name.encode())
        rows = C.cast(p, C.POINTER(C.POINTER(C.c_double)))
        flat = np.ctypeslib.as_array(rows[0], shape=(n * 3,))
        return flat.reshape(n, 3).copy()

    def atom_int(self, name: str, n: int) -> np.ndarray:
        p = self.lib.lammps_extract_atom(self.h, name.encode())
        arr = C.cast(p, C.POINTER(C.c_int))
        return np.ctypeslib.as_array(arr, shape=(n,)).copy()

    def compute_peratom(self, cid: str, n: int, ncols: int = 0) -> np.ndarray:
        """per-atom vector (ncols = 0) or array of a compute, first n atoms (library.h:
        LMP_STYLE_ATOM = 1, LMP_TYPE_VECTOR = 1, LMP_TYPE_ARRAY = 2); the compute must be current
        (used by the thermo output of this step)"""
        p = self.lib.lammps_extract_compute(self.h, cid.encode(), 1, 2 if ncols else 1)
        self._check()
        if not p:
            raise RuntimeError(f"compute {cid} has no per-atom data")
        if ncols:
            rows = C.cast(p, C.POINTER(C.POINTER(C.c_double)))
            return np.ctypeslib.as_array(rows[0], shape=(n * ncols,)).reshape(n, ncols).copy()
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)).copy()

    def pair_extract_scalar(self, name: str) -> float:
        p = self.lib.lammps_extract_pair(self.h, name.encode())
        return C.cast(p, C.POINTER(C.c_double))[0] if p else float("nan")

    def neighbor_pairs(self, style: str):
        """All (i, j) of the pair style's perpetual list as local indices, j with the
        special-bond bits masked off (lmptype.h:61-65)."""
        idx = self.lib.lammps_find_pair_neighlist(self.h, style.encode(), 1, 0, 0)
        if idx < 0:
            raise RuntimeError(f"no neighbor list for pair {style}")
        n = self.lib.lammps_neighlist_num_elements(self.h, idx)
        ii, jj = [], []
        iatom = C.c_int()
        num = C.c_int()
        ptr = C.POINTER(C.c_int)()
        for e in range(n):
            self.lib.lammps_neighlist_element_neighbors(self.h, idx, e, C.byref(iatom),
                                                        C.byref(num), C.byref(ptr))
            if num.value:
                js = np.ctypeslib.as_array(ptr, shape=(num.value,)) & 0x1FFFFFFF
                jj.append(js.copy())
                ii.append(np.full(num.value, iatom.value, dtype=np.int32))
        if not ii:
            return np.zeros(0, np.int32), np.zeros(0, np.int32)
        return np.concatenate(ii), np.concatenate(jj)


# The two bench inputs, restated (same commands as bench/in.lj and bench/in.eam of the
# reference; the files themselves are not copied).  `nx,ny,nz` are the -var x/y/z factors.
def lj_input(nx=1, ny=1, nz=1, run=100, cells=20, extra="") -> str:
    return f"""
units           lj
atom_style      atomic
lattice         fcc 0.8442
region          box block 0 {cells * nx} 0 {cells * ny} 0 {cells * nz}
create_box      1 box
create_atoms    1 box
mass            1 1.0
velocity        all create 1.44 87287 loop geom
pair_style      lj/cut 2.5
pair_coeff      1 1 1.0 1.0 2.5
neighbor        0.3 bin
neigh_modify    delay 0 every 20 check no
fix             1 all nve
{extra}
run             {run}
"""


def eam_input(nx=1, ny=1, nz=1, run=100, cells=20, potential=None, extra="") -> str:
    potential = potential or str(POTENTIALS / "Cu_u3.eam")
    return f"""
units           metal
atom_style      atomic
lattice         fcc 3.615
region          box block 0 {cells * nx} 0 {cells * ny} 0 {cells * nz}
create_box      1 box
create_atoms    1 box
pair_style      eam
pair_coeff      1 1 {potential}
velocity        all create 1600.0 376847 loop geom
neighbor        1.0 bin
neigh_modify    every 1 delay 5 check yes
fix             1 all nve
timestep        0.005
thermo          50
{extra}
run             {run}
"""


def run_exe(script: str, threads: int = 1, suffix: str | None = None, cwd=None, timeout=3600):
    """Run lmp_ref on a script; returns stdout.  threads>1 uses the reference's OPENMP
    package (-sf omp -pk omp N), the only multi-core path available without MPI."""
    import subprocess
    import tempfile
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as fh:
        fh.write(script)
        path = fh.name
    cmd = [str(EXE), "-in", path, "-log", "none"]
    if suffix:
        cmd += ["-sf", suffix]
        if suffix == "omp":
            cmd += ["-pk", "omp", str(threads)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=cwd, timeout=timeout)
    finally:
        os.unlink(path)
    if r.returncode != 0:
        raise RuntimeError(f"lmp_ref failed: {r.stdout[-2000:]} {r.stderr[-2000:]}")
    return r.stdout
