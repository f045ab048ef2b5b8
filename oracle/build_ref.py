#!/usr/bin/env python3
"""Build recipe for oracle/_ref: the UNMODIFIED reference (LAMMPS) CPU code path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.

What it does
------------
Compiles the reference's own C++ sources *where they lie* under /root/reference/src
(never copied into this repo) with /usr/bin/g++ into

    oracle/_ref/liblammps_ref.so   (library API: lammps_open_no_mpi, lammps_command, ...)
    oracle/_ref/lmp_ref            (the `lmp` executable)
    oracle/_ref/potentials/        (potential tables the bench inputs read: Cu_u3.eam, Al_jnp.eam)

using  core src/*.cpp + STUBS (serial MPI) + MANYBODY/pair_eam*.cpp + the OPENMP and OPT
styles whose base style is in that set.  The reference's own build systems (cmake /
src/Makefile) are NOT run; the only "generated code" LAMMPS needs is the list of style
headers (what src/Make.sh `style` greps for), which `gen_style_headers` re-derives with
the same rule (a header that mentions FOO_CLASS is included from style_foo.h).

oracle/_ref/ is git-ignored but NOT gpurun-ignored, so the built files travel to the
GPU box where /root/reference does not exist.

Usage:  python oracle/build_ref.py [-j N] [--b200]   (--b200 also links lmp_b200, see
        lammps_b200/lammps_pkg/build_pkg.py which imports this module)
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("LAMMPS_REF", "/root/reference"))
SRC = REF / "src"
OUT = REPO / "oracle" / "_ref"
CXX = os.environ.get("REF_CXX", "/usr/bin/g++")
CXXFLAGS = ["-O3", "-std=c++17", "-fopenmp", "-fPIC", "-DLAMMPS_SMALLBIG", "-DLMP_OPENMP",
            "-DLAMMPS_EXCEPTIONS", "-w"]

# (macro, filename prefix, style file name)  -- same table as src/Make.sh:75-97
STYLE_TABLE = [
    ("ANGLE_CLASS", "angle_", "angle"), ("ATOM_CLASS", "atom_vec_", "atom"),
    ("BODY_CLASS", "body_", "body"), ("BOND_CLASS", "bond_", "bond"),
    ("COMMAND_CLASS", "", "command"), ("COMPUTE_CLASS", "compute_", "compute"),
    ("DIHEDRAL_CLASS", "dihedral_", "dihedral"), ("DUMP_CLASS", "dump_", "dump"),
    ("FIX_CLASS", "fix_", "fix"), ("GRAN_SUB_MOD_CLASS", "gran_sub_mod_", "gran_sub_mod"),
    ("IMPROPER_CLASS", "improper_", "improper"), ("INTEGRATE_CLASS", "", "integrate"),
    ("KSPACE_CLASS", "", "kspace"), ("MINIMIZE_CLASS", "min_", "minimize"),
    ("NBIN_CLASS", "nbin_", "nbin"), ("NPAIR_CLASS", "npair_", "npair"),
    ("NSTENCIL_CLASS", "nstencil_", "nstencil"), ("NTOPO_CLASS", "ntopo_", "ntopo"),
    ("PAIR_CLASS", "pair_", "pair"), ("READER_CLASS", "reader_", "reader"),
    ("REGION_CLASS", "region_", "region"),
]


def source_set() -> tuple[list[Path], list[Path]]:
    """Return (cpp files, include dirs) of the reference subset we build."""
    core = sorted(SRC.glob("*.cpp"))
    chosen = list(core)
    have = {p.name for p in core}
    many = [SRC / "MANYBODY" / n for n in ("pair_eam.cpp", "pair_eam_alloy.cpp", "pair_eam_fs.cpp")]
    chosen += many
    have |= {p.name for p in many}
    # OPENMP / OPT: a suffix style is installed only if its base style exists
    # (rule of src/OPENMP/Install.sh:31-35 and src/OPT/Install.sh)
    for pkg, suf in (("OPENMP", "_omp"), ("OPT", "_opt")):
        for p in sorted((SRC / pkg).glob(f"*{suf}.cpp")):
            base = p.name.replace(f"{suf}.cpp", ".cpp")
            if p.name == "thr_omp.cpp" or base in have:
                chosen.append(p)
    chosen.append(SRC / "OPENMP" / "thr_data.cpp")
    if SRC / "OPENMP" / "thr_omp.cpp" not in chosen:
        chosen.append(SRC / "OPENMP" / "thr_omp.cpp")
    chosen.append(SRC / "STUBS" / "mpi.cpp")
    incs = [SRC, SRC / "STUBS", SRC / "MANYBODY", SRC / "OPENMP", SRC / "OPT"]
    return chosen, incs


def gen_style_headers(gen: Path, cpp_files: list[Path], extra_headers: list[Path] = ()) -> None:
    """Re-derive style_*.h / packages_*.h / lmpinstalledpkgs.h / lmpgitversion.h."""
    gen.mkdir(parents=True, exist_ok=True)
    headers = []
    for c in cpp_files:
        h = c.with_suffix(".h")
        if h.exists():
            headers.append(h)
    headers += list(extra_headers)
    texts = {h: h.read_text(errors="replace") for h in headers}
    for macro, prefix, name in STYLE_TABLE:
        lines = [f'#include "{h.name}"' for h in sorted(headers, key=lambda p: p.name)
                 if h.name.startswith(prefix) and macro in texts[h]]
        _write_if_changed(gen / f"style_{name}.h", "\n".join(lines) + ("\n" if lines else ""))
        _write_if_changed(gen / f"packages_{name}.h", "")
    pk = ('const char * LAMMPS_NS::LAMMPS::installed_packages[] = '
          '{"MANYBODY", "OPENMP", "OPT", NULL};\n')
    _write_if_changed(gen / "lmpinstalledpkgs.h", pk)
    gv = ('bool LAMMPS_NS::LAMMPS::has_git_info() { return false; }\n'
          'const char *LAMMPS_NS::LAMMPS::git_commit() { return "(unknown)"; }\n'
          'const char *LAMMPS_NS::LAMMPS::git_branch() { return "(unknown)"; }\n'
          'const char *LAMMPS_NS::LAMMPS::git_descriptor() { return "(unknown)"; }\n')
    _write_if_changed(gen / "lmpgitversion.h", gv)


def _write_if_changed(p: Path, s: str) -> None:
    if not p.exists() or p.read_text() != s:
        p.write_text(s)


def compile_all(files: list[Path], incs: list[Path], objdir: Path, jobs: int,
                extra_flags: list[str] = ()) -> list[Path]:
    objdir.mkdir(parents=True, exist_ok=True)
    objs, todo = [], []
    for f in files:
        tag = f.parent.name if f.parent != SRC else "core"
        o = objdir / f"{tag}__{f.stem}.o"
        objs.append(o)
        if not o.exists() or o.stat().st_mtime < f.stat().st_mtime:
            todo.append((f, o))
    inc_flags = [f"-I{i}" for i in incs]

    def one(fo):
        f, o = fo
        cmd = [CXX, *CXXFLAGS, *extra_flags, *inc_flags, "-c", str(f), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return f, r.returncode, r.stderr

    if todo:
        print(f"[build_ref] compiling {len(todo)} files with -j{jobs} ...", flush=True)
    with cf.ThreadPoolExecutor(jobs) as ex:
        for f, rc, err in ex.map(one, todo):
            if rc != 0:
                sys.stderr.write(err[-4000:])
                raise SystemExit(f"[build_ref] failed: {f}")
    return objs


def build(jobs: int = 8) -> None:
    if not SRC.exists():
        raise SystemExit(f"[build_ref] reference sources not found at {SRC}")
    files, incs = source_set()
    gen = OUT / "gen"
    gen_style_headers(gen, files)
    objs = compile_all(files, [gen, *incs], OUT / "obj", jobs)
    main_o = OUT / "obj" / "core__main.o"
    lib_objs = [o for o in objs if o != main_o]
    lib = OUT / "liblammps_ref.so"
    if not lib.exists() or any(o.stat().st_mtime > lib.stat().st_mtime for o in lib_objs):
        subprocess.check_call([CXX, "-shared", "-fopenmp", "-o", str(lib), *map(str, lib_objs),
                               "-ldl", "-lpthread"])
    exe = OUT / "lmp_ref"
    if not exe.exists() or exe.stat().st_mtime < lib.stat().st_mtime:
        subprocess.check_call([CXX, "-fopenmp", "-o", str(exe), str(main_o), f"-L{OUT}",
                               "-llammps_ref", f"-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"])
    pot = OUT / "potentials"
    pot.mkdir(exist_ok=True)
    for name in ("Cu_u3.eam", "Al_jnp.eam"):
        if not (pot / name).exists():
            shutil.copy(REF / "potentials" / name, pot / name)
    print(f"[build_ref] ok: {lib} {exe}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=os.cpu_count() or 8)
    a = ap.parse_args()
    build(a.j)
