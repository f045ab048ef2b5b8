#!/usr/bin/env python3
"""Recipe for compiling the UNMODIFIED reference (LAMMPS) sources where they lie.

Shared by two builds, neither of which runs the reference's own build system:
  * oracle/build_ref.py            -> oracle/_ref/{liblammps_ref.so, lmp_ref}   (the checker)
  * lammps_b200/lammps_pkg/build_pkg.py -> lmp_b200 = the same LAMMPS + the B200 package (the host
                                          program the package plugs into)
Objects are cached under build/ref_obj (git-ignored, not shipped to the GPU box).

Sources compiled: core src/*.cpp + STUBS (serial MPI) + MANYBODY/pair_eam*.cpp + the OPENMP and
OPT styles whose base style is in that set.  The only "generated code" LAMMPS needs is the list
of style headers (what src/Make.sh `style` greps for), which gen_style_headers re-derives with
the same rule (a header that mentions FOO_CLASS is included from style_foo.h).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("LAMMPS_REF", "/root/reference"))
SRC = REF / "src"
OBJ = REPO / "build" / "ref_obj"
GEN = REPO / "build" / "ref_gen"
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
                extra_flags: list[str] = (), deps: list[Path] = ()) -> list[Path]:
    """Compile `files` into objdir; an object is stale when its source or any of `deps`
    (headers outside the reference tree, e.g. the C ABI) is newer."""
    objdir.mkdir(parents=True, exist_ok=True)
    objs, todo = [], []
    dep_time = max((d.stat().st_mtime for d in deps), default=0.0)
    for f in files:
        tag = f.parent.name if f.parent != SRC else "core"
        o = objdir / f"{tag}__{f.stem}.o"
        objs.append(o)
        if not o.exists() or o.stat().st_mtime < max(f.stat().st_mtime, dep_time):
            todo.append((f, o))
    inc_flags = [f"-I{i}" for i in incs]

    def one(fo):
        f, o = fo
        cmd = [CXX, *CXXFLAGS, *extra_flags, *inc_flags, "-c", str(f), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return f, r.returncode, r.stderr

    if todo:
        print(f"[ref_compile] compiling {len(todo)} files with -j{jobs} ...", flush=True)
    with cf.ThreadPoolExecutor(jobs) as ex:
        for f, rc, err in ex.map(one, todo):
            if rc != 0:
                sys.stderr.write(err[-4000:])
                raise SystemExit(f"[ref_compile] failed: {f}")
    return objs



def compile_reference(jobs: int = 8) -> list[Path]:
    """All reference objects (cached); returns their paths."""
    if not SRC.exists():
        raise SystemExit(f"[ref_compile] reference sources not found at {SRC}")
    files, incs = source_set()
    gen_style_headers(GEN, files)
    return compile_all(files, [GEN, *incs], OBJ, jobs)
