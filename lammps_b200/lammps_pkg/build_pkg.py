#!/usr/bin/env python3
"""Builds lammps_b200/lammps_pkg/lmp_b200: the reference LAMMPS with the B200 package compiled in.

What is compiled
----------------
* lammps_b200/lammps_pkg/B200/*.cpp      -- the package (pair lj/cut/b200, eam/b200, eam/alloy/b200,
                                            eam/fs/b200, fix nve/b200, fix B200, run_style verlet/b200)
* force.cpp, update.cpp                  -- UNMODIFIED reference sources, recompiled only because they
                                            include the generated style_pair.h / style_fix.h /
                                            style_integrate.h, which now also list the B200 headers
                                            (what cmake's RegisterStylesExt / Make.sh `style` do)
* input.cpp, lammps.cpp, modify.cpp      -- reference sources + the core patch below (applied to a
                                            scratch copy under _build/, never committed): the
                                            `package b200` command and the `-sf b200` default
* every other object                     -- the reference's, compiled once into build/ref_obj by
                                            tools/ref_compile.py (same cache the checker build uses)
and linked against lammps_b200/libb200md.so (the C ABI, include/b200_md.h).

The reference's own build system is not run; nothing is written under /root/reference.
Usage: python lammps_b200/lammps_pkg/build_pkg.py
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))

import ref_compile as R  # noqa: E402

PKG = HERE / "B200"
BUILD = HERE / "_build"
EXE = HERE / "lmp_b200"

# ---- the core patch: (file, anchor text that must occur exactly once, replacement) ----------
PACKAGE_BRANCH = '''  } else if (strcmp(arg[0],"b200") == 0) {
    if (!modify->check_package("B200"))
      error->all(FLERR, Error::ARGZERO, "Package b200 command without B200 package installed");

    std::string fixcmd = "package_b200 all B200";
    for (int i = 1; i < narg; i++) fixcmd += std::string(" ") + arg[i];
    modify->add_fix(fixcmd);

'''
CORE_PATCH = [
    # fix B200 may be created before the box exists, like fix GPU / OMP / INTEL
    ("modify.cpp", '{"GPU", "OMP", "INTEL", "property/atom",', '{"GPU", "OMP", "INTEL", "B200", "property/atom",'),
    ("input.cpp",
     '  } else error->all(FLERR, Error::ARGZERO, "Unknown package keyword: {}", arg[0]);',
     PACKAGE_BRANCH +
     '  } else error->all(FLERR, Error::ARGZERO, "Unknown package keyword: {}", arg[0]);'),
    ("lammps.cpp",
     '        if (strcmp("intel", pkg_name) == 0) package_issued |= Suffix::INTEL;',
     '        if (strcmp("intel", pkg_name) == 0) package_issued |= Suffix::INTEL;\n'
     '        if (strcmp("b200", pkg_name) == 0) package_issued |= (1 << 5);    // Suffix::B200'),
    ("lammps.cpp",
     '    if (strcmp(suffix,"omp") == 0 && !(package_issued & Suffix::OMP))\n'
     '      input->one("package omp 0");\n',
     '    if (strcmp(suffix,"omp") == 0 && !(package_issued & Suffix::OMP))\n'
     '      input->one("package omp 0");\n'
     '    if (strcmp(suffix,"b200") == 0 && !modify->check_package("B200"))\n'
     '      error->all(FLERR,"Using suffix b200 without B200 package installed");\n'
     '    if (strcmp(suffix,"b200") == 0 && !(package_issued & (1 << 5)))\n'
     '      input->one("package b200");\n'),
]


def patched_core() -> list[Path]:
    out = []
    BUILD.mkdir(parents=True, exist_ok=True)
    texts = {}
    for fname, anchor, repl in CORE_PATCH:
        t = texts.get(fname) or (R.SRC / fname).read_text()
        if t.count(anchor) != 1:
            raise SystemExit(f"[build_pkg] core patch anchor not found exactly once in {fname}")
        texts[fname] = t.replace(anchor, repl)
    for fname, t in texts.items():
        p = BUILD / fname
        if not p.exists() or p.read_text() != t:
            p.write_text(t)
        out.append(p)
    return out


def build(jobs: int = 8) -> Path:
    lib = REPO / "lammps_b200" / "libb200md.so"
    if not lib.exists():
        from lammps_b200 import build as B
        B.build()
    all_ref_objs = R.compile_reference(jobs)
    ref_files, incs = R.source_set()
    gen = BUILD / "gen"
    pkg_cpp = sorted(PKG.glob("*.cpp"))
    pkg_h = sorted(PKG.glob("*.h"))
    R.gen_style_headers(gen, ref_files, extra_headers=pkg_h)
    (gen / "lmpinstalledpkgs.h").write_text(
        'const char * LAMMPS_NS::LAMMPS::installed_packages[] = '
        '{"B200", "MANYBODY", "OPENMP", "OPT", NULL};\n')
    inc = [gen, PKG, REPO / "include", *incs]
    objdir = BUILD / "obj"
    patched = patched_core()
    restyle = [R.SRC / n for n in ("force.cpp", "update.cpp")]
    new_objs = R.compile_all(pkg_cpp + patched + restyle, inc, objdir, jobs,
                             deps=[*pkg_h, REPO / "include" / "b200_md.h"])
    replaced = {"core__force.o", "core__modify.o", "core__update.o", "core__input.o", "core__lammps.o"}
    ref_objs = [o for o in all_ref_objs if o.name not in replaced]
    newest = max(o.stat().st_mtime for o in new_objs + [lib])
    if not EXE.exists() or EXE.stat().st_mtime < newest:
        cmd = [R.CXX, "-fopenmp", "-o", str(EXE), *map(str, new_objs), *map(str, ref_objs),
               f"-L{lib.parent}", "-lb200md", "-Wl,-rpath,$ORIGIN/..", "-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    # the reference's own bench inputs, untouched (git-ignored; they travel to the GPU box where
    # tests run them with -sf b200)
    bench = HERE / "bench_inputs"
    bench.mkdir(exist_ok=True)
    for name in ("in.lj", "in.eam", "Cu_u3.eam"):
        src = R.REF / "bench" / name
        if src.exists() and not (bench / name).exists():
            shutil.copy(src, bench / name)
    print(f"[build_pkg] ok: {EXE}")
    return EXE


if __name__ == "__main__":
    build(os.cpu_count() or 8)
