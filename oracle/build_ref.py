#!/usr/bin/env python3
"""Build recipe for oracle/_ref: the UNMODIFIED reference (LAMMPS) CPU code path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.

Compiles the reference's own C++ sources *where they lie* under /root/reference/src (never
copied into this repo; recipe in tools/ref_compile.py, the reference's cmake / Makefile are not
run) with /usr/bin/g++ into

    oracle/_ref/liblammps_ref.so   (library API: lammps_open_no_mpi, lammps_command, ...)
    oracle/_ref/lmp_ref            (the `lmp` executable)
    oracle/_ref/potentials/        (potential tables the bench inputs read: Cu_u3.eam, Al_jnp.eam)

oracle/_ref/ is git-ignored but NOT gpurun-ignored, so the built files travel to the GPU box
where /root/reference does not exist.

Usage:  python oracle/build_ref.py [-j N]
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "tools"))
import ref_compile as RC  # noqa: E402

OUT = REPO / "oracle" / "_ref"


def build(jobs: int = 8) -> None:
    objs = RC.compile_reference(jobs)
    OUT.mkdir(parents=True, exist_ok=True)
    main_o = RC.OBJ / "core__main.o"
    lib_objs = [o for o in objs if o != main_o]
    lib = OUT / "liblammps_ref.so"
    if not lib.exists() or any(o.stat().st_mtime > lib.stat().st_mtime for o in lib_objs):
        subprocess.check_call([RC.CXX, "-shared", "-fopenmp", "-o", str(lib), *map(str, lib_objs),
                               "-ldl", "-lpthread"])
    exe = OUT / "lmp_ref"
    if not exe.exists() or exe.stat().st_mtime < lib.stat().st_mtime:
        subprocess.check_call([RC.CXX, "-fopenmp", "-o", str(exe), str(main_o), f"-L{OUT}",
                               "-llammps_ref", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"])
    pot = OUT / "potentials"
    pot.mkdir(exist_ok=True)
    for name in ("Cu_u3.eam", "Al_jnp.eam"):
        if not (pot / name).exists():
            shutil.copy(RC.REF / "potentials" / name, pot / name)
    print(f"[build_ref] ok: {lib} {exe}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=os.cpu_count() or 8)
    a = ap.parse_args()
    build(a.j)
