"""Builds lammps_b200/libb200md.so from csrc/ with nvcc for sm_100a (cross-compiles without a
GPU).  In-tree on purpose: the .so travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = Path(os.environ.get("B200_LIB_OUT", HERE / "libb200md.so"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("B200_NVCC_EXTRA", "").split()
FLAGS = [*EXTRA, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(CSRC.glob("*.cu")), sorted(list(CSRC.glob("*.cuh")) +
                                              [HERE.parent / "include" / "b200_md.h"])


def build(force: bool = False, verbose: bool = False) -> Path:
    cu, deps = sources()
    newest = max(p.stat().st_mtime for p in cu + deps)
    if not force and LIB.exists() and LIB.stat().st_mtime >= newest:
        return LIB
    cmd = [NVCC, *FLAGS, "-o", str(LIB), *map(str, cu)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = HERE / "build.log"
    log.write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-6000:] + r.stderr[-6000:])
        raise RuntimeError("nvcc failed building libb200md.so")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
