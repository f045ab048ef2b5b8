"""CPU tests of the drop-in boundary: libb200md.so loads, exports every symbol include/b200_md.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from lammps_b200 import engine

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "b200_md.h").read_text()
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200_md.h but not exported"
    assert sorted(engine.EXPORTS) == names


def test_header_cites_reference_interfaces():
    text = (ROOT / "include" / "b200_md.h").read_text()
    for cite in ("fix_nve.cpp", "comm_brick.cpp", "neighbor.cpp", "pair_lj_cut.cpp", "pair_eam.cpp",
                 "verlet.cpp", "GPU/pair_lj_cut_gpu.cpp"):
        assert cite in text


def test_no_cpu_fallback():
    lib = engine.load_library()
    if lib.b200_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(engine.B200Error):
        engine.Engine(0)
    h = C.c_void_p()
    assert lib.b200_create(C.byref(h), 0, 0) != 0


def test_product_never_imports_oracle():
    for p in (ROOT / "lammps_b200").rglob("*.py"):
        t = p.read_text()
        assert "import oracle" not in t and "from oracle" not in t, p
    for p in (ROOT / "lammps_b200" / "csrc").glob("*"):
        assert "oracle" not in p.read_text()


def test_stats_struct_layout_matches_the_header(tmp_path):
    """engine.Stats (ctypes) and the LAMMPS package both mirror `b200_stats`; a C program
    compiled against the header says what the library really writes (size and field offsets)"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    fields = [n for n, _ in engine.Stats._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "b200_md.h"}"',
             'int main(void) {', '  printf("%zu\\n", sizeof(b200_stats));']
    lines += [f'  printf("%zu\\n", offsetof(b200_stats, {n}));' for n in fields]
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert out[0] == C.sizeof(engine.Stats)
    assert out[1:] == [getattr(engine.Stats, n).offset for n in fields]
