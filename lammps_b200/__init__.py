"""lammps_b200 -- B200-native short-range MD hot path (lj/cut, eam, nve, half Verlet lists).

The compute path is hand-written CUDA for sm_100a behind a C ABI (include/b200_md.h,
lammps_b200/csrc).  This Python package is the thin host mirror used by tests and bench.py.
"""
__version__ = "0.1.0"
