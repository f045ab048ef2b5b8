#!/bin/bash
# compare library variants x tile sizes (LJ 4M, one GPU)
for lib in ${LIBS:-libb200md.so libb200md_cpa.so libb200md_minb3.so}; do
for t in 8,4,4 8,8,4; do
  for p in double mixed; do
    echo "== $lib tile $t $p"
    B200_LIBPATH=$PWD/lammps_b200/$lib B200_TILE=$t timeout 300 python tools/perf_probe.py lj 100 100 $p 2>&1 | grep -E "steps:|pair |neigh_build|rror"
  done
done
done
