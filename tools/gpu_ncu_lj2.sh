#!/bin/bash
# ncu --set full of the lj/cut tile kernels on 4 M atoms: $1 = tag
tag=${1:-r02b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B200_LJ2=352,2,2 ncu --set full --clock-control none --import-source on -k regex:k_tile_lj2 -s 10 -c 1 -o gpurun_out/${tag}_full_lj2 -f python tools/perf_probe.py lj 100 25 double > /dev/null 2>&1
B200_LJ2=0 ncu --set full --clock-control none --import-source on -k regex:k_tile_lj -s 10 -c 1 -o gpurun_out/${tag}_full_lj1 -f python tools/perf_probe.py lj 100 25 double > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
