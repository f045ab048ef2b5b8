#!/bin/bash
# ncu --set full of the tile kernels (LJ 4M, one GPU); $1 = tag, $2 = precision list
tag=${1:-r01g}
mkdir -p gpurun_out
for p in ${2:-double mixed}; do
ncu --set full --clock-control none --import-source on -k regex:k_tile_lj -s 10 -c 1 -o gpurun_out/${tag}_full_k_tile_lj_$p -f python tools/perf_probe.py lj 100 25 $p > gpurun_out/ncu_$p.log 2>&1
done
