#!/bin/bash
# launch-shape / tile-shape sweep of the fused FP64 and mixed lj kernels at 4 M atoms; $1 = tag
tag=${1:-r02aj}
mkdir -p gpurun_out
out=gpurun_out/${tag}_lj2_sweep.txt
: > $out
run() { label=$1; shift; echo "== $label" >> $out; env "$@" timeout 300 python tools/perf_probe.py lj 100 60 ${PREC:-double} 2>&1 | grep -E "steps:|pair |neigh_build|rror" >> $out; }
run "default 8x8x4 352,2,2"
run "352,2,4" B200_LJ2=352,2,4
run "448,2,2" B200_LJ2=448,2,2
run "tile 8x4x4 320,3,2" B200_TILE=8,4,4 B200_LJ2=320,3,2
run "tile 8x4x4 256,4,2" B200_TILE=8,4,4 B200_LJ2=256,4,2
run "tile 8x8x2 320,3,2" B200_TILE=8,8,2 B200_LJ2=320,3,2
run "tile 16x4x4 352,2,2" B200_TILE=16,4,4
run "tile 4x8x4 256,4,2" B200_TILE=4,8,4 B200_LJ2=256,4,2
run "tile 8x6x4 352,2,2" B200_TILE=8,6,4
run "tile 10x8x4 448,2,2" B200_TILE=10,8,4 B200_LJ2=448,2,2
PREC=mixed
run "mixed default"
run "mixed tile 8x4x4" B200_TILE=8,4,4
run "mixed tile 16x4x4" B200_TILE=16,4,4
run "mixed tile 8x8x8" B200_TILE=8,8,8
cat $out
