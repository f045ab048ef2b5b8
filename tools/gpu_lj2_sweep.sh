#!/bin/bash
# k_tile_lj2 launch-shape sweep on 4 M LJ atoms (one GPU): B200_LJ2=threads,minb,ilp [+ B200_TILE]
mkdir -p gpurun_out
out=gpurun_out/${1:-r02a}_lj2_sweep.txt
: > $out
run() {  # label, env...
  label=$1; shift
  echo "== $label" >> $out
  env "$@" timeout 300 python tools/perf_probe.py lj 100 60 double 2>&1 | grep -E "steps:|pair |neigh_build|rror" >> $out
}
if [ -z "$2" ]; then
run "old k_tile_lj" B200_LJ2=0
run "352,2,2" B200_LJ2=352,2,2
run "352,2,4" B200_LJ2=352,2,4
run "448,2,2" B200_LJ2=448,2,2
run "448,2,4" B200_LJ2=448,2,4
run "tile 8x4x4 320,3,2" B200_TILE=8,4,4 B200_LJ2=320,3,2
run "tile 8x4x4 256,4,2" B200_TILE=8,4,4 B200_LJ2=256,4,2
else
shift
for cfg in "$@"; do run "$cfg" B200_LJ2=$cfg; done
fi
cat $out
