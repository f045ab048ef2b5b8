#!/bin/bash
# mixed-precision tile kernel variants (LJ 4M, one GPU)
python -m pytest tests/test_gpu_mixed.py tests/test_gpu_tile_list.py -x -q 2>&1 | tail -3
for lib in libb200md.so libb200md_fx3.so; do
for t in 8,8,4 8,4,4; do
  echo "== $lib tile $t mixed fx"
  B200_LIBPATH=$PWD/lammps_b200/$lib B200_TILE=$t timeout 300 python tools/perf_probe.py lj 100 100 mixed 2>&1 | grep -E "steps:|pair |rror|thermo"
done
done
echo "== FP64-staged mixed"; B200_MIXED_FX=0 python tools/perf_probe.py lj 100 100 mixed 2>&1 | grep -E "steps:|pair |thermo"
echo "== double"; python tools/perf_probe.py lj 100 100 double 2>&1 | grep -E "steps:|pair |thermo"
