#!/bin/bash
# tile-size sweep (LJ 4M, one GPU); optional ncu captures of the tile kernels with "ncu" as $1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed.py -x -q 2>&1 | tail -5
for t in ${TILES:-8,4,4 8,8,4 4,4,4}; do
  for p in double mixed; do
    echo "== tile $t $p"
    B200_TILE=$t timeout 300 python tools/perf_probe.py lj 100 100 $p 2>&1 | grep -E "steps:|pair |neigh_build|rror"
  done
done
if [ "$1" == "ncu" ]; then
tag=${2:-r01e}
ncu --set full --clock-control none --import-source on -k regex:k_tile_lj -s 10 -c 1 -o gpurun_out/${tag}_full_k_tile_lj -f python tools/perf_probe.py lj 100 25 double > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_lj -s 10 -c 1 -o gpurun_out/${tag}_full_k_tile_lj_mixed -f python tools/perf_probe.py lj 100 25 mixed > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_build -s 1 -c 1 -o gpurun_out/${tag}_full_k_tile_build -f python tools/perf_probe.py lj 100 25 double > gpurun_out/ncu_b.log 2>&1
fi
