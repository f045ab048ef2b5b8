#!/bin/bash
# sweep lanes-per-atom (B200_TPA) for the pair kernels, double and mixed
python -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed.py -q 2>&1 | tail -3
for t in 1 2 4 8; do
  for prec in double mixed; do
    echo "== TPA=$t $prec LJ 4M"; B200_TPA=$t python tools/perf_probe.py lj 100 100 $prec 2>&1 | grep -E "steps:|  pair |neigh_build"
  done
done
for t in 1 4 8; do
  for prec in double mixed; do
    echo "== TPA=$t $prec EAM 2M"; B200_TPA=$t python tools/perf_probe.py eam 80 100 $prec 2>&1 | grep -E "steps:|  pair "
  done
done
B200_TPA=8 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed.py -q 2>&1 | tail -3
