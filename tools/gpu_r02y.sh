#!/bin/bash
# r02y: (1) parity subset, (2) k_tile_lj2 staged-layout A/B ({x,y}+z vs 24-byte records), (3) eam baseline split
mkdir -p gpurun_out
out=gpurun_out/r02y_probe.txt
: > $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile_list.py tests/test_gpu_fused_nve.py -m gpu -x -q 2>&1 | tail -5 >> $out
probe() { # label, env..., -- args
  label=$1; shift
  echo "== $label" >> $out
  env "$@" 2>&1 | grep -E "steps:|pair |neigh|initial|final|clear|comm|rror" >> $out
}
for rep in 1 2; do
probe "lj4m double base rep$rep" B200_LIBPATH=$PWD/lammps_b200/libb200md_base.so timeout 300 python tools/perf_probe.py lj 100 60 double
probe "lj4m double xy+z rep$rep" timeout 300 python tools/perf_probe.py lj 100 60 double
done
probe "eam2m double flat" timeout 300 python tools/perf_probe.py eam 80 100 double
probe "eam2m double tile" B200_LIST=tile timeout 300 python tools/perf_probe.py eam 80 100 double
probe "eam2m mixed flat" timeout 300 python tools/perf_probe.py eam 80 100 mixed
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_eam -s 40 -c 60 --csv --log-file gpurun_out/r02y_launches_eam.csv python tools/perf_probe.py eam 80 30 double > /dev/null 2>&1
cat $out
