#!/bin/bash
# eam-related GPU tests + the 2 M-atom eam probe (flat kernels) and the 32 k one (tile kernels)
mkdir -p gpurun_out
out=gpurun_out/${1:-eamq}_probe.txt
: > $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_eam2.py tests/test_gpu_tile_list.py tests/test_gpu_subdomains.py tests/test_gpu_mixed.py -m gpu -x -q -k "eam or mixed" 2>&1 | tail -6 >> $out
probe() { label=$1; shift; echo "== $label" >> $out; env "$@" 2>&1 | grep -E "steps:|pair |neigh|initial|clear|comm|rror" >> $out; }
probe "eam2m double" timeout 300 python tools/perf_probe.py eam 80 100 double
probe "eam2m mixed" timeout 300 python tools/perf_probe.py eam 80 100 mixed
probe "eam256k double" timeout 300 python tools/perf_probe.py eam 40 100 double
cat $out
