#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02ae_probe.txt
: > $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $out
probe() { label=$1; shift; echo "== $label" >> $out; env "$@" 2>&1 | grep -E "steps:|pair |neigh|rror" >> $out; }
probe "lj4m double" timeout 300 python tools/perf_probe.py lj 100 60 double
probe "lj4m mixed" timeout 300 python tools/perf_probe.py lj 100 60 mixed
probe "eam2m double" timeout 300 python tools/perf_probe.py eam 80 100 double
probe "eam32k double" timeout 300 python tools/perf_probe.py eam 20 100 double
cat $out
