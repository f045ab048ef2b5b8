#!/bin/bash
# quick single-GPU pass: all GPU tests + per-phase probes (double + mixed, LJ + EAM)
# usage: gpu_check.sh [probe-only]
mkdir -p gpurun_out
if [ "$1" != "probe-only" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -15
fi
for w in "lj 100 double" "lj 100 mixed" "eam 80 double" "eam 80 mixed"; do
  set -- $w
  timeout 300 python tools/perf_probe.py $1 $2 100 $3 2>&1 | grep -E "steps:|pair |neigh|initial|final|clear|comm|stats|rror"
done
