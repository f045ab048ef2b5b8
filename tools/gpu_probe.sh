#!/bin/bash
# quick single-GPU pass: parity, probes (double + mixed), ncu full capture of the pair kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed.py -q 2>&1 | tail -5
python tools/perf_probe.py lj 100 100 double 2>&1 | grep -E "steps:|pair |neigh |initial|final|clear"
python tools/perf_probe.py lj 100 100 mixed 2>&1 | grep -E "steps:|pair |neigh |initial|final|clear"
python tools/perf_probe.py eam 80 100 double 2>&1 | grep -E "steps:|pair "
python tools/perf_probe.py eam 80 100 mixed 2>&1 | grep -E "steps:|pair "
ncu --set full --clock-control none --import-source on -k regex:k_pair_lj -s 10 -c 1 -o gpurun_out/prof_pair_lj_${1:-x} -f python tools/perf_probe.py lj 100 25 double > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_lj_mixed -s 10 -c 1 -o gpurun_out/prof_pair_lj_mixed_${1:-x} -f python tools/perf_probe.py lj 100 25 mixed > gpurun_out/ncu_b.log 2>&1
