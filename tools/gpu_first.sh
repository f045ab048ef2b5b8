#!/bin/bash
# first GPU pass: parity tests, probes, bench, ncu launch list + one full capture
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/perf_probe.py lj 100 100 > gpurun_out/probe_lj4m.log 2>&1
python tools/perf_probe.py eam 80 100 > gpurun_out/probe_eam2m.log 2>&1
python bench.py --workload lj4m --steps 100 --warmup 20 > gpurun_out/bench_lj4m.json 2> gpurun_out/bench_lj4m.err
python bench.py --steps 100 --warmup 20 > gpurun_out/bench_lj32m.json 2> gpurun_out/bench_lj32m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_lj4m.csv python tools/perf_probe.py lj 100 25 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_lj -s 10 -c 2 -o gpurun_out/prof_pair_lj python tools/perf_probe.py lj 100 25 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_half -s 1 -c 1 -o gpurun_out/prof_build python tools/perf_probe.py lj 100 25 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/probe_lj4m.log; cat gpurun_out/probe_eam2m.log; cat gpurun_out/bench_lj4m.json; cat gpurun_out/bench_lj32m.json; tail -5 gpurun_out/bench_lj32m.err
