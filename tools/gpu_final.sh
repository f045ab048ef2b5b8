#!/bin/bash
# final evidence of a round on one GPU; $1 = tag
#  1. ncu launch list of the driver's bench command (side measurements off)
#  2. ncu --set full of the dominant kernel (k_tile_lj2, FP64, fused fix nve)
#  3. all GPU tests, smoke, the driver-protocol bench line
tag=${1:-r02bc}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_bench_lj32m.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/${tag}_launches_bench_lj32m.csv > gpurun_out/${tag}_launches_bench_lj32m.txt 2>&1; head -8 gpurun_out/${tag}_launches_bench_lj32m.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_lj2 -s 10 -c 1 -o gpurun_out/${tag}_full_lj2_double -f python tools/perf_probe.py lj 100 25 double > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/${tag}_full_lj2_double.ncu-rep > gpurun_out/${tag}_ncu_full_k_tile_lj2_double.txt 2>&1; head -12 gpurun_out/${tag}_ncu_full_k_tile_lj2_double.txt
python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_lj32m_n1_driver_protocol.json 2> gpurun_out/${tag}_bench.err
cut -c1-400 gpurun_out/${tag}_bench_lj32m_n1_driver_protocol.json; tail -2 gpurun_out/${tag}_bench.err
