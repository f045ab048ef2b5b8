#!/bin/bash
# round-2 evidence on one GPU: ncu --set full of the fused lj/cut tile kernels (FP64 + mixed), launch
# list of a bench run, the driver-protocol bench line (reference arm first); $1 = tag
tag=${1:-r02ag}
mkdir -p gpurun_out
for p in double mixed; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_lj2 -s 10 -c 1 -o gpurun_out/${tag}_full_lj2_$p -f python tools/perf_probe.py lj 100 25 $p > /dev/null 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_lj4m.csv python bench.py --workload lj4m --steps 20 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_lj32m_n1.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/${tag}_bench_lj32m_n1.json; tail -3 gpurun_out/bench.err
ls -la gpurun_out | grep $tag
