#!/bin/bash
# evidence pass on one GPU: all GPU tests, bench lines, launch list and ncu --set full of the
# tile kernels; $1 = tag
tag=${1:-r01k}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for w in "lj 100 double" "lj 100 mixed" "eam 80 double" "eam 80 mixed"; do
  set -- $w
  python tools/perf_probe.py $1 $2 100 $3 2>&1 | tail -14 > gpurun_out/${tag}_probe_$1_$3.txt
  grep -H "steps:" gpurun_out/${tag}_probe_$1_$3.txt
done
python bench.py --workload lj4m --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/${tag}_bench_lj4m.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/${tag}_bench_lj4m.json
python bench.py --steps 100 --warmup 20 > gpurun_out/${tag}_bench_lj32m.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/${tag}_bench_lj32m.json
python bench.py --steps 100 --warmup 20 --precision mixed --no-cpu-baseline > gpurun_out/${tag}_bench_lj32m_mixed.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/${tag}_bench_lj32m_mixed.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_lj4m.csv python tools/perf_probe.py lj 100 40 double > /dev/null 2>&1
for p in double mixed; do
ncu --set full --clock-control none --import-source on -k regex:k_tile_lj -s 10 -c 1 -o gpurun_out/${tag}_full_k_tile_lj_$p -f python tools/perf_probe.py lj 100 25 $p > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_tile_build -s 1 -c 1 -o gpurun_out/${tag}_full_k_tile_build -f python tools/perf_probe.py lj 100 25 double > /dev/null 2>&1
du -sh gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_tile_lj_fx -s 10 -c 1 -o gpurun_out/${tag}_full_k_tile_lj_fx -f python tools/perf_probe.py lj 100 25 mixed > /dev/null 2>&1
python bench.py --workload eam2m --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/${tag}_bench_eam2m.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/${tag}_bench_eam2m.json
python bench.py --workload lj32k --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/${tag}_bench_lj32k.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/${tag}_bench_lj32k.json
python tools/skin_sweep.py > gpurun_out/${tag}_skin_sweep_lj4m.jsonl 2>> gpurun_out/bench.err; tail -3 gpurun_out/${tag}_skin_sweep_lj4m.jsonl
