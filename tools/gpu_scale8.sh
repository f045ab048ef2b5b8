#!/bin/bash
# 8-GPU pass: parity at 2x2x2, then the strong-scaled LJ 32M and weak-scaled EAM points
N=${1:-8}
mkdir -p gpurun_out
for k in "lj 24 100" "eam 12 60"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py $k 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -6 | tee -a gpurun_out/multi_check_n$N.log
done
for wl in lj32m eam2m; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 20 --workload $wl > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
echo "rc=$?"; cut -c1-2500 gpurun_out/bench_${wl}_n$N.json; tail -3 gpurun_out/bench_${wl}_n$N.err
done
