#!/bin/bash
# multi-GPU pass: single-GPU parity regression, then N-rank parity and scaling probes
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_multi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_multi.log
tail -15 gpurun_out/pytest_gpu_multi.log
for kind in lj eam; do
  NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py $kind 12 100 > gpurun_out/multi_check_${kind}_n$N.log 2>&1
  echo "rc=$?" >> gpurun_out/multi_check_${kind}_n$N.log
  tail -12 gpurun_out/multi_check_${kind}_n$N.log
done
for wl in lj4m lj32m; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 20 --workload $wl > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
echo "rc=$?"; cat gpurun_out/bench_${wl}_n$N.json; tail -5 gpurun_out/bench_${wl}_n$N.err
done
