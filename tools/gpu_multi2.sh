#!/bin/bash
# round-2 multi-GPU pass on N GPUs: the whole GPU suite (multi-rank tests included), the N-rank
# parity check for lj and eam, and the driver-protocol bench line; $1 = N, $2 = tag
N=${1:-2}; tag=${2:-r02ah}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu_n$N.log
tail -6 gpurun_out/${tag}_pytest_gpu_n$N.log
for kind in lj eam; do
  NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py $kind 12 100 > gpurun_out/${tag}_multi_check_${kind}_n$N.log 2>&1
  echo "rc=$?" >> gpurun_out/${tag}_multi_check_${kind}_n$N.log
  tail -6 gpurun_out/${tag}_multi_check_${kind}_n$N.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_lj32m_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "bench rc=$?"; cut -c1-260 gpurun_out/${tag}_bench_lj32m_n$N.json; tail -3 gpurun_out/${tag}_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 100 --warmup 20 --workload eam2m --no-cpu-baseline > gpurun_out/${tag}_bench_eam2m_weak_n$N.json 2>> gpurun_out/${tag}_bench_n$N.err
echo "bench eam rc=$?"; cut -c1-260 gpurun_out/${tag}_bench_eam2m_weak_n$N.json
