#!/bin/bash
# driver-protocol bench lines on N GPUs (lj32m strong scaling, eam 2 M atoms/GPU weak scaling); $1 = N, $2 = tag
N=${1:-8}; tag=${2:-r02ai}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py lj 16 100 > gpurun_out/${tag}_multi_check_lj_n$N.log 2>&1; tail -3 gpurun_out/${tag}_multi_check_lj_n$N.log
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tests/multi_rank_check.py eam 14 60 > gpurun_out/${tag}_multi_check_eam_n$N.log 2>&1; tail -3 gpurun_out/${tag}_multi_check_eam_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_lj32m_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "bench rc=$?"; cut -c1-260 gpurun_out/${tag}_bench_lj32m_n$N.json; tail -2 gpurun_out/${tag}_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 100 --warmup 20 --workload eam2m --no-cpu-baseline > gpurun_out/${tag}_bench_eam2m_weak_n$N.json 2>> gpurun_out/${tag}_bench_n$N.err
echo "bench eam rc=$?"; cut -c1-260 gpurun_out/${tag}_bench_eam2m_weak_n$N.json
