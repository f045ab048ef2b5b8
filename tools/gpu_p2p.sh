#!/bin/bash
# 2-GPU check of the peer-memory halo against the NCCL transport
N=${1:-2}
for k in "lj 14 100" "eam 10 100"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py $k 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^W1017\|^\[rank" | tail -5
done
for mode in p2p nccl; do
for wl in lj4m eam2m; do
echo "== $mode $wl"
B200_HALO=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 20 --workload $wl 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['ms']/v['calls']*1e3,1) for k,v in d['phases'].items()})"
done
done
