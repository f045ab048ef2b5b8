#!/bin/bash
# the LAMMPS package driving N GPUs from one process (`-pk b200 gpus N`) on the 32 M-atom LJ melt,
# next to bench.py on the same N GPUs; $1 = N, $2 = tag
N=${1:-2}; tag=${2:-r02al}
mkdir -p gpurun_out
out=gpurun_out/${tag}_lmp_b200_gpus$N.txt
: > $out
cd lammps_b200/lammps_pkg/bench_inputs
echo "== lmp_b200 -sf b200 -pk b200 gpus $N -var x 10 -var y 10 -var z 10 -in in.lj" >> ../../../$out
timeout 900 ../lmp_b200 -sf b200 -pk b200 gpus $N -var x 10 -var y 10 -var z 10 -in in.lj 2>&1 | grep -E "Loop time|B200 package|Step|^ +[0-9]+ |Neighbor list builds|Total # of neighbors|atoms" | head -20 >> ../../../$out
echo "== lmp_b200 -sf b200 -pk b200 gpus 1 (same input, one GPU)" >> ../../../$out
timeout 900 ../lmp_b200 -sf b200 -var x 10 -var y 10 -var z 10 -in in.lj 2>&1 | grep -E "Loop time|Neighbor list builds|Total # of neighbors" >> ../../../$out
cd ../../..
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench.py --gpus', d['n_gpus'], 'steps 100:', d['value'], 'atom-steps/s', d['ms_per_step'], 'ms/step')" >> $out
cat $out
