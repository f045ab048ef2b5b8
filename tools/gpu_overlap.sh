#!/bin/bash
# 2-GPU check of the halo/compute overlap: parity with overlap on, then bench on/off
N=${1:-2}
mkdir -p gpurun_out
for k in "lj 24 100" "lj 12 100" "eam 8 60"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py $k 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -5
done
for ov in 1 0; do
for wl in lj4m lj32m; do
B200_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 20 --workload $wl > gpurun_out/ov${ov}_bench_${wl}_n$N.json 2> gpurun_out/ov_bench.err
echo "overlap=$ov $wl rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/ov${ov}_bench_${wl}_n$N.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("halo_overlap"), {k:v["ms"] for k,v in d["phases"].items()})
PY
done
done
