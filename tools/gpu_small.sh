#!/bin/bash
# small systems (the reference's own 32 k-atom bench inputs): fused integrator vs CUDA graph; $1 = tag
tag=${1:-r02ak}
mkdir -p gpurun_out
out=gpurun_out/${tag}_small.txt
: > $out
run() { label=$1; shift; echo "== $label" >> $out; env "$@" timeout 300 python tools/perf_probe.py $KIND 20 ${STEPS:-1000} double 2>&1 | grep -E "steps:|rror|launches" | cut -c1-200 >> $out; }
KIND=lj
run "lj32k default (graph)"
run "lj32k fused" B200_FUSE_MIN=0
run "lj32k no graph no fuse" B200_GRAPH=0
KIND=eam STEPS=400
run "eam32k default (eam2, unfused)"
run "eam32k fused" B200_FUSE_MIN=0
run "eam32k flat" B200_EAM2=0
cd lammps_b200/lammps_pkg/bench_inputs
for v in "" "B200_FUSE_MIN=0"; do
echo "== lmp_b200 in.lj $v" >> ../../../$out; env $v ../lmp_b200 -sf b200 -in in.lj 2>&1 | grep -E "Loop time" >> ../../../$out
echo "== lmp_b200 in.eam $v" >> ../../../$out; env $v ../lmp_b200 -sf b200 -in in.eam 2>&1 | grep -E "Loop time" >> ../../../$out
done
cd ../../..
cat $out
