#!/bin/bash
# evidence pass on one GPU: launch lists + ncu --set full of every kernel class (reports kept
# small: one launch per kernel, no source import -- gpurun_out/ is capped at 64 MiB)
tag=${1:-r01d}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in "lj 100 double" "lj 100 mixed" "eam 80 double" "eam 80 mixed"; do
  set -- $w
  python tools/perf_probe.py $1 $2 100 $3 2>&1 | tail -14 > gpurun_out/${tag}_probe_$1_$3.txt
  grep -H "steps:" gpurun_out/${tag}_probe_$1_$3.txt
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_lj4m.csv python tools/perf_probe.py lj 100 40 double > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_eam2m.csv python tools/perf_probe.py eam 80 40 double > /dev/null 2>&1
for k in k_pair_lj k_nve_initial k_nve_final_initial k_build_half k_pbc_bin k_permute_owned k_unpack_forward k_pack_reverse; do
  ncu --set full --clock-control none -k regex:"^$k" -s 3 -c 1 -o gpurun_out/${tag}_full_$k -f python tools/perf_probe.py lj 100 25 double > /dev/null 2>&1
done
for k in k_pair_lj_mixed k_merge_ff; do
  ncu --set full --clock-control none -k regex:"^$k" -s 3 -c 1 -o gpurun_out/${tag}_full_$k -f python tools/perf_probe.py lj 100 25 mixed > /dev/null 2>&1
done
for k in k_eam_rho k_eam_force k_eam_embed; do
  ncu --set full --clock-control none -k regex:"^$k" -s 3 -c 1 -o gpurun_out/${tag}_full_$k -f python tools/perf_probe.py eam 80 25 double > /dev/null 2>&1
done
du -sh gpurun_out
