#!/bin/bash
# ncu evidence for round 2: launch lists (per-launch device time) and one full capture per dominant kernel
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-legs"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_$tag.csv $SHORT > $out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_fit_kernel -s 4 -c 1 -f -o $out/prof_fit_$tag $SHORT > $out/ncu_fit_$tag.log 2>&1
echo "ncu fit rc=$?"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_eval_kernel -s 1 -c 1 -f -o $out/prof_eval_$tag $SHORT > $out/ncu_eval_$tag.log 2>&1
echo "ncu eval rc=$?"
LW="python bench.py --workload neuron1024_nb4 --steps 2 --warmup 3"
$LW > $out/plain_lw_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv --log-file $out/launches_lw_$tag.csv $LW > $out/ncu_launch_lw_$tag.log 2>&1
echo "lw launch list rc=$?"
$LW > $out/plain_lw_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lw_gemm_kernel -s 22 -c 1 -f -o $out/prof_lw_gemm_$tag $LW > $out/ncu_lw_gemm_$tag.log 2>&1
echo "ncu lw gemm rc=$?"
$LW > $out/plain_lw_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lw_dw_kernel -s 15 -c 1 -f -o $out/prof_lw_dw_$tag $LW > $out/ncu_lw_dw_$tag.log 2>&1
echo "ncu lw dw rc=$?"
ls -la $out/*.ncu-rep | tail -6
