#!/bin/bash
tag=r02
out=gpurun_out; mkdir -p $out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -4 $out/smoke_$tag.log
if [ $rc -eq 124 ]; then echo "ABORT: smoke hung"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_$tag.log
timeout 500 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-600 $out/bench_$tag.json
timeout 300 python bench.py --impl reference > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err; echo "ref rc=$?"; cut -c1-400 $out/bench_ref_$tag.json
timeout 300 python bench.py --workload hipct256 --steps 300 --no-side-legs > $out/bench_hipct256_$tag.json 2>> $out/bench_$tag.err; echo "hipct256 rc=$?"
timeout 200 python bench.py --workload neuron1024_nb4 > $out/bench_nb4_$tag.json 2>> $out/bench_$tag.err; echo "nb4 rc=$?"
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-legs"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_$tag.csv $SHORT > $out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_fit_kernel -s 4 -c 1 -f -o $out/prof_fit_$tag $SHORT > $out/ncu_fit_$tag.log 2>&1
echo "ncu fit rc=$?"
LW="python bench.py --workload neuron1024_nb4 --steps 2 --warmup 3"
$LW > $out/plain_lw_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lw_gemm_kernel -s 27 -c 1 -f -o $out/prof_lw_bwd_$tag $LW > $out/ncu_lw_bwd_$tag.log 2>&1
echo "ncu lw bwd rc=$?"
