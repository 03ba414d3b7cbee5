#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench line, the ncu launch list and one full capture of the fit and
# decompress kernels.  Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01'
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json
python bench.py --impl reference > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err; echo "ref rc=$?"; cat $out/bench_ref_$tag.json
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $SHORT > $out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_fit_kernel -s 4 -c 1 -f -o $out/prof_fit_$tag $SHORT > $out/ncu_fit_$tag.log 2>&1
echo "ncu fit rc=$?"
$SHORT > $out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_eval_kernel -s 1 -c 1 -f -o $out/prof_eval_$tag $SHORT > $out/ncu_eval_$tag.log 2>&1
echo "ncu eval rc=$?"
WIDE="python bench.py --workload hipct256 --steps 3 --warmup 3 --no-cpu-baseline"
$WIDE > $out/plain_wide_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_fit_wide_kernel -s 4 -c 1 -f -o $out/prof_fit_wide_$tag $WIDE > $out/ncu_fit_wide_$tag.log 2>&1
echo "ncu wide fit rc=$?"
python bench.py --workload hipct256 > $out/bench_hipct256_$tag.json 2>> $out/bench_$tag.err; echo "hipct bench rc=$?"; cat $out/bench_hipct256_$tag.json
