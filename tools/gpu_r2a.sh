#!/bin/bash
# round-2 first GPU call: all GPU tests, smoke, default bench, stage timing of the fit kernel (baseline before kernel work)
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -5 $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json; tail -5 $out/bench_$tag.err
python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/tc_stage_timing.py 56 7 100000 4 > $out/timing_$tag.txt 2>&1; echo "timing rc=$?"; cat $out/timing_$tag.txt
