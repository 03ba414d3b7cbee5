#!/bin/bash
# smoke first (hang check), then tests in two stages (established paths / new layer-wise widths), bench, stage timing
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -6 $out/smoke_$tag.log
if [ $rc -eq 124 ]; then echo "ABORT: smoke hung"; exit 1; fi
LW='140 or 180 or 228 or 254 or 127-4'
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread -k "not ($LW)" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -15 $out/pytest_$tag.log
timeout 400 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread -k "$LW" > $out/pytest_lw_$tag.log 2>&1; echo "pytest lw rc=$?" | tee -a $out/pytest_lw_$tag.log
tail -40 $out/pytest_lw_$tag.log
timeout 500 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json; tail -5 $out/bench_$tag.err
timeout 100 python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/tc_stage_timing.py 56 7 100000 4 > $out/timing_$tag.txt 2>&1; echo "timing rc=$?"; cat $out/timing_$tag.txt
