#!/bin/bash
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -4 $out/smoke_$tag.log
if [ $rc -eq 124 ]; then echo "ABORT: smoke hung"; exit 1; fi
timeout 700 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -15 $out/pytest_$tag.log
timeout 500 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json | cut -c1-1500; tail -5 $out/bench_$tag.err
timeout 300 python bench.py --workload neuron1024_nb4 > $out/bench_nb4_$tag.json 2> $out/bench_nb4_$tag.err; echo "nb4 rc=$?"; cat $out/bench_nb4_$tag.json | cut -c1-3000; tail -5 $out/bench_nb4_$tag.err
