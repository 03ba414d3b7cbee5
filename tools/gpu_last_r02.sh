#!/bin/bash
# Last GPU call of round 2 (about 3 GPU-minutes were left): the whole -m gpu suite with the new sampler / schedule / half
# tests first, then smoke(), then the default bench line with whatever time remains.
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 125 python -m pytest tests/test_gpu_cubes.py tests -m gpu -q --tb=short --timeout=50 -p no:cacheprovider > $out/pytest_r02_last.log 2>&1
echo "pytest rc=$? t=$SECONDS"
tail -n 30 $out/pytest_r02_last.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r02_last.log 2>&1
echo "smoke rc=$? t=$SECONDS"
tail -n 4 $out/smoke_r02_last.log
left=$((165 - SECONDS))
if [ $left -gt 20 ]; then
  timeout $left python bench.py > $out/bench_r02_last.json 2> $out/bench_r02_last.err
  echo "bench rc=$? t=$SECONDS"
  head -c 600 $out/bench_r02_last.json
fi
