#!/bin/bash
P='import sys,json; d=json.loads(sys.stdin.read()); print("value",d["value"]/1e9,"b2b",d["back_to_back"]["value"]/1e9,"kernel_ms",d["roofline"]["kernel_ms"],"e2e",d["e2e"]["value"]/1e9)'
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02j.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 gpurun_out/smoke_r02j.log
if [ $rc -eq 124 ]; then echo "ABORT: smoke hung"; exit 1; fi
for w in vessel neuron128; do
echo "== $w product (two samplers)"; timeout 200 python bench.py --workload $w --steps 400 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "$P"
echo "== $w one sampler"; timeout 200 python tools/exp_variant.py onesampler "-DBRIEF_FIT_TWO_SAMPLERS=0" -- bench.py --workload $w --steps 400 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "$P"
done
timeout 100 python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/tc_stage_timing.py 56 7 100000 4 --flush 2>&1 | grep -E "us per step|wait sampler|TOTAL|kernel total"
timeout 100 python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/tc_stage_timing.py 56 7 100000 4 2>&1 | grep -E "us per step|wait sampler|TOTAL|kernel total"
