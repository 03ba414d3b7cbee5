#!/bin/bash
P='import sys,json; d=json.loads(sys.stdin.read()); print("value",d["value"]/1e9,"b2b",d["back_to_back"]["value"]/1e9,"kernel_ms",d["roofline"]["kernel_ms"],"e2e",d["e2e"]["value"]/1e9)'
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02l.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 gpurun_out/smoke_r02l.log
if [ $rc -ne 0 ]; then echo "ABORT: smoke failed"; tail -5 gpurun_out/smoke_r02l.log; exit 1; fi
for w in vessel neuron128 config1; do
echo "== $w product (dz_NH in group A)"; timeout 120 python bench.py --workload $w --steps 400 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "$P"
echo "== $w dz_NH in group B"; timeout 120 python tools/exp_variant.py dzb "-DBRIEF_FIT_DZ_IN_A=0" -- bench.py --workload $w --steps 400 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "$P"
done
timeout 100 python tools/exp_variant.py timing "-DBRIEF_TC_TIMING" -- tools/tc_stage_timing.py 56 7 100000 4 2>&1 | tail -21
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/pytest_r02l.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r02l.log
