#!/bin/bash
out=gpurun_out; mkdir -p $out
P='import sys,json; d=json.loads(sys.stdin.read()); print("value",d["value"]/1e9,"b2b",d["back_to_back"]["value"]/1e9,"kernel_ms",d["roofline"]["kernel_ms"],"dec",d["decompress"]["value"]/1e9, "deblock_ms", (d.get("deblock") or {}).get("ms"))'
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r02i.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 $out/smoke_r02i.log
if [ $rc -eq 124 ]; then echo "ABORT: smoke hung"; exit 1; fi
echo "== product"; timeout 200 python bench.py --steps 400 --no-cpu-baseline 2>/dev/null | python -c "$P"
echo "== pad skip (4-column groups, forward epilogues)"; timeout 200 python tools/exp_variant.py padskip "-DBRIEF_PAD_SKIP=1" -- bench.py --steps 400 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "$P"
timeout 200 python bench.py --workload neuron1024_nb4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nb4: value',d['value']/1e6,'M/s ms',d['ms_per_step'],'TF',d['roofline']['achieved'],'dec',d['decompress']['value']/1e9)"
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_deblock.py -m gpu -q -k "wide_networks or deblock" --timeout 150 --timeout-method=thread 2>&1 | tail -3
