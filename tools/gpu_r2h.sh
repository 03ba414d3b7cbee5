#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_r02h.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 $out/smoke_r02h.log
if [ $rc -eq 124 ]; then echo "ABORT: smoke hung"; exit 1; fi
timeout 300 python bench.py --steps 300 --no-cpu-baseline > $out/bench_r02h.json 2> $out/bench_r02h.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02h.json"))
print("value",d["value"]/1e9,"b2b",d["back_to_back"]["value"]/1e9,"e2e",d["e2e"]["value"]/1e9,"kernel_ms",d["roofline"]["kernel_ms"])
print("decompress",d["decompress"]["value"]/1e9, "deblock", d.get("deblock"))
print({k:(round(v.get("value",0)/1e9,3),round(v.get("decompress_voxels_per_s",0)/1e9,2)) for k,v in d["workloads"].items()})
print("strong", d["strong_scaling"]["value"]/1e9)
PY
timeout 200 python bench.py --workload neuron1024_nb4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nb4 fill-wave slicing: value',d['value']/1e6,'M/s ms',d['ms_per_step'],'TF',d['roofline']['achieved'],'dec',d['decompress']['value']/1e9)"
LW="python bench.py --workload neuron1024_nb4 --steps 2 --warmup 3"
$LW > $out/plain_lw_r02.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lw_|opt_kernel" -s 66 -c 44 --csv --log-file $out/launches_lw_r02.csv $LW > $out/ncu_launch_lw_r02.log 2>&1
echo "lw launch list rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > $out/pytest_r02h.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_r02h.log
