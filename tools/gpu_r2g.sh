#!/bin/bash
P='import sys,json; d=json.loads(sys.stdin.read()); print("value",d["value"]/1e9,"b2b",d["back_to_back"]["value"]/1e9,"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"])'
for w in neuron128 config1; do
  echo "== $w product (two issuers for F>=48)"; timeout 200 python bench.py --workload $w --steps 300 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "$P"
  echo "== $w two issuers for F>=32"; timeout 200 python tools/exp_variant.py issuer32 "-DBRIEF_FIT_TWO_ISSUERS_MIN_F=32" -- bench.py --workload $w --steps 300 --no-cpu-baseline --no-side-legs 2>/dev/null | python -c "$P"
done
bash tools/gpu_r2_prof.sh r02
