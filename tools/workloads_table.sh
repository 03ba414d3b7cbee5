#!/bin/bash
# bench.py on every block geometry of SURVEY 8(d) that fits one GPU -> gpurun_out/workloads_<tag>.txt
tag=${1:-r01}
out=gpurun_out/workloads_$tag.txt
echo "# bench.py --workload W --no-cpu-baseline on one B200; fit = coord-samples/s, fwd+bwd+Adamax" > $out
echo "# workload      blocks/GPU  f    prec  fit samples/s  e2e samples/s  ms/step  fit-kernel ms  algorithmic TF  frac of peak   decompress vox/s" >> $out
for w in vessel vessel64 config1 neuron128 hipct256; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 300 --warmup 20 2>/dev/null | python -c "
import json, sys
l = json.loads(sys.stdin.read()); c = l['config']; r = l['roofline']
print(f\"{'$w':14s} {c['blocks_per_gpu']:9d} {c.get('features', 0):4d}  {l['dtype']:4s} {l['value']:14.3e} {l['e2e']['value']:14.3e} {l['ms_per_step']:8.3f} {r['kernel_ms']:14.3f} {r['achieved']:15.1f} {r['frac']:13.4f}   {l['decompress']['value']:.3e}\")
" >> $out
done
cat $out
