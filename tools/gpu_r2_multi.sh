#!/bin/bash
# named shapes of BASELINE configs 3-5 on N GPUs of one box (strong scaling: ONE volume sharded by LPT) + the default line
N=${1:-8}
tag=${2:-r02}
out=gpurun_out
mkdir -p $out
run() {  # name, extra args...
  name=$1; shift
  if [ "$N" -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $out/${name}_n${N}_$tag.json 2> $out/${name}_n${N}_$tag.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $out/${name}_n${N}_$tag.json 2> $out/${name}_n${N}_$tag.err
  fi
  echo "$name N=$N rc=$?"; cat $out/${name}_n${N}_$tag.json | cut -c1-1800; tail -3 $out/${name}_n${N}_$tag.err | cut -c1-400
}
for w in ${WORKLOADS:-neuron1024 hipct2048_equal decomp4096 hipct2048 neuron1024_nb4}; do
  run bench_$w --workload $w --reproducible
done
if [ "${DEFAULT:-1}" = "1" ]; then run bench_default --steps 200 --warmup 20 --reproducible; fi
if [ "$N" -eq 2 ]; then timeout 300 python -m pytest tests/test_sharding.py -m gpu -q > $out/pytest_nccl_$tag.log 2>&1; echo "nccl test rc=$?"; tail -3 $out/pytest_nccl_$tag.log; fi
