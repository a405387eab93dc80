#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-b}
N=${2:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
fi
tail -c 3000 gpurun_out/bench_${TAG}*.json; tail -5 gpurun_out/bench_${TAG}*.err
