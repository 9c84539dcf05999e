#!/bin/bash
# multi-GPU quick session: strong-scaling sharded bench at N ranks, no side legs
tag=${1:-mg}; N=${2:-8}; shift; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 2 \
  --no-cpu-baseline --no-other-configs --no-e2e "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err; echo "rc=$?" >> gpurun_out/${tag}.err
grep "step \|parity\|Error\|error\|rc=" gpurun_out/${tag}.err | tail -8
