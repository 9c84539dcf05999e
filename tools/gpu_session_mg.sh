#!/bin/bash
# multi-GPU session: strong-scaling bench at N ranks, sharded and replicated build
tag=${1:-mg}; N=${2:-2}
mkdir -p gpurun_out
F="--no-cpu-baseline --no-other-configs"
run() { # name, extra flags
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 2 $F $2 \
    > gpurun_out/${tag}_$1.json 2> gpurun_out/${tag}_$1.err; echo "rc=$?" >> gpurun_out/${tag}_$1.err
  grep "step \|parity\|e2e pass\|Error\|error" gpurun_out/${tag}_$1.err | tail -6
}
run sharded ""
run replicated "--build replicated --no-e2e"
