#!/bin/bash
# N-GPU session: strong-scaling bench with the e2e legs, and the box's aggregate D2H bandwidth
tag=${1:-mg}; N=${2:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 \
  > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?" >> gpurun_out/${tag}_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/d2h_probe.py > gpurun_out/${tag}_d2h.json 2> gpurun_out/${tag}_d2h.err
grep "step \|parity\|e2e pass\|rc=" gpurun_out/${tag}_bench.err | tail -12; cat gpurun_out/${tag}_d2h.json
