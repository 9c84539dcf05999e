#!/bin/bash
# 2-GPU box: parity tests that touch the build paths + N=1 and N=2 strong bench lines (no side legs)
tag=${1:-q2}
mkdir -p gpurun_out
F="--no-e2e --no-cpu-baseline --no-other-configs"
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "not concurrent" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 4 --warmup 2 $F > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; echo "rc=$?" >> gpurun_out/${tag}_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 $F > gpurun_out/${tag}_n2.json 2> gpurun_out/${tag}_n2.err; echo "rc=$?" >> gpurun_out/${tag}_n2.err
tail -3 gpurun_out/${tag}_pytest.log; grep "step \|parity\|rc=" gpurun_out/${tag}_n1.err | tail -3 | cut -c1-220; grep "step \|parity\|rc=" gpurun_out/${tag}_n2.err | tail -3 | cut -c1-220
