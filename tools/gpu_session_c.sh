#!/bin/bash
# quick session: GPU tests + strong and weak bench lines without the side legs
tag=${1:-s}
mkdir -p gpurun_out
F="--no-e2e --no-cpu-baseline --no-other-configs"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 2 --warmup 1 $F > gpurun_out/${tag}_strong.json 2> gpurun_out/${tag}_strong.err; echo "rc=$?" >> gpurun_out/${tag}_strong.err
timeout 600 python bench.py --scaling weak --steps 3 --warmup 2 $F --no-parity > gpurun_out/${tag}_weak.json 2> gpurun_out/${tag}_weak.err; echo "rc=$?" >> gpurun_out/${tag}_weak.err
tail -3 gpurun_out/${tag}_pytest.log
for f in strong weak; do grep "step " gpurun_out/${tag}_$f.err | tail -1; done
