#!/bin/bash
# One GPU-box session: probe the host, run the GPU tests, the default bench, and an ncu launch list. Output under gpurun_out/.
# usage: tools/gpu_session.sh <tag> [tests|notests]
tag=${1:-s}
mkdir -p gpurun_out
{
  echo "== host"; nproc; free -g | head -2; (java -version 2>&1 | head -1) || true; nvidia-smi -L
} > gpurun_out/${tag}_host.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
fi
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs > gpurun_out/${tag}_ncu.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log 2>/dev/null; tail -3 gpurun_out/${tag}_bench.err; head -c 600 gpurun_out/${tag}_bench.json
