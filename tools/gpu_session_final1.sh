#!/bin/bash
# final N=1 session: all GPU tests, the default bench line, one full ncu capture of the scoring kernel, the launch list
tag=${1:-fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 4 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?" >> gpurun_out/${tag}_bench.err
F="--no-e2e --no-cpu-baseline --no-other-configs --no-parity"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:score_rows_kernel -s 6 -c 1 -o gpurun_out/${tag}_score512 \
  python bench.py --steps 1 --warmup 0 $F > gpurun_out/${tag}_ncu_score.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 $F > gpurun_out/${tag}_ncu_list.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log; grep "step \|parity\|e2e pass\|rc=" gpurun_out/${tag}_bench.err | tail -8 | cut -c1-220
