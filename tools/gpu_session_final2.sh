#!/bin/bash
# last N=1 session of round 2: all GPU tests (they now end in netclu_cc -g / pangenes --clus / pandelos.sh), smoke(), a short bench line
tag=${1:-fin2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?" >> gpurun_out/${tag}_bench.err
tail -4 gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_smoke.log; grep "step \|parity\|rc=" gpurun_out/${tag}_bench.err | tail -4 | cut -c1-220
