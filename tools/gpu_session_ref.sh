#!/bin/bash
# the reference arm on the full 1,000-genome index (host RAM / time check) and the default N=1 bench line
tag=${1:-ref}
mkdir -p gpurun_out
( /usr/bin/time -v timeout 1700 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference.json ) 2> gpurun_out/${tag}_reference.err
echo "rc=$?" >> gpurun_out/${tag}_reference.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?" >> gpurun_out/${tag}_bench.err
grep "reference\|Maximum resident\|Elapsed\|rc=" gpurun_out/${tag}_reference.err | tail -12; head -c 1500 gpurun_out/${tag}_reference.json; echo
grep "step \|parity\|e2e pass\|rc=" gpurun_out/${tag}_bench.err | tail -12
