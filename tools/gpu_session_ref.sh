#!/bin/bash
# the reference arm on the full 1,000-genome index (host RAM / time check) and the default N=1 bench line
tag=${1:-ref}
mkdir -p gpurun_out
timeout 1700 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err
echo "rc=$?" >> gpurun_out/${tag}_reference.err
grep "reference\|Maximum resident\|Elapsed\|rc=" gpurun_out/${tag}_reference.err | tail -12; head -c 1500 gpurun_out/${tag}_reference.json; echo
