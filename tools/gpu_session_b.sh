#!/bin/bash
# Scoring-kernel A/B session: tests, strong + weak bench with and without the tier-1 slot rotation, one full ncu capture.
tag=${1:-s}
mkdir -p gpurun_out
F="--no-e2e --no-cpu-baseline --no-other-configs"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 2 --warmup 1 $F > gpurun_out/${tag}_strong.json 2> gpurun_out/${tag}_strong.err; echo "rc=$?" >> gpurun_out/${tag}_strong.err
timeout 600 python bench.py --scaling weak --steps 3 --warmup 2 $F > gpurun_out/${tag}_weak.json 2> gpurun_out/${tag}_weak.err; echo "rc=$?" >> gpurun_out/${tag}_weak.err
PD_T1ROT=0 timeout 600 python bench.py --scaling weak --steps 3 --warmup 2 $F > gpurun_out/${tag}_weak_rot0.json 2> gpurun_out/${tag}_weak_rot0.err
PD_T1ROT=3 timeout 600 python bench.py --scaling weak --steps 3 --warmup 2 $F > gpurun_out/${tag}_weak_rot3.json 2> gpurun_out/${tag}_weak_rot3.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:score_rows_kernel -s 6 -c 1 -o gpurun_out/${tag}_score512 \
  python bench.py --scaling weak --steps 1 --warmup 0 $F --no-parity > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"onesweep_kernel|kmer_hist_kernel" -c 4 -o gpurun_out/${tag}_sort \
  python bench.py --scaling weak --query-genomes 4 --steps 1 --warmup 0 $F --no-parity > gpurun_out/${tag}_ncu_sort.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
for f in strong weak weak_rot0 weak_rot3; do grep "step " gpurun_out/${tag}_$f.err | tail -1; done
