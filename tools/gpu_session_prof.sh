#!/bin/bash
# profiles of the final code: launch list of one strong N=1 step, full captures of the scoring kernel and the sort kernels
tag=${1:-prof}
mkdir -p gpurun_out
F="--no-e2e --no-cpu-baseline --no-other-configs --no-parity"
timeout 600 python bench.py --steps 2 --warmup 1 $F > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "rc=$?" >> gpurun_out/${tag}_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 $F > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:score_rows_kernel -s 6 -c 1 -o gpurun_out/${tag}_score512 \
  python bench.py --steps 1 --warmup 0 $F > gpurun_out/${tag}_ncu_score.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"onesweep_kernel|kmer_hist_kernel|entry_apply_kernel|fwd_partition_kernel|fwd_place_kernel" -c 8 -o gpurun_out/${tag}_build \
  python bench.py --scaling weak --query-genomes 4 --steps 1 --warmup 0 $F > gpurun_out/${tag}_ncu_build.log 2>&1
grep "step " gpurun_out/${tag}_plain.err | tail -2; ls -la gpurun_out/${tag}_*
