#!/bin/bash
# ncu full capture (with source counters) of the generation-5 kernel, one product
O=${1:-gpurun_out/prof5}; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 1 -c 1 -o $O/prof_gen5 \
  python bench.py --steps 1 --warmup 1 --m-per-gpu 75776 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py $O/prof_gen5.ncu-rep > $O/gen5_ncu_summary.txt 2>&1; cat $O/gen5_ncu_summary.txt
