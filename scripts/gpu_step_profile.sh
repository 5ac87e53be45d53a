#!/bin/bash
# ncu evidence for one bench step: launch list of the step kernels + one full capture of the fused kernel
O=${1:-gpurun_out/r01d}; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused|band_|kstar|rs_|acq_kernel|argmax|gather|contract" -c 80 --csv --log-file $O/launches_step.csv \
  python bench.py --steps 2 --warmup 2 --m-per-gpu 303104 --no-cpu-baseline > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 1 -c 1 -o $O/prof_fused_v2 \
  python bench.py --steps 1 --warmup 1 --m-per-gpu 37888 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
