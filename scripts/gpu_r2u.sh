#!/bin/bash
# round-2 call U: is the L2 hit rate of the scratch a capacity effect?  ncu of generations 5 and 6 at N = 2048 (C5) -- scratch 74 / 37 MB
O=gpurun_out/${1:-r2u}; mkdir -p $O
for GEN in 6 5; do B200BO_FAST_KERNEL=$GEN timeout 900 ncu --set full --clock-control none -k regex:predict_fused -s 1 -c 1 -o $O/prof_C5_gen$GEN python bench.py --workload C5 --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/ncu_C5_gen$GEN.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py $O/prof_C5_gen$GEN.ncu-rep > $O/C5_gen${GEN}_ncu_summary.txt 2>&1; grep -E "kernel|duration|tensor|hit_rate|dram__bytes|lts__through" $O/C5_gen${GEN}_ncu_summary.txt; done
