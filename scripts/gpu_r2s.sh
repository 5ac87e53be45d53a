#!/bin/bash
# round-2 call S: TMA-staged kernel-matrix assembly: parity (fit / state tests), timing A/B, ncu full capture
O=gpurun_out/${1:-r2s}; mkdir -p $O
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_fit_gpu.py tests/test_genexp_gpu.py tests/test_matern_nu_gpu.py tests/test_restricted_gpu.py -x -q -m gpu > $O/test_fit.log 2>&1; echo "fit tests rc=$?"; tail -3 $O/test_fit.log
for W in C3 C4 C5; do for V in 0 2 3; do
  B200BO_ASSEMBLE_TMA=$V timeout 300 python scripts/fit_time.py $W 8 2>&1 | tail -1 | sed "s/^/ASSEMBLE_TMA=$V /" | cut -c1-230 | tee -a $O/fit_time_assemble.txt
done; done
for V in 0 2; do B200BO_ASSEMBLE_TMA=$V timeout 600 ncu --set full --clock-control none -k regex:kmat_assemble -s 2 -c 1 -o $O/prof_assemble$V python scripts/fit_time.py C3 4 > $O/ncu_assemble$V.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py $O/prof_assemble$V.ncu-rep > $O/assemble${V}_ncu_summary.txt 2>&1; grep -E "kernel|duration|dram|fma|alu|lts__t_sector_hit|registers|warps_active|issue_active" $O/assemble${V}_ncu_summary.txt; done
