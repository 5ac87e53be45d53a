#!/bin/bash
# round-2 call T: assembly kernel specialised per correlation id: parity, timing A/B, ncu; SM placement of the generation-6 grid
O=gpurun_out/${1:-r2t}; mkdir -p $O
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_fit_gpu.py tests/test_genexp_gpu.py tests/test_matern_nu_gpu.py tests/test_restricted_gpu.py tests/test_trend_gpu.py -x -q -m gpu > $O/test_fit.log 2>&1; echo "fit tests rc=$?"; tail -3 $O/test_fit.log
for W in C3 C4 C5 C2; do for V in 0 1; do
  B200BO_ASSEMBLE_TMA=$V timeout 300 python scripts/fit_time.py $W 8 2>&1 | tail -1 | sed "s/^/ASSEMBLE_TMA=$V /" | cut -c1-230 | tee -a $O/fit_time_assemble.txt
done; done
B200BO_ASSEMBLE_TMA=1 timeout 600 ncu --set full --clock-control none -k regex:kmat_assemble -s 2 -c 1 -o $O/prof_assemble1 python scripts/fit_time.py C3 4 > $O/ncu_assemble1.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py $O/prof_assemble1.ncu-rep > $O/assemble1_ncu_summary.txt 2>&1; grep -E "kernel|duration|dram|lts__t_sector_hit|registers|warps_active|issue_active" $O/assemble1_ncu_summary.txt
B200BO_SMID_DUMP=$O/smid.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/smid.log 2>&1; echo "smid rc=$?"; head -3 $O/smid.txt
