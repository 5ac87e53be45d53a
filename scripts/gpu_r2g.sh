#!/bin/bash
O=gpurun_out/${1:-r2g}; mkdir -p $O
for DBG in 0 1 2 3; do echo "== OZ_DEBUG=$DBG"; B200BO_OZ_DEBUG=$DBG timeout 300 python scripts/oz_check.py 3968 2>&1 | grep -E "^3968 " | cut -c1-100; done
for G in 4 8 16; do echo "== OZ_G=$G"; B200BO_OZ_G=$G timeout 300 python scripts/oz_check.py 3968 2>&1 | grep -E "^3968 [78]" | cut -c1-60; done
timeout 300 python scripts/oz_check.py 64 192 1024 3968 8064 --out $O/oz_check.json > $O/oz_check.log 2>&1; echo "oz_check rc=$?"; cut -c1-150 $O/oz_check.log
timeout 900 python -m pytest tests/test_oz_gpu.py -x -q -m gpu > $O/test_oz.log 2>&1; echo "oz tests rc=$?"; tail -5 $O/test_oz.log
for W in C3 C4; do for TC in 0 7 8; do
  B200BO_CHOL_TC=$TC timeout 300 python scripts/fit_time.py $W 8 2>&1 | tail -1 | sed "s/^/TC=$TC /" | tee -a $O/fit_time.txt
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_syrk -s 2 -c 1 -o $O/prof_oz python scripts/oz_check.py 3968 > $O/ncu_oz.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py $O/prof_oz.ncu-rep > $O/oz_ncu_summary.txt 2>&1; cat $O/oz_ncu_summary.txt
