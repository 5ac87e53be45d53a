#!/bin/bash
# round-2 call D: kstar_small fix (append test), the tensor-core trailing update: unit check, timings, fit tests, fit timing A/B
O=gpurun_out/${1:-r2d}; mkdir -p $O
timeout 600 python -m pytest tests/test_update_gpu.py -x -q -m gpu > $O/test_update.log 2>&1; echo "update tests rc=$?"; tail -3 $O/test_update.log
timeout 300 python scripts/oz_check.py 64 192 1024 3968 8064 --out $O/oz_check.json > $O/oz_check.log 2>&1; echo "oz_check rc=$?"; tail -20 $O/oz_check.log
timeout 900 python -m pytest tests/test_oz_gpu.py -x -q -m gpu > $O/test_oz.log 2>&1; echo "oz tests rc=$?"; tail -15 $O/test_oz.log
for W in C3 C4; do for TC in 0 7 8; do
  B200BO_CHOL_TC=$TC timeout 300 python scripts/fit_time.py $W 8 2>&1 | tail -1 | sed "s/^/TC=$TC /" | tee -a $O/fit_time.txt
done; done
