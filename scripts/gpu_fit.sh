#!/bin/bash
timeout 600 python -m pytest tests/test_fit_gpu.py tests/test_parity_gpu.py tests/test_restricted_gpu.py tests/test_trend_gpu.py tests/test_multi_target_gpu.py -q -x --timeout 300 2>&1 | tail -4
for CFG in "0 0" "1 0" "1 1"; do set -- $CFG; for W in C3 C4 C2; do
B200BO_CHOL_LOOKAHEAD=$1 B200BO_GRAPHS=$2 timeout 300 python scripts/fit_time.py $W 10 2>&1 | tail -1
done; done
