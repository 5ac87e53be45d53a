#!/bin/bash
timeout 600 python -m pytest tests/test_fit_gpu.py tests/test_parity_gpu.py tests/test_restricted_gpu.py tests/test_trend_gpu.py tests/test_multi_target_gpu.py tests/test_genexp_gpu.py tests/test_grad_gpu.py -q -x --timeout 300 2>&1 | tail -4
for BT in ${BTS:-0 1}; do for W in C3 C4 C2 C5; do
B200BO_BIG_TILES=$BT timeout 300 python scripts/fit_time.py $W 8 2>&1 | tail -1 | sed "s/^/big_tiles=$BT /"
done; done
