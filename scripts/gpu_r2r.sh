#!/bin/bash
# round-2 call R: the argmax tests that used generation 3, compute-sanitizer memcheck / racecheck on generation 6 + the int8 trailing update
O=gpurun_out/${1:-r2r}; mkdir -p $O
timeout 1200 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_fast_argmax_is_exact" > $O/test_argmax.log 2>&1; echo "argmax tests rc=$?"; tail -3 $O/test_argmax.log
timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_fast.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_fast.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "Race reported|SUMMARY" $O/sanitizer_racecheck.log | cut -c1-200 | sort | uniq -c | head -20
