#!/bin/bash
# round-2 call N: developer timeline of generation 6 (CTA 0: MMA issuer + producer group 0), gen 5 for comparison
O=gpurun_out/${1:-r2n}; mkdir -p $O
B200BO_TRACE=$O/trace_gen6.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/trace6.log 2>&1; echo "rc=$?"
B200BO_FAST_KERNEL=5 B200BO_TRACE=$O/trace_gen5.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/trace5.log 2>&1; echo "rc=$?"
ls -la $O; head -5 $O/trace_gen6.txt
