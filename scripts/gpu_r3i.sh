#!/bin/bash
# round-2 call 3I: timelines of both sides of group 0 (are the sides balanced?)
O=gpurun_out/${1:-r3i}; mkdir -p $O
for C in 0 2; do B200BO_TRACE_CTA=$C B200BO_TRACE=$O/trace_gen6_cta$C.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/trace$C.log 2>&1; echo "trace cta $C rc=$?"; done
