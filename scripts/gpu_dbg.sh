#!/bin/bash
# kernel-only timing experiments on the generation-5 kernel: which producer-side activity slows the MMA stream
for B in ${BITS:-0 1 2 4 8 15 0}; do
  B200BO_DEBUG_BITS=$B timeout 300 python scripts/fused_time.py ${WL:-C3} ${M:-303104} 1 3 2>&1 | tail -1
done
