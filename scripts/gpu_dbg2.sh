#!/bin/bash
for WL in C5 C4; do for B in 0 8; do
  B200BO_DEBUG_BITS=$B timeout 300 python scripts/fused_time.py $WL ${M:-303104} 1 3 2>&1 | tail -1
done; done
