#!/bin/bash
# round-2 call 3E: L2 prefetch of the scratch replays P uses ahead (generation 6)
O=gpurun_out/${1:-r3e}; mkdir -p $O
for P in 0 4 8 16 0 12; do B200BO_GEN6_PF_AHEAD=$P timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_P$P.json 2> $O/bench_P$P.err
  python - <<PY
import json
d=json.loads(open('$O/bench_P$P.json').read().strip().splitlines()[-1])
print('pf_ahead=$P value %.4e ms %.2f frac %.3f clocks %s'%(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
PY
done
