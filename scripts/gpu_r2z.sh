#!/bin/bash
# round-2 call Z: fewer CTAs (smaller scratch, higher clock under the cap?) for generation 6
O=gpurun_out/${1:-r2z}; mkdir -p $O
for S in 148 136 128 120 148 112; do B200BO_FAST_MAX_SMS=$S timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_S$S.json 2> $O/bench_S$S.err
  python - <<PY
import json
d=json.loads(open('$O/bench_S$S.json').read().strip().splitlines()[-1])
print('max_sms=$S (ran %s) value %.4e e2e %.4e ms %.2f frac %.3f clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
PY
done
