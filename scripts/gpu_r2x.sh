#!/bin/bash
# round-2 call X: cooperative launch of generation 6: parity subset + A/B
O=gpurun_out/${1:-r2x}; mkdir -p $O
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_rt_and_moments and 6" > $O/test_rt.log 2>&1; echo "rt tests rc=$?"; tail -2 $O/test_rt.log
timeout 900 python -m pytest tests/test_scale_gpu.py -x -q -m gpu > $O/test_scale.log 2>&1; echo "scale tests rc=$?"; tail -2 $O/test_scale.log
for COOP in 1 0 1 0; do B200BO_GEN6_COOPERATIVE=$COOP timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_coop$COOP.json 2> $O/bench_coop$COOP.err
  python - <<PY
import json
d=json.loads(open('$O/bench_coop$COOP.json').read().strip().splitlines()[-1])
print('cooperative=$COOP (ran %s) value %.4e e2e %.4e ms %.2f frac %.3f clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
PY
done
