#!/bin/bash
# round-2 call W: mailbox split over two lanes: parity subset, C3 / C5 / C4 A/B
O=gpurun_out/${1:-r2w}; mkdir -p $O
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_rt_and_moments and 6" > $O/test_rt.log 2>&1; echo "rt tests rc=$?"; tail -2 $O/test_rt.log
timeout 900 python -m pytest tests/test_scale_gpu.py -x -q -m gpu > $O/test_scale.log 2>&1; echo "scale tests rc=$?"; tail -2 $O/test_scale.log
for CFG in "C3 6" "C3 5" "C3 6" "C5 6" "C5 5" "C2 6" "C2 5"; do set -- $CFG
  B200BO_GEN6_MIN_LD=1024 B200BO_FAST_KERNEL=$2 timeout 600 python bench.py --workload $1 --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_$1_gen$2.json 2> $O/bench_$1_gen$2.err
  python - <<PY
import json
d=json.loads(open('$O/bench_$1_gen$2.json').read().strip().splitlines()[-1])
print('$1 gen=$2 (ran %s) value %.4e e2e %.4e ms %.2f frac %.3f clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
PY
done
