#!/bin/bash
# round-2 call 3B: packed (FFMA2) against scalar fp32 producers in generation 6: two builds on the same box
O=gpurun_out/${1:-r3b}; mkdir -p $O
run() { for i in 1 2; do timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_$1_$i.json 2> $O/bench_$1_$i.err
  python - <<PY
import json
d=json.loads(open('$O/bench_$1_$i.json').read().strip().splitlines()[-1])
print('$1 run $i value %.4e ms %.2f frac %.3f clocks %s'%(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
PY
done; }
run packed
B200BO_EXTRA_DEFINES=B200BO_FK6_SCALAR python -m bayesian_optimization_b200.build --force > $O/build_scalar.log 2>&1; echo "build rc=$?"
run scalar
python -m bayesian_optimization_b200.build --force > $O/build_packed.log 2>&1
run packed2
