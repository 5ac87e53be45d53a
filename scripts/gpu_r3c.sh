#!/bin/bash
# round-2 call 3C: generation 5 with packed fp32 producers (parity, C2 / C5 bench), whole fast-path test file
O=gpurun_out/${1:-r3c}; mkdir -p $O
timeout 1500 python -m pytest tests/test_fast_gpu.py tests/test_scale_gpu.py tests/test_candidates.py -q -m gpu > $O/test_fast.log 2>&1; echo "fast tests rc=$?"; tail -3 $O/test_fast.log
for W in C2 C5; do timeout 600 python bench.py --workload $W --steps 8 --warmup 3 --no-extras > $O/bench_$W.json 2> $O/bench_$W.err; python - <<PY
import json
d=json.loads(open('$O/bench_$W.json').read().strip().splitlines()[-1])
print('$W (ran gen %s) value %.4e e2e %.4e ms %.2f frac %.3f cpu %s clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks']))
PY
done
timeout 900 python bench.py --workload C4 --steps 5 --warmup 3 --no-extras > $O/bench_C4.json 2> $O/bench_C4.err; python - <<PY
import json
d=json.loads(open('$O/bench_C4.json').read().strip().splitlines()[-1])
print('C4 (ran gen %s) value %.4e e2e %.4e ms %.2f frac %.3f cpu %s clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks']))
PY
