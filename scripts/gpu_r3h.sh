#!/bin/bash
# round-2 call 3H: survivor-list capacity by q: C2 back on the one-product pass?
O=gpurun_out/${1:-r3h}; mkdir -p $O
B200BO_BAND_DEBUG=1 timeout 600 python bench.py --workload C2 --steps 3 --warmup 1 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "^band:" | sort | uniq -c | head -8
for W in C2 C5; do timeout 900 python bench.py --workload $W --steps 8 --warmup 3 --no-extras > $O/bench_$W.json 2> $O/bench_$W.err; python - <<PY
import json
d=json.loads(open('$O/bench_$W.json').read().strip().splitlines()[-1])
print('$W (gen %s, %s product) value %.4e e2e %.4e ms %.2f frac %.3f rescored %s kernels %s cpu %s argmax ok %s clocks %s'%(d['roofline'].get('generation'), d['roofline'].get('products_per_mac'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['config'].get('rescored_per_step'), d['kernel_ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline'].get('argmax_matches_gpu'), d['clocks']))
PY
done
timeout 1500 python -m pytest tests/test_fast_gpu.py tests/test_scale_gpu.py tests/test_candidates.py -q -m gpu -k "not test_rt_and_moments" > $O/test_fast.log 2>&1; echo "fast tests rc=$?"; tail -3 $O/test_fast.log
