#!/bin/bash
# round-2 call 3F: launches sized in time (tiles per launch scaled by (4096 / N)^2): C2 / C5 / C3
O=gpurun_out/${1:-r3f}; mkdir -p $O
for W in C2 C5 C3; do timeout 600 python bench.py --workload $W --steps 8 --warmup 3 --no-extras > $O/bench_$W.json 2> $O/bench_$W.err; python - <<PY
import json
d=json.loads(open('$O/bench_$W.json').read().strip().splitlines()[-1])
print('$W (ran gen %s) value %.4e e2e %.4e ms %.2f e2e ms %.2f frac %.3f launches %s kernels %s clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['kernel_ms_per_step'], d['clocks']))
PY
done
timeout 600 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "device_resident or flat or predict_tol" 2>&1 | tail -2
