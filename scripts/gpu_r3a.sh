#!/bin/bash
# round-2 call 3A: packed fp32 (FFMA2) producers in generation 6: parity subset, A/B against generation 5 as the box reference, timeline
O=gpurun_out/${1:-r3a}; mkdir -p $O
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_rt_and_moments and 6" > $O/test_rt.log 2>&1; echo "rt tests rc=$?"; tail -2 $O/test_rt.log
timeout 900 python -m pytest tests/test_scale_gpu.py -x -q -m gpu > $O/test_scale.log 2>&1; echo "scale tests rc=$?"; tail -2 $O/test_scale.log
for GEN in 6 5 6 5; do B200BO_FAST_KERNEL=$GEN timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_gen$GEN.json 2> $O/bench_gen$GEN.err
  python - <<PY
import json
d=json.loads(open('$O/bench_gen$GEN.json').read().strip().splitlines()[-1])
print('gen=$GEN (ran %s) value %.4e e2e %.4e ms %.2f frac %.3f whole %.3f clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['clocks']))
PY
done
B200BO_TRACE=$O/trace_gen6.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/trace.log 2>&1; echo "trace rc=$?"
