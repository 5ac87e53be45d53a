#!/bin/bash
# round-2 call Y: generation 6 with the first K chunks of a tile double-buffered in the scratch: parity + sweep
O=gpurun_out/${1:-r2y}; mkdir -p $O
B200BO_GEN6_DB_CHUNKS=16 timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_rt_and_moments and 6" > $O/test_rt.log 2>&1; echo "rt tests (K=16) rc=$?"; tail -2 $O/test_rt.log
B200BO_GEN6_DB_CHUNKS=32 timeout 900 python -m pytest tests/test_scale_gpu.py -x -q -m gpu > $O/test_scale.log 2>&1; echo "scale tests (K=32) rc=$?"; tail -2 $O/test_scale.log
for K in 0 16 32 64 0 24; do B200BO_GEN6_DB_CHUNKS=$K timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_K$K.json 2> $O/bench_K$K.err
  python - <<PY
import json
d=json.loads(open('$O/bench_K$K.json').read().strip().splitlines()[-1])
print('db_chunks=$K (ran %s) value %.4e e2e %.4e ms %.2f frac %.3f clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
PY
done
B200BO_GEN6_DB_CHUNKS=32 B200BO_TRACE=$O/trace_gen6_K32.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/trace.log 2>&1; echo "trace rc=$?"
