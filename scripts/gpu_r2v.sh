#!/bin/bash
# round-2 call V: generation 6 against generation 5 by N (whole-step bench of the four workloads + N = 3072 / 6144 probes)
O=gpurun_out/${1:-r2v}; mkdir -p $O
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_rt_and_moments and 6" > $O/test_rt.log 2>&1; echo "rt tests rc=$?"; tail -2 $O/test_rt.log
for W in C2 C5 C4; do for GEN in 5 6; do
  B200BO_GEN6_MIN_LD=1024 B200BO_FAST_KERNEL=$GEN timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_${W}_gen$GEN.json 2> $O/bench_${W}_gen$GEN.err
  python - <<PY
import json
d=json.loads(open('$O/bench_${W}_gen$GEN.json').read().strip().splitlines()[-1])
print('$W gen=$GEN (ran %s) value %.4e e2e %.4e ms %.2f frac %.3f clocks %s'%(d['roofline'].get('generation'), d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
PY
done; done | tee $O/gen6_vs_gen5_by_N.txt
