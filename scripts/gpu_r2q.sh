#!/bin/bash
# round-2 call Q: whole-tree check: smoke, every GPU test, default bench + reference arm, chunk-size A/B for generation 6
O=gpurun_out/${1:-r2q}; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
S=$(date +%s); timeout 1700 python -m pytest tests -q -m gpu --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -8 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]); print('C3 value=%.4g e2e=%.4g ms=%.2f frac=%.3f whole=%.3f clocks=%s cpu=%s launches=%s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['clocks'], d['cpu_baseline']['value'], d['gpu_launches']))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
for CT in 4 16 32; do B200BO_DEV_CHUNK_TILES=$CT timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_chunk$CT.json 2> $O/bench_chunk$CT.err; python - <<PY
import json
d=json.loads(open('$O/bench_chunk$CT.json').read().strip().splitlines()[-1])
print('chunk_tiles=$CT value %.4e e2e %.4e ms %.2f clocks %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']))
PY
done
