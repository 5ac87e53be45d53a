#!/bin/bash
# round-end check: smoke, all GPU tests, default bench
O=${1:-gpurun_out/final}; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
S=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_C3.json').read().strip().splitlines()[-1]); print('C3 value=%.4g e2e=%.4g ms=%.2f frac=%.3f clocks=%s cpu=%s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value']))"
