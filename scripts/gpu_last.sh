#!/bin/bash
# last sanity check of the committed tree: smoke, the at-scale fast-path tests, update / oz tests, default bench
O=gpurun_out/${1:-last}; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python -m pytest tests/test_scale_gpu.py tests/test_update_gpu.py tests/test_oz_gpu.py tests/test_parity_gpu.py -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_C3.json').read().strip().splitlines()[-1]); print('C3 value=%.4g e2e=%.4g ms=%.2f frac=%.3f whole=%.3f clocks=%s cpu=%s launches=%s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['clocks'], d['cpu_baseline']['value'], d['gpu_launches']))"
