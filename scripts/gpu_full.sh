#!/bin/bash
# full GPU check of a round state: smoke, all GPU tests, default bench (+ reference arm), other workloads, ncu launch list
O=${1:-gpurun_out/full}; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
S=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench rc=$?"; cat $O/bench_C3.json
for W in C2 C4 C5; do timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$W.json 2> $O/bench_$W.err; echo "bench $W rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_$W.json').read().strip().splitlines()[-1]); print('$W value=%.4g e2e=%.4g ms=%.2f frac=%.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused|band_|kstar|rs_|acq_kernel|argmax|gather|contract" -c 80 --csv --log-file $O/launches_step_gen5.csv \
  python bench.py --steps 2 --warmup 2 --m-per-gpu 303104 --no-cpu-baseline > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
