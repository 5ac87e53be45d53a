#!/bin/bash
# round-2 call K: generation 6 with 256-bit scratch stores: parity subset, bench, ncu
O=gpurun_out/${1:-r2k}; mkdir -p $O
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_rt_and_moments and 6" > $O/test_rt.log 2>&1; echo "rt tests rc=$?"; tail -3 $O/test_rt.log
for CFG in "6 0" "5 0" "6 0"; do set -- $CFG; B200BO_FAST_KERNEL=$1 B200BO_GEN6_DEAD_HINT=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_gen$1_$2.json 2> $O/bench_gen$1_$2.err; python - <<PY
import json
d=json.loads(open('$O/bench_gen$1_$2.json').read().strip().splitlines()[-1])
print('gen=$1 dead_hint=$2 value %.4e e2e %.4e ms %.2f fused %.2f band %.2f frac %.3f clocks %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step']['contract_or_fused'], d['kernel_ms_per_step']['acq_argmax_or_band'], d['roofline']['frac'], d['clocks']))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 1 -c 1 -o $O/prof_gen6 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py $O/prof_gen6.ncu-rep > $O/gen6_ncu_summary.txt 2>&1; cat $O/gen6_ncu_summary.txt
for W in C2 C4 C5; do timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-extras > $O/bench_$W.json 2> $O/bench_$W.err; echo "bench $W rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_$W.json').read().strip().splitlines()[-1]); print('$W value=%.4g e2e=%.4g ms=%.2f frac=%.3f cpu=%s clocks=%s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline'], d['clocks']))"; done
