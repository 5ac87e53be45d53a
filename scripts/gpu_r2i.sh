#!/bin/bash
# round-2 call I: generation 6 (two CTA pairs share a candidate tile): parity tests, A/B bench against generation 5, ncu
O=gpurun_out/${1:-r2i}; mkdir -p $O
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "test_rt_and_moments and (6 or 5)" > $O/test_rt.log 2>&1; echo "rt tests rc=$?"; tail -4 $O/test_rt.log
timeout 900 python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "not test_rt_and_moments" > $O/test_fast.log 2>&1; echo "fast tests rc=$?"; tail -4 $O/test_fast.log
timeout 900 python -m pytest tests/test_scale_gpu.py -x -q -m gpu > $O/test_scale.log 2>&1; echo "scale tests rc=$?"; tail -4 $O/test_scale.log
for GEN in 5 6 5 6; do B200BO_FAST_KERNEL=$GEN timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_gen$GEN.json 2> $O/bench_gen$GEN.err; python - <<PY
import json
d=json.loads(open('$O/bench_gen$GEN.json').read().strip().splitlines()[-1])
print('gen=$GEN value %.4e e2e %.4e ms %.2f fused %.2f band %.2f frac %.3f clocks %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step']['contract_or_fused'], d['kernel_ms_per_step']['acq_argmax_or_band'], d['roofline']['frac'], d['clocks']))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 1 -c 1 -o $O/prof_gen6 python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py $O/prof_gen6.ncu-rep > $O/gen6_ncu_summary.txt 2>&1; cat $O/gen6_ncu_summary.txt
