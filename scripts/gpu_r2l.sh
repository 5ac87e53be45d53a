#!/bin/bash
# round-2 call L: persisting-L2 window over the generation-6 scratch: A/B bench + ncu
O=gpurun_out/${1:-r2l}; mkdir -p $O
for P in 0 1 0 1; do B200BO_L2_PERSIST=$P timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_persist$P.json 2> $O/bench_persist$P.err; python - <<PY
import json
d=json.loads(open('$O/bench_persist$P.json').read().strip().splitlines()[-1])
print('persist=$P value %.4e e2e %.4e ms %.2f fused %.2f band %.2f frac %.3f clocks %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step']['contract_or_fused'], d['kernel_ms_per_step']['acq_argmax_or_band'], d['roofline']['frac'], d['clocks']))
PY
done
B200BO_L2_PERSIST=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 1 -c 1 -o $O/prof_gen6p python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --no-cpu-baseline --no-extras > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py $O/prof_gen6p.ncu-rep > $O/gen6_persist_ncu_summary.txt 2>&1; grep -E "duration|tensor|hit_rate|dram__bytes|lts__throughput" $O/gen6_persist_ncu_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused|band_|kstar|rowdot|acq_kernel|argmax|gather|moments|check|reset" -c 60 --csv --log-file $O/launches_step_gen6.csv python bench.py --steps 1 --warmup 1 --m-per-gpu 1212416 --no-cpu-baseline --no-extras > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
