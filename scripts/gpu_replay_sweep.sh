#!/bin/bash
# replay kernel (generation 4): parity first, then the scratch-budget sweep on the bench workload
O=${1:-gpurun_out/replay}; mkdir -p $O
timeout 600 python -m pytest tests/test_fast_gpu.py -q -x -k "rt_and_moments" --timeout 300 > $O/pytest_rt.log 2>&1; echo "pytest rt rc=$?"; tail -4 $O/pytest_rt.log
timeout 600 python -m pytest tests/test_fast_gpu.py -q -x -k "not rt_and_moments" --timeout 300 > $O/pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -4 $O/pytest_rest.log
for MB in ${SWEEP:-0 32 64 96 160}; do
  B200BO_REPLAY_MB=$MB timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_mb$MB.json 2> $O/bench_mb$MB.err; echo "MB=$MB rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$O/bench_mb$MB.json").read().strip().splitlines()[-1])
    print("MB=$MB value=%.3e e2e=%.3e ms=%.2f frac=%.3f clocks=%s rescored=%s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], d["config"].get("rescored_per_step")))
except Exception as e:
    print("MB=$MB parse failed", e); print(open("$O/bench_mb$MB.err").read()[-1500:])
PY
done
