#!/bin/bash
O=${1:-gpurun_out/gen5}; mkdir -p $O
timeout 600 python -m pytest tests/test_fast_gpu.py -q -x -k "rt_and_moments" --timeout 300 > $O/pytest_rt.log 2>&1; echo "pytest rt rc=$?"; tail -6 $O/pytest_rt.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 600 $O/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print("value=%.3e e2e=%.3e ms=%.2f frac=%.3f clocks=%s rescored=%s argmax=%s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], d["config"].get("rescored_per_step"), d["config"].get("argmax")))
except Exception as e:
    print("parse failed", e)
PY
B200BO_TRACE=$O/trace_gen5.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 75776 --no-cpu-baseline > $O/trace_run.log 2>&1; echo "trace rc=$?"
[ -n "$SKIP_REST" ] || timeout 900 python -m pytest tests/test_fast_gpu.py -q -x -k "not rt_and_moments" --timeout 300 > $O/pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -4 $O/pytest_rest.log
