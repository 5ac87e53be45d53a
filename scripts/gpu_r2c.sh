#!/bin/bash
# round-2 call C: whole-tree check after the re-entry -- smoke, every GPU test (no -x: list every failure), default bench,
# reference arm, launch list of a step
O=gpurun_out/${1:-r2c}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/box.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
S=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -15 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 6000 $O/bench_default.json; tail -5 $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; tail -c 1500 $O/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_step.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"; tail -3 $O/ncu_list.log
